// zkattest.cu — host orchestration + C ABI (include/zkattest.h) of libzkattest.
//
// One context = one CUDA device + one stream + a grow-only workspace.  A prove/verify call
// is a fixed sequence of ~25 batch kernels over flat arrays in HBM; there is no CPU compute
// path (the ZKA_HOSTSIM build of this file exists only for the CPU unit tests, see
// zk_launch.cuh).
#include "zkattest.h"

#include <math.h>
#include <stdio.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <mutex>
#include <chrono>
#include <thread>
#include <string>
#include <vector>

#include "zk_launch.cuh"
#include "zk_prove.cuh"
#include "zk_verify_agg.cuh"

namespace zk {
// ---- small helper tasks of the C ABI layer (namespace scope: kernel template arguments) ----
struct ParsePointsTask {   // host bytes -> affine Montgomery (+ validity), used for params / sub-ops
  const uint8_t* nist;   // [count][65] or null
  const uint8_t* tom;    // [count][67] or null
  uint32_t* nist_aff;    // [count][16]
  uint32_t* tom_aff;     // [count][18]
  uint8_t* bad;          // [count]
  uint8_t* inf;          // [count] (P-256 identity given as 65 zero bytes)
  ZK_HD void operator()(int t) const {
    if (nist) {
      const uint8_t* b = nist + (size_t)t * 65;
      uint32_t x[8], y[8];
      limbs_from_be<8>(x, b + 1, 32);
      limbs_from_be<8>(y, b + 33, 32);
      bool allz = (b[0] == 0) && is_zero_n<8>(x) && is_zero_n<8>(y);
      reduce_once<FpP256>(x);
      reduce_once<FpP256>(y);
      P256Aff a;
      P256p::to_mont(a.x, x);
      P256p::to_mont(a.y, y);
      bool ok = allz || (b[0] == 0x04 && p256_on_curve(a.x, a.y));
      if (!ok || allz) p256_set_generator(a);
      p256_st_aff(nist_aff + (size_t)t * 16, a);
      if (bad) bad[t] = ok ? 0 : 1;
      if (inf) inf[t] = allz ? 1 : 0;
    }
    if (tom) {
      const uint8_t* b = tom + (size_t)t * WP;
      uint32_t xm[PGL], ym[PGL];
      bool ok = tom_parse(xm, ym, b);   // tag, coordinate range and curve equation (edwards.ts:70-86 / weier.ts:74-89)
      if (!ok) {
        TomPt g;
        tom_set_generator(g);
        copy_n<PGL>(xm, g.x);
        copy_n<PGL>(ym, g.y);
      }
      st<PGL>(tom_aff + (size_t)t * TOM_AFF_WORDS, xm);
      st<PGL>(tom_aff + (size_t)t * TOM_AFF_WORDS + PGL, ym);
      if (bad) bad[t] = ok ? 0 : 1;
    }
  }
};

struct GenAffTask {   // generator constants -> device
  uint32_t* p256_g;   // [16]
  uint32_t* tom_g;    // [18]
  ZK_HD void operator()(int) const {
    P256Aff g;
    p256_set_generator(g);
    p256_st_aff(p256_g, g);
    TomPt tg;
    tom_set_generator(tg);     // affine generator of the proof group (z = 1)
    st<PGL>(tom_g, tg.x);
    st<PGL>(tom_g + PGL, tg.y);
  }
};

template <class F, int NB>
struct FieldOpTask {
  const uint8_t *a, *b;
  uint8_t* out;
  int op;
  ZK_HD void operator()(int t) const {
    constexpr int N = F::N;
    uint32_t x[N], y[N], r[N];
    limbs_from_be<N>(x, a + (size_t)t * NB, NB);
    zero_n<N>(y);
    if (b) limbs_from_be<N>(y, b + (size_t)t * NB, NB);
    F::to_mont(x, x);
    F::to_mont(y, y);
    if (op == 0) F::mul(r, x, y);
    else if (op == 1) F::add(r, x, y);
    else if (op == 2) F::sub(r, x, y);
    else if (op == 3) F::inv(r, x);
    else F::inv_fermat(r, x);
    F::from_mont(r, r);
    limbs_to_be<N>(out + (size_t)t * NB, r, NB);
  }
};
struct GProjTask {
  const uint32_t* g;
  uint32_t* proj;
  ZK_HD void operator()(int) const {
    uint32_t one[PGL];
    PGp::set_one(one);
    st<PGL>(proj, g);
    st<PGL>(proj + PGL, g + PGL);
    st<PGL>(proj + 2 * PGL, one);
  }
};

struct CommitConvTask {
  const uint8_t *v, *r;
  uint32_t *jv, *jr;
  ZK_HD void operator()(int t) const {
    uint32_t a[8];
    limbs_from_be<8>(a, v + (size_t)t * 32, 32);
    reduce_once<FpP256>(a);
    st<8>(jv + (size_t)t * 8, a);
    limbs_from_be<8>(a, r + (size_t)t * 32, 32);
    reduce_once<FpP256>(a);
    st<8>(jr + (size_t)t * 8, a);
  }
};

struct PackTomTask {
  const uint8_t* bytes;
  uint8_t* out;
  ZK_HD void operator()(int t) const {
    for (int i = 0; i < WP; i++) out[(size_t)t * WP + i] = bytes[(size_t)t * BSTRIDE + i];
  }
};

struct P256MulTask {
  const uint8_t* k;
  const uint32_t* g8;
  const uint32_t* rtab;
  const uint8_t* binf;
  uint32_t* proj;
  ZK_HD void operator()(int t) const {
    uint32_t s[8];
    limbs_from_be<8>(s, k + (size_t)t * 32, 32);
    reduce_once<FnP256>(s);
    P256Pt acc;
    p256_set_identity(acc);
    if (rtab) {
      if (!binf[t]) p256_accum_rtab(acc, rtab + (size_t)t * RT_ENTRIES * P256_AFF_WORDS, s);
    } else {
      p256_accum_fixed(acc, g8, s, 8);
    }
    p256_st_proj(proj + (size_t)t * P256_PROJ_WORDS, acc);
  }
};

struct PackP256Task {
  const uint8_t* bytes;
  uint8_t* out;
  ZK_HD void operator()(int t) const {
    for (int i = 0; i < NP; i++) out[(size_t)t * NP + i] = bytes[(size_t)t * BSTRIDE + i];
  }
};

struct Hash80Task {
  const uint8_t* m;
  size_t stride;
  const uint32_t* len;
  uint8_t* out;
  ZK_HD void operator()(int t) const {
    Sha256 h;
    h.init();
    h.update(m + (size_t)t * stride, (int)len[t]);
    uint32_t c3[3];
    h.final80(c3);
    uint8_t* o = out + (size_t)t * 10;
    o[0] = (uint8_t)(c3[2] >> 8); o[1] = (uint8_t)c3[2];
    for (int i = 0; i < 4; i++) { o[2 + i] = (uint8_t)(c3[1] >> (24 - 8 * i)); o[6 + i] = (uint8_t)(c3[0] >> (24 - 8 * i)); }
  }
};

struct GenConvTask {
  const uint8_t* v;
  uint32_t *jv, *jr;
  ZK_HD void operator()(int) const {
    uint32_t a[8];
    limbs_from_be<8>(a, v, 32);
    reduce_once<FpP256>(a);
    st<8>(jv, a);
    zero_n<8>(a);
    st<8>(jr, a);
  }
};

struct KeyToIntTask {
  const uint32_t* aff;
  const uint8_t *bad, *inf;
  uint8_t* x;
  int32_t* status;
  ZK_HD void operator()(int t) const {
    uint32_t c[8];
    P256p::from_mont(c, aff + (size_t)t * 16);
    limbs_to_be<8>(x + (size_t)t * 32, c, 32);
    status[t] = (bad[t] || inf[t]) ? ZKA_ERR_INVALID_PK : ZKA_OK;
  }
};

// ---- multi-GPU helpers: proofs leave a rank as ONE contiguous block (zka_proofs_pack / zka_proofs_unpack)
ZK_HD uint64_t pack_align16(uint32_t len) { return ((uint64_t)len + 15u) & ~(uint64_t)15u; }
struct PackScanTask {   // off[b] = sum_{i<b} align16(len[i]); one thread (B <= a few thousand per group)
  const uint32_t* len;
  uint64_t* off;        // [B + 1]
  int B;
  ZK_HD void operator()(int) const {
    uint64_t acc = 0;
    for (int b = 0; b < B; b++) { off[b] = acc; acc += pack_align16(len[b]); }
    off[B] = acc;
  }
};
struct PackCopyTask {   // one thread per (proof, 16-byte piece); dir 0: rows -> packed, 1: packed -> rows
  uint8_t* rows;
  size_t stride;
  const uint32_t* len;
  const uint64_t* off;
  uint8_t* packed;
  size_t cap;
  int pieces;           // ceil(stride / 16)
  int dir;
  ZK_HD void operator()(int t) const {
    const int b = t / pieces, j = t % pieces;
    const size_t o = (size_t)16 * j;
    if (o >= len[b]) return;
    const size_t po = (size_t)off[b] + o;
    if (po + 16 > cap) return;   // the caller checks off[B] <= cap
    uint8_t* r = rows + (size_t)b * stride + o;
    uint8_t* q = packed + po;
    if (dir == 0) { for (int i = 0; i < 16; i++) q[i] = o + i < stride ? r[i] : 0; }
    else { for (int i = 0; i < 16; i++) if (o + i < stride) r[i] = q[i]; }
  }
};
struct PackCopy16Task {   // same, both sides 16-byte aligned: one uint4 per thread
  uint8_t* rows;
  size_t stride;
  const uint32_t* len;
  const uint64_t* off;
  uint8_t* packed;
  size_t cap;
  int pieces;
  int dir;
  ZK_HD void operator()(int t) const {
    const int b = t / pieces, j = t % pieces;
    const size_t o = (size_t)16 * j;
    if (o >= len[b]) return;
    const size_t po = (size_t)off[b] + o;
    if (po + 16 > cap) return;
    U4* r = reinterpret_cast<U4*>(rows + (size_t)b * stride + o);
    U4* q = reinterpret_cast<U4*>(packed + po);
    if (dir == 0) *q = *r; else *r = *q;
  }
};

}  // namespace zk

using namespace zk;

namespace {

struct FixedTable {   // positional table of one base point
  DevBuf buf;
  uint32_t* tab = nullptr;
};

}  // namespace

// A lane = one compute stream + two copy streams + its own grow-only workspace.  A prove / verify call cuts
// the batch into chunks and deals them round-robin to the lanes; every lane beyond the first is driven by
// its own host thread for the duration of the call, so the mid-pipeline host synchronisation of one lane
// (the item count after the Fiat-Shamir scan) never stalls the others, the latency-bound stages of one
// chunk (doubling chains, 16 KB hashes, scans) overlap the multiplier-bound kernels of another, and
// host<->device copies of one lane overlap the kernels of the others.
struct Lane {
  Stream st;
  // copy streams + events of the host-buffer pipeline; slot = (lane-local chunk index) & 1
  Stream cs_in, cs_out;
  Event ev_small[2], ev_tape[2], ev_done[2], ev_out[2];
  Stream aux[2];            // verifier: the torsion guard and the P-256 part of the aggregate check run beside the tomEdwards256 MSM
  Event ev_fork, ev_join[2];
  // workspace (grow-only)
  DevBuf w[64];
  DevBuf in[16], out[8];
  DevBuf agg[48];   // workspace of the verifier's chunk-wide aggregate check (zk_verify_agg.cuh)
  std::string err;
};

struct zka_ctx : Lane {
  int device = 0;
  std::vector<Lane*> extra;   // lanes 1 .. nlanes-1 (lane 0 is the context itself)
  int nlanes = 3;             // ZKA_LANES
  DevBuf ring_in, ring_m;     // the ring of the current call (shared by all lanes, read-only while they run)
  Lane& lane(int i) { return i == 0 ? *this : *extra[i - 1]; }
  int tom_w = 22, tom_nwin = 12;   // per base: 12 windows x (2^21 + 1) signed-digit entries x 128 B = 3.2 GB of HBM (ZKA_TOM_W)
                                   // (profiles/window_sweep_r1.txt: w=13 39.9k, w=16 58.8k proofs/s)
  int chunk = 4096;       // largest chunk of a call whose buffers are all device memory (ZKA_CHUNK)
  int host_chunk = 2048;  // chunk size when proofs return to host memory: copies of one chunk overlap the next
  volatile uint32_t* progress = nullptr;   // zka_set_progress: flags[k] = 1 when chunk k of the running prove call is complete
  uint32_t progress_cap = 0;
  int agg = 1;            // verifier: chunk-wide aggregate check before the per-proof MSMs (ZKA_AGG=0 disables it)
  int agg_c = 0;          // window bits of the aggregate MSM (0: chosen from the chunk size; ZKA_AGG_C)
  uint64_t agg_pass = 0, agg_fail = 0;   // chunks decided by the aggregate / sent to the per-proof path (zka_stat)
  std::mutex stat_mu;
  std::mutex copy_mu;     // keeps the copies of one chunk together on the shared copy-in stream (verify)
  int agg_c_last = 0;
  bool tape_split = true; // host tapes travel in two strided copies: the 3 + 4S draws before the challenge, then only the
                          // item / GK draws up to the longest proof of the chunk (ZKA_TAPE_SPLIT=0: one full-stride copy)
  int p256_hw = 20;       // window bits of the P-256 G table and of the per-params NistGroup.h table
                          // (13 windows x 2^20 entries x 64 B = 872 MB each; 16 -> 20 -> 22: PhaseA 13.4 -> 12.9 -> 12.6 ms)
  FixedTable g8;          // P-256 generator, w=8 [32][256][16]
  FixedTable gw;          // P-256 generator, p256_hw-bit windows (prover phase A)
  FixedTable tg;          // tomEdwards256 generator [nwin][2^w][32]
  DevBuf tg_bytes;        // 67-byte encoding of g
  DevBuf lag;             // GK Lagrange matrix cache
  int lag_n = -1;
};

struct zka_params {
  zka_ctx* ctx = nullptr;
  uint32_t sec_level = 80;
  FixedTable h8;          // NistGroup.h fixed-base table, h_w-bit windows (16 x 65536 x 64 B = 67 MB)
  int h_w = 20;
  FixedTable th;          // ProofGroup.h table
  uint8_t h_nist[65];
  uint8_t h_proof[WP];
};

namespace {

int fail(zka_ctx* ctx, int code, const std::string& msg) {
  if (ctx) ctx->err = msg;
  return code;
}

// ---- chunk-wide aggregate check of the verifier (zk_verify_agg.cuh) ------------------------------------------------
struct AggPlan {
  AggDigits D;
  int levels;
  int lm[AGG_MAX_LEVELS];   // log2 fan-in of every level of the bucket reduction (sum = c - 1)
};
// window bits from a cost model fitted to profiles/agg_window_sweep_r2i.md: ceil(258/c) windows x (entries / warp
// efficiency + 3 x 2^(c-1) bucket-tree additions); the warp efficiency accounts for the spread of the bucket sizes inside
// a warp (Poisson: about mean + sigma)
AggPlan agg_plan(double entries, int c_forced) {
  int best = 4;
  double best_cost = 1e300;
  for (int c = 4; c <= 16; c++) {
    const double nb = (double)(1u << (c - 1)), load = entries / nb;
    const double eff = load / (load + std::sqrt(load > 1.0 ? load : 1.0));
    const double cost = std::ceil(258.0 / c) * (entries / (eff > 0.05 ? eff : 0.05) + 3.0 * nb);
    if (cost < best_cost) { best_cost = cost; best = c; }
  }
  const int c = c_forced ? c_forced : best;
  AggPlan pl;
  memset(&pl, 0, sizeof(pl));
  pl.D.c = c;
  pl.D.nwin = (258 + c - 1) / c;
  pl.D.nb = 1 << (c - 1);
  for (int j = 0; j < pl.D.nwin; j++) {   // offs = sum_j 2^(c-1) 2^(c j)
    const int pos = c * j + c - 1;
    pl.D.offs[pos >> 5] |= 1u << (pos & 31);
  }
  // top window: digits 0 .. top_max, spread over 2^top_shift sub-buckets each
  const int tb = 256 - c * (pl.D.nwin - 1);
  const int top_max = 1 << (tb > 0 ? tb : 0);
  pl.D.top_shift = 0;
  while (((top_max + 1) << (pl.D.top_shift + 1)) <= pl.D.nb) pl.D.top_shift++;
  const int bits = c - 1;
  pl.levels = (bits + AGG_FAN_BITS - 1) / AGG_FAN_BITS;
  for (int i = 0; i < pl.levels; i++) pl.lm[i] = bits / pl.levels + (i < bits % pl.levels ? 1 : 0);
  return pl;
}
// enqueue histogram, prefix sums, scatter, bucket sums and the reduction tree of one group; returns the root sums
template <class Src>
void agg_msm(Stream& st, DevBuf* A, const Src& src, const AggPlan& pl, const uint32_t* ctl, const uint32_t** rootA,
             const uint32_t** rootB) {
  const AggDigits& D = pl.D;
  const int nwin = D.nwin, nb = D.nb, nseg = (nb + 1 + AGG_SEG - 1) / AGG_SEG;
  const size_t cap = (size_t)src.slots();
  uint32_t* hist = A[0].get<uint32_t>((size_t)nwin * (nb + 1));
  uint32_t* bstart = A[1].get<uint32_t>((size_t)nwin * (nb + 2));
  uint32_t* segtot = A[2].get<uint32_t>((size_t)nwin * nseg);
  uint32_t* sorted = A[3].get<uint32_t>((size_t)nwin * cap);
  uint32_t* bsum = A[4].get<uint32_t>((size_t)nwin * nb * Src::PTW);
  dev_memset(st, hist, 0, (size_t)nwin * (nb + 1) * 4);
  launch(st, (long long)cap, AggHistTask<Src>{src, D, ctl, hist});
  launch(st, (long long)nwin * nseg, AggSegSumTask{ctl, hist, segtot, nb, nseg});
  launch(st, nwin, AggSegScanTask{ctl, segtot, nseg});
  launch(st, (long long)nwin * nseg, AggOffsetsTask{ctl, segtot, hist, bstart, nb, nseg});
  launch(st, (long long)cap, AggScatterTask<Src>{src, D, ctl, hist, sorted, cap});
  launch(st, (long long)nwin * nb, AggBucketTask<Src>{src, ctl, bstart, sorted, bsum, cap, nb});
  const uint32_t *inA = bsum, *inB = nullptr;
  int nin = nb, ll = 0;
  for (int lv = 0; lv < pl.levels; lv++) {
    const int nout = nin >> pl.lm[lv];
    uint32_t* oA = A[5 + 2 * lv].get<uint32_t>((size_t)nwin * nout * Src::PTW);
    uint32_t* oB = A[6 + 2 * lv].get<uint32_t>((size_t)nwin * nout * Src::PTW);
    launch(st, (long long)nwin * nout, AggLevelTask<Src>{ctl, inA, inB, oA, oB, nin, pl.lm[lv], ll, nwin, D.top_shift});
    inA = oA; inB = oB; nin = nout; ll += pl.lm[lv];
  }
  *rootA = inA;
  *rootB = inB;
}

int ceil_log2(uint32_t v) {
  int n = 0;
  while ((1ull << n) < v) n++;
  return n;
}


// Normalisation launches: points per thread (= per binary inversion).  A thread costs about
// per_point * chunk + 70 multiplications in sequence (the inversion is worth ~70), and the grid runs in
// waves of (SMs x 4 resident CTAs): pick the chunk that minimises waves x thread length.  (The first
// heuristic, count / 100000, put 1.4 waves on the GPU for a 1024-proof batch.)
static int g_norm_slots = 148 * 4;
static bool g_norm_model = true;
inline int norm_chunk_for(long long count, int per_point = 11) {
  if (!g_norm_model) {
    long long c = count / 100000;
    if (c < 8) c = 8;
    if (c > NORM_CHUNK_MAX) c = NORM_CHUNK_MAX;
    return (int)c;
  }
  int best = 8;
  long long best_cost = -1;
  for (int c = 8; c <= NORM_CHUNK_MAX; c++) {
    const long long threads = (count + c - 1) / c, ctas = (threads + 127) / 128;
    const long long waves = (ctas + g_norm_slots - 1) / g_norm_slots;
    const long long cost = waves * ((long long)per_point * c + 70);
    if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = c; }
  }
  return best;
}
inline void launch_p256_norm(Stream& st, const uint32_t* proj, uint32_t* aff, uint8_t* bytes, uint8_t* inf, long long count) {
  if (count <= 0) return;
  const int ch = norm_chunk_for(count, bytes ? 7 : 5);
  launch(st, (count + ch - 1) / ch, P256NormTask{proj, aff, bytes, inf, (int)count, ch});
}
// e2 = 1: the points come from TomCommitTask (a = -1 image curve E2); e2 = 0: E1 projective
// aff may be null; otherwise the E1 affine pair is written for points with (index % aff_mod) < aff_lim
inline void launch_tom_norm(Stream& st, const uint32_t* proj, uint32_t* aff, uint8_t* bytes, long long count, int e2,
                            int aff_mod = 1, int aff_lim = 1) {
  if (count <= 0) return;
  const int ch = norm_chunk_for(count);
  launch(st, (count + ch - 1) / ch, TomNormTask{proj, aff, bytes, (int)count, ch, e2, aff_mod, aff_lim});
}

// ---- table construction -------------------------------------------------------------------
// P-256 positional table (signed digits, fb_entries(w) multiples per window) from one affine Montgomery base
// (device pointer, 16 words)
void build_p256_tab(zka_ctx* ctx, const uint32_t* base_aff_dev, FixedTable& out, int w) {
  Stream& st = ctx->st;
  const int nwin = fb_windows(w);
  const size_t E = (size_t)fb_entries(w), count = (size_t)nwin * E;
  DevBuf pows, rows;
  uint32_t* d_pows = pows.get<uint32_t>((size_t)nwin * P256_PROJ_WORDS);
  uint32_t* d_rows = rows.get<uint32_t>(count * P256_PROJ_WORDS);
  out.tab = out.buf.get<uint32_t>(count * P256_AFF_WORDS);
  launch(st, 1, P256PowsTask{base_aff_dev, nullptr, d_pows, 1, nwin, w});
  if (w > 9) {
    DevBuf hi;
    const int nh = 1 << (w - 9);
    uint32_t* d_hi = hi.get<uint32_t>((size_t)nwin * nh * P256_PROJ_WORDS);
    launch(st, nwin, P256RowsHiTask{d_pows, d_hi, d_rows, w});
    launch(st, (long long)nwin * nh, P256RowsLoTask{d_pows, d_hi, d_rows, w});
    sync(st);
    hi.release();
  } else {
    launch(st, nwin, P256RowsTask{d_pows, d_rows, w});
  }
  launch_p256_norm(st, d_rows, out.tab, nullptr, nullptr, (long long)(count));
  sync(st);
  pows.release();
  rows.release();
}
#if defined(ZKA_PG_WAR256)
// war256 positional table [fb_windows(w)][fb_entries(w)] of affine points from one affine base (16 words, device)
void build_tom_tab(zka_ctx* ctx, const uint32_t* base_aff_dev, FixedTable& out) {
  Stream& st = ctx->st;
  const int w = ctx->tom_w, nwin = fb_windows(w);
  const size_t E = (size_t)fb_entries(w), count = (size_t)nwin * E;
  DevBuf pows, rows;
  uint32_t* d_pows = pows.get<uint32_t>((size_t)nwin * P256_PROJ_WORDS);
  uint32_t* d_rows = rows.get<uint32_t>(count * P256_PROJ_WORDS);
  out.tab = out.buf.get<uint32_t>(count * TOM_PRE_WORDS);
  launch(st, 1, WarPowsTask{base_aff_dev, nullptr, d_pows, 1, nwin, w});
  if (w > 9) {
    DevBuf hi;
    const int nh = 1 << (w - 9);
    uint32_t* d_hi = hi.get<uint32_t>((size_t)nwin * nh * P256_PROJ_WORDS);
    launch(st, nwin, WarRowsHiTask{d_pows, d_hi, d_rows, w});
    launch(st, (long long)nwin * nh, WarRowsLoTask{d_pows, d_hi, d_rows, w});
    sync(st);
    hi.release();
  } else {
    launch(st, nwin, WarRowsTask{d_pows, d_rows, w});
  }
  {
    const int ch = norm_chunk_for((long long)count, 5);
    launch(st, ((long long)count + ch - 1) / ch, WarNormTask{d_rows, out.tab, nullptr, nullptr, (int)count, ch});
  }
  sync(st);
  pows.release();
  rows.release();
}
#else
// tomEdwards256 positional table [nwin][fb_entries(w)] from one image-curve affine base (18 words, device)
void build_tom_tab(zka_ctx* ctx, const uint32_t* base_aff_dev, FixedTable& out) {
  Stream& st = ctx->st;
  const int w = ctx->tom_w, nwin = ctx->tom_nwin;
  const size_t ne = (size_t)fb_entries(w), count = (size_t)nwin * ne;
  DevBuf pows, rows;
  uint32_t* d_pows = pows.get<uint32_t>((size_t)nwin * 36);
  uint32_t* d_rows = rows.get<uint32_t>(count * TOM_PROJ_WORDS);
  out.tab = out.buf.get<uint32_t>(count * TOM_PRE_WORDS);
  launch(st, 1, TomPowsTask{base_aff_dev, d_pows, 1, nwin, w});
  if (w > 9) {
    DevBuf hi;
    const int nh = 1 << (w - 9);
    uint32_t* d_hi = hi.get<uint32_t>((size_t)nwin * nh * 36);
    launch(st, nwin, TomRowsHiTask{d_pows, d_hi, d_rows, w});
    launch(st, (long long)nwin * nh, TomRowsLoTask{d_pows, d_hi, d_rows, w});
    sync(st);
    hi.release();
  } else {
    launch(st, nwin, TomRowsTask{d_pows, d_rows, w});
  }
  // rows (E1 projective) -> entries of the prover's a = -1 image curve (v - w, v + w, 2 d2 w v)
  launch(st, (long long)(count + 15) / 16, TomTabE2Task{d_rows, out.tab, (int)count});
  sync(st);
  pows.release();
  rows.release();
}

#endif

// stage a caller buffer on the device if it is a host pointer
template <class T>
const T* stage_in(Stream& st, DevBuf& buf, const T* p, size_t count) {
  if (!p || count == 0) return p;
  if (is_device_ptr(p)) return p;
  T* d = buf.get<T>(count);
  copy_h2d(st, d, p, count * sizeof(T));
  return d;
}
template <class T>
const T* stage_in(zka_ctx* ctx, DevBuf& buf, const T* p, size_t count) {
  return stage_in(ctx->st, buf, p, count);
}

// Chunk schedule of a call: boundaries off[0..nchunks] of the batch.  `cmax` = largest chunk, `lanes` = lanes that
// will run.  Plain: near-equal chunks, at least one per lane when chunks of >= 256 proofs allow it.  Tapered (used
// when a buffer lives in host memory): the first chunks are small so that the kernels start after a short input
// copy, the last ones small so that little output copy is left exposed after the last kernel, and lanes that
// claim chunks dynamically drift out of phase (one copies while another computes):
//   cmax/4, cmax/2, [full chunks], cmax/2, cmax/4.
std::vector<uint32_t> chunk_schedule(uint32_t B, uint32_t cmax, int lanes, bool taper) {
  std::vector<uint32_t> off{0};
  auto r32 = [](uint32_t v) { return v < 32 ? std::max<uint32_t>(v, 1) : ((v + 31) & ~31u); };
  if (lanes > 1) {
    uint32_t per = r32((B + (uint32_t)lanes - 1) / (uint32_t)lanes);
    cmax = std::min(cmax, std::max<uint32_t>(per, 256));
  }
  uint32_t done = 0;
  auto push = [&](uint32_t n) { n = std::min(n, B - done); if (n) { done += n; off.push_back(done); } };
  if (taper && cmax >= 1024 && B >= 3 * cmax) {
    const uint32_t q = r32(cmax / 4), h = r32(cmax / 2);
    push(q);
    push(h);
    const uint32_t mid = B - done - (h + q);
    const uint32_t nm = (mid + cmax - 1) / cmax;
    const uint32_t each = r32((mid + nm - 1) / nm);
    for (uint32_t i = 0; i + 1 < nm; i++) push(each);
    push(B - done - (h + q));
    push(h);
    push(q);
  } else {
    const uint32_t nk = (B + cmax - 1) / cmax;
    const uint32_t each = r32((B + nk - 1) / nk);
    while (done < B) push(each);
  }
  return off;
}

// flags[k] = 1 once everything enqueued on `st` so far has run (a CUDA host callback: no thread of ours waits)
#if !defined(ZKA_HOSTSIM)
void CUDART_CB progress_cb(void* p) { *reinterpret_cast<volatile uint32_t*>(p) = 1u; }
void notify_progress(Stream& st, volatile uint32_t* flag) { ZK_CUDA_CHECK(cudaLaunchHostFunc(st.s, progress_cb, (void*)flag)); }
#else
void notify_progress(Stream&, volatile uint32_t* flag) { *flag = 1u; }
#endif

// Run fn(lane index) for lanes 0..used-1: lane 0 on the calling thread, the others on their own host threads
// (each binds the context's device).  The first exception of any lane is rethrown on the caller.
template <class Fn>
void run_lanes(zka_ctx* ctx, int used, Fn fn) {
  if (used <= 1) { fn(0); return; }
  std::vector<std::thread> th;
  std::vector<std::string> errs((size_t)used);
  for (int li = 1; li < used; li++)
    th.emplace_back([&, li] {
      try {
#if !defined(ZKA_HOSTSIM)
        ZK_CUDA_CHECK(cudaSetDevice(ctx->device));
#endif
        fn(li);
      } catch (const std::exception& e) {
        errs[(size_t)li] = e.what()[0] ? e.what() : "lane failed";
      }
    });
  try { fn(0); } catch (const std::exception& e) { errs[0] = e.what()[0] ? e.what() : "lane failed"; }
  for (auto& t : th) t.join();
  for (auto& e : errs)
    if (!e.empty()) throw std::runtime_error(e);
}

}  // namespace

// =============================================================================== C ABI
extern "C" {

int zka_version(void) { return 1; }

const char* zka_last_error(const zka_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

uint64_t zka_launch_count(const zka_ctx* ctx) {
  if (!ctx) return 0;
  uint64_t n = ctx->st.launches + ctx->aux[0].launches + ctx->aux[1].launches;
  for (const Lane* l : ctx->extra) n += l->st.launches + l->aux[0].launches + l->aux[1].launches;
  return n;
}

void* zka_get_stream(zka_ctx* ctx) {
#if !defined(ZKA_HOSTSIM)
  return ctx ? (void*)ctx->st.s : nullptr;
#else
  return nullptr;
#endif
}
int zka_set_profiling(zka_ctx* ctx, int enable) {
  if (!ctx) return ZKA_E_ARG;
  try {
    for (int i = 0; i < 1 + (int)ctx->extra.size(); i++) { sync(ctx->lane(i).st); ctx->lane(i).st.profiling = enable != 0; }
  } catch (...) { return ZKA_E_CUDA; }
  return 0;
}
int zka_profile_reset(zka_ctx* ctx) {
  if (!ctx) return ZKA_E_ARG;
  try {
    for (int i = 0; i < 1 + (int)ctx->extra.size(); i++) { sync(ctx->lane(i).st); ctx->lane(i).st.prof.clear(); }
  } catch (...) { return ZKA_E_CUDA; }
  return 0;
}
// knobs that may change between calls (tests, sweeps): "lanes", "chunk", "host_chunk"
int zka_set_option(zka_ctx* ctx, const char* key, long value) {
  if (!ctx || !key || value < 1) return ZKA_E_ARG;
  const std::string k(key);
  try {
    if (k == "lanes") {
      if (value > 8) return ZKA_E_ARG;
      while ((int)ctx->extra.size() + 1 < value) {
        Lane* l = new Lane();
        stream_create(l->st);
        stream_create(l->cs_in);
        stream_create(l->cs_out);
        stream_create(l->aux[0]);
        stream_create(l->aux[1]);
        l->st.profiling = ctx->st.profiling;
        ctx->extra.push_back(l);
      }
      ctx->nlanes = (int)value;
    } else if (k == "chunk") {
      ctx->chunk = (int)value;
    } else if (k == "host_chunk") {
      ctx->host_chunk = (int)value;
    } else if (k == "agg") {          // 1: off, 2: on (values start at 1)
      ctx->agg = (int)value - 1;
    } else if (k == "agg_c") {
      if (value < 4 || value > 16) return ZKA_E_ARG;
      ctx->agg_c = (int)value;
    } else {
      return ZKA_E_ARG;
    }
  } catch (const std::exception& e) {
    return fail(ctx, ZKA_E_CUDA, e.what());
  }
  return 0;
}
long long zka_stat(zka_ctx* ctx, const char* key) {
  if (!ctx || !key) return -1;
  const std::string k(key);
  std::lock_guard<std::mutex> g(ctx->stat_mu);
  if (k == "agg_pass") return (long long)ctx->agg_pass;
  if (k == "agg_fail") return (long long)ctx->agg_fail;
  if (k == "agg_c") return (long long)ctx->agg_c_last;        // window bits of the last tomEdwards256 aggregate MSM
  return -1;
}
size_t zka_profile_json(zka_ctx* ctx, char* buf, size_t cap) {
  if (!ctx) return 0;
  std::map<std::string, ProfEntry> all;
  try {
    for (int i = 0; i < 1 + (int)ctx->extra.size(); i++) {
      sync(ctx->lane(i).st);
      for (auto& kv : ctx->lane(i).st.prof) {
        ProfEntry& e = all[kv.first];
        e.launches += kv.second.launches; e.ms += kv.second.ms; e.items += kv.second.items;
      }
    }
  } catch (...) { return 0; }
  std::string j = "{";
  bool first = true;
  for (auto& kv : all) {
    char tmp[512];
    snprintf(tmp, sizeof tmp, "%s\"%s\": {\"launches\": %llu, \"ms\": %.6f, \"items\": %llu}", first ? "" : ", ",
             kv.first.c_str(), (unsigned long long)kv.second.launches, kv.second.ms,
             (unsigned long long)kv.second.items);
    j += tmp;
    first = false;
  }
  j += "}";
  if (buf && cap) {
    size_t n = std::min(cap - 1, j.size());
    memcpy(buf, j.data(), n);
    buf[n] = 0;
  }
  return j.size() + 1;
}
int zka_proof_group(char* name, size_t cap, int* point_bytes, int* scalar_bytes) {
#if defined(ZKA_PG_WAR256)
  const char* n = "war256";
#else
  const char* n = "tomEdwards256";
#endif
  if (name && cap) {
    strncpy(name, n, cap - 1);
    name[cap - 1] = 0;
  }
  if (point_bytes) *point_bytes = WP;
  if (scalar_bytes) *scalar_bytes = WS;
  return 0;
}
int zka_set_progress(zka_ctx* ctx, volatile uint32_t* flags, uint32_t cap) {
  if (!ctx) return ZKA_E_ARG;
  ctx->progress = flags;
  ctx->progress_cap = flags ? cap : 0;
  return 0;
}
int zka_chunk_schedule(zka_ctx* ctx, uint32_t B, int host_buffers, uint32_t* off, uint32_t cap) {
  if (!ctx || !off || cap == 0) return ZKA_E_ARG;
  const std::vector<uint32_t> v = chunk_schedule(B, (uint32_t)(host_buffers ? std::min(ctx->chunk, ctx->host_chunk) : ctx->chunk), ctx->nlanes,
                                                 host_buffers != 0);
  if (v.size() > cap) return ZKA_E_ARG;
  for (size_t i = 0; i < v.size(); i++) off[i] = v[i];
  return (int)v.size() - 1;
}
int zka_lanes(const zka_ctx* ctx) { return ctx ? ctx->nlanes : 0; }
int zka_config(const zka_ctx* ctx, int* tom_w, int* tom_nwin, int* chunk) {
  if (!ctx) return ZKA_E_ARG;
  if (tom_w) *tom_w = ctx->tom_w;
  if (tom_nwin) *tom_nwin = ctx->tom_nwin;
  if (chunk) *chunk = ctx->chunk;
  return 0;
}

int zka_init(int device, zka_ctx** out) {
  if (!out) return ZKA_E_ARG;
  *out = nullptr;
  zka_ctx* ctx = new zka_ctx();
  try {
#if !defined(ZKA_HOSTSIM)
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0 || device >= ndev) {
      cudaGetLastError();
      delete ctx;
      return ZKA_E_CUDA;   // no GPU: fail loudly, there is no CPU fallback
    }
    ZK_CUDA_CHECK(cudaSetDevice(device));
    ZK_CUDA_CHECK(cudaStreamCreateWithFlags(&ctx->st.s, cudaStreamNonBlocking));
#endif
    ctx->device = device;
    if (const char* e = getenv("ZKA_TOM_W")) {
      int w = atoi(e);
      if (w >= 2 && w <= 24) ctx->tom_w = w;
    }
    ctx->tom_nwin = fb_windows(ctx->tom_w);
    if (const char* e = getenv("ZKA_CHUNK")) {
      int c = atoi(e);
      if (c >= 1) ctx->chunk = c;
    }
    if (const char* e = getenv("ZKA_NORM_MODEL")) g_norm_model = atoi(e) != 0;
#if !defined(ZKA_HOSTSIM)
    {
      int sms = 0;
      if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) == cudaSuccess && sms > 0) g_norm_slots = sms * 4;
    }
#endif
    if (const char* e = getenv("ZKA_HOST_CHUNK")) {
      int c = atoi(e);
      if (c >= 1) ctx->host_chunk = c;
    }
    stream_create(ctx->cs_in);
    stream_create(ctx->cs_out);
    stream_create(ctx->aux[0]);
    stream_create(ctx->aux[1]);
    {
      int lanes = 3;
#if defined(ZKA_HOSTSIM)
      lanes = 1;
#endif
      if (const char* e = getenv("ZKA_LANES")) lanes = atoi(e);
      if (lanes < 1) lanes = 1;
      if (lanes > 8) lanes = 8;
      if (zka_set_option(ctx, "lanes", lanes) != 0) throw std::runtime_error("lanes");
    }
    if (const char* e = getenv("ZKA_TAPE_SPLIT")) ctx->tape_split = atoi(e) != 0;
    if (const char* e = getenv("ZKA_AGG")) ctx->agg = atoi(e);
    if (const char* e = getenv("ZKA_AGG_C")) {
      int c = atoi(e);
      if (c >= 4 && c <= 16) ctx->agg_c = c;
    }
    if (const char* e = getenv("ZKA_P256_HW")) {   // window bits of the P-256 G / NistGroup.h tables: 8..24
      int w = atoi(e);
      if (w >= 8 && w <= 24) ctx->p256_hw = w;
    }
    DevBuf gen;
    uint32_t* d_gen = gen.get<uint32_t>(16 + TOM_AFF_WORDS);
    launch(ctx->st, 1, GenAffTask{d_gen, d_gen + 16});
    build_p256_tab(ctx, d_gen, ctx->g8, 8);
    build_p256_tab(ctx, d_gen, ctx->gw, ctx->p256_hw);
    build_tom_tab(ctx, d_gen + 16, ctx->tg);
    // encoding of g (C_14 = params.g in pi_8, pointAdd.ts:144,220): normalise the table entry 1*g
    DevBuf proj, aff;
    uint32_t* d_proj = proj.get<uint32_t>(TOM_PROJ_WORDS);
    uint32_t* d_aff = aff.get<uint32_t>(TOM_AFF_WORDS);
    uint8_t* d_bytes = ctx->tg_bytes.get<uint8_t>(BSTRIDE);
    launch(ctx->st, 1, GProjTask{d_gen + 16, d_proj});
    launch_tom_norm(ctx->st, d_proj, d_aff, d_bytes, 1, 0);
    sync(ctx->st);
    gen.release();
    proj.release();
    aff.release();
  } catch (const std::exception& e) {
    fprintf(stderr, "zka_init: %s\n", e.what());
    delete ctx;
    return ZKA_E_CUDA;
  }
  *out = ctx;
  return 0;
}

void zka_shutdown(zka_ctx* ctx) {
  if (!ctx) return;
  ctx->g8.buf.release();
  ctx->gw.buf.release();
  ctx->tg.buf.release();
  ctx->tg_bytes.release();
  ctx->lag.release();
  for (auto& b : ctx->w) b.release();
  for (auto& b : ctx->in) b.release();
  for (auto& b : ctx->out) b.release();
  for (int li = 0; li < 1 + (int)ctx->extra.size(); li++)
    for (auto& b : ctx->lane(li).agg) b.release();
  ctx->ring_in.release();
  ctx->ring_m.release();
  for (int li = 0; li < 1 + (int)ctx->extra.size(); li++) {
    Lane& l = ctx->lane(li);
    if (li > 0) {
      for (auto& b : l.w) b.release();
      for (auto& b : l.in) b.release();
      for (auto& b : l.out) b.release();
    }
    for (int i = 0; i < 2; i++) {
      ev_destroy(l.ev_small[i]); ev_destroy(l.ev_tape[i]); ev_destroy(l.ev_done[i]); ev_destroy(l.ev_out[i]);
    }
    stream_destroy(l.cs_in);
    stream_destroy(l.cs_out);
    stream_destroy(l.aux[0]);
    stream_destroy(l.aux[1]);
    ev_destroy(l.ev_fork); ev_destroy(l.ev_join[0]); ev_destroy(l.ev_join[1]);
    stream_destroy(l.st);
  }
  for (Lane* l : ctx->extra) delete l;
  delete ctx;
}

int zka_params_create(zka_ctx* ctx, const uint8_t h_nist[65], const uint8_t h_proof[67], uint32_t sec_level,
                      zka_params** out) {
  if (!ctx || !h_nist || !h_proof || !out) return ZKA_E_ARG;
  if (sec_level < 1 || sec_level > MAX_REPS) return fail(ctx, ZKA_E_ARG, "sec_level must be in [1,80]");
  try {
    zka_params* P = new zka_params();
    P->ctx = ctx;
    P->sec_level = sec_level;
    P->h_w = ctx->p256_hw;
    memcpy(P->h_nist, h_nist, 65);
    memcpy(P->h_proof, h_proof, WP);
    DevBuf bn, bt, an, at, bad, inf;
    uint8_t* d_bn = bn.get<uint8_t>(65);
    uint8_t* d_bt = bt.get<uint8_t>(WP);
    uint32_t* d_an = an.get<uint32_t>(16);
    uint32_t* d_at = at.get<uint32_t>(TOM_AFF_WORDS);
    uint8_t* d_bad = bad.get<uint8_t>(2);
    uint8_t* d_inf = inf.get<uint8_t>(1);
    copy_h2d(ctx->st, d_bn, h_nist, 65);
    copy_h2d(ctx->st, d_bt, h_proof, WP);
    launch(ctx->st, 1, ParsePointsTask{d_bn, nullptr, d_an, nullptr, d_bad, d_inf});
    launch(ctx->st, 1, ParsePointsTask{nullptr, d_bt, nullptr, d_at, d_bad + 1, nullptr});
    uint8_t hb[3];
    copy_d2h(ctx->st, hb, d_bad, 2);
    copy_d2h(ctx->st, hb + 2, d_inf, 1);
    sync(ctx->st);
    if (hb[0] || hb[1] || hb[2]) {
      delete P;
      return fail(ctx, ZKA_E_ARG, "params: h point not on its group");
    }
    build_p256_tab(ctx, d_an, P->h8, P->h_w);
    build_tom_tab(ctx, d_at, P->th);
    for (DevBuf* b : {&bn, &bt, &an, &at, &bad, &inf}) b->release();
    *out = P;
    return 0;
  } catch (const std::exception& e) {
    return fail(ctx, ZKA_E_CUDA, e.what());
  }
}

void zka_params_destroy(zka_params* P) {
  if (!P) return;
  P->h8.buf.release();
  P->th.buf.release();
  delete P;
}

size_t zka_proof_max_len(uint32_t ring_size, uint32_t sec_level) {
  return (size_t)proof_len((int)sec_level, ceil_log2(ring_size), (int)sec_level);
}
size_t zka_prove_tape_len(uint32_t ring_size, uint32_t sec_level) {
  return (size_t)32 * prove_draws((int)sec_level, ceil_log2(ring_size), (int)sec_level);
}
size_t zka_verify_tape_len(uint32_t ring_size, uint32_t sec_level) {
  return verify_tape_len(ceil_log2(ring_size), (int)sec_level);
}

// ------------------------------------------------------------------------------ sub-ops
int zka_tom_commit_batch(zka_ctx* ctx, const zka_params* P, uint32_t count, const uint8_t* v, const uint8_t* r,
                         uint8_t* out) {
  if (!ctx || !P || !v || !r || !out) return ZKA_E_ARG;
  if (count == 0) return 0;
  try {
    Stream& st = ctx->st;
    const uint8_t* dv = stage_in(ctx, ctx->in[0], v, (size_t)count * 32);
    const uint8_t* dr = stage_in(ctx, ctx->in[1], r, (size_t)count * 32);
    uint32_t* jv = ctx->w[0].get<uint32_t>((size_t)count * 8);
    uint32_t* jr = ctx->w[1].get<uint32_t>((size_t)count * 8);
    uint32_t* proj = ctx->w[2].get<uint32_t>((size_t)count * TOM_PROJ_WORDS);
    uint32_t* aff = ctx->w[3].get<uint32_t>((size_t)count * TOM_AFF_WORDS);
    uint8_t* bytes = ctx->w[4].get<uint8_t>((size_t)count * BSTRIDE);
    launch(st, count, CommitConvTask{dv, dr, jv, jr});
    launch(st, count, TomCommitTask{jv, jr, ctx->tg.tab, P->th.tab, proj, ctx->tom_w, ctx->tom_nwin});
    launch_tom_norm(st, proj, aff, bytes, (long long)(count), 1);
    if (is_device_ptr(out)) {
      launch(st, count, PackTomTask{bytes, out});
    } else {
      uint8_t* packed = ctx->out[0].get<uint8_t>((size_t)count * WP);
      launch(st, count, PackTomTask{bytes, packed});
      copy_d2h(st, out, packed, (size_t)count * WP);
    }
    sync(st);
    return 0;
  } catch (const std::exception& e) {
    return fail(ctx, ZKA_E_CUDA, e.what());
  }
}

int zka_p256_mul_batch(zka_ctx* ctx, uint32_t count, const uint8_t* base, const uint8_t* k, uint8_t* out) {
  if (!ctx || !k || !out) return ZKA_E_ARG;
  if (count == 0) return 0;
  try {
    Stream& st = ctx->st;
    const uint8_t* dk = stage_in(ctx, ctx->in[0], k, (size_t)count * 32);
    const uint8_t* db = base ? stage_in(ctx, ctx->in[1], base, (size_t)count * 65) : nullptr;
    uint32_t* proj = ctx->w[0].get<uint32_t>((size_t)count * P256_PROJ_WORDS);
    uint32_t* aff = ctx->w[1].get<uint32_t>((size_t)count * P256_AFF_WORDS);
    uint8_t* bytes = ctx->w[2].get<uint8_t>((size_t)count * BSTRIDE);
    uint8_t* inf = ctx->w[3].get<uint8_t>(count);
    uint32_t* rtab = nullptr;
    uint8_t* binf = nullptr;
    if (db) {
      // per-base w=4 positional tables: the same path the prover uses for R (PhaseAP256Task)
      uint32_t* baff = ctx->w[4].get<uint32_t>((size_t)count * 16);
      uint8_t* bad = ctx->w[5].get<uint8_t>(count);
      binf = ctx->w[6].get<uint8_t>(count);
      uint32_t* pows = ctx->w[7].get<uint32_t>((size_t)count * RT_NWIN * P256_PROJ_WORDS);
      uint32_t* rows = ctx->w[8].get<uint32_t>((size_t)count * RT_ENTRIES * P256_PROJ_WORDS);
      rtab = ctx->w[9].get<uint32_t>((size_t)count * RT_ENTRIES * P256_AFF_WORDS);
      launch(st, count, ParsePointsTask{db, nullptr, baff, nullptr, bad, binf});
      launch(st, count, P256PowsTask{baff, binf, pows, (int)count, RT_NWIN, RT_W});
      launch(st, (long long)count * RT_NWIN, P256RowsSignedTask{pows, rows});
      const long long np = (long long)count * RT_ENTRIES;
      launch_p256_norm(st, rows, rtab, nullptr, nullptr, (long long)(np));
    }
    launch(st, count, P256MulTask{dk, ctx->g8.tab, rtab, binf, proj});
    launch_p256_norm(st, proj, aff, bytes, inf, (long long)(count));
    if (is_device_ptr(out)) {
      launch(st, count, PackP256Task{bytes, out});
    } else {
      uint8_t* packed = ctx->out[0].get<uint8_t>((size_t)count * NP);
      launch(st, count, PackP256Task{bytes, packed});
      copy_d2h(st, out, packed, (size_t)count * NP);
    }
    sync(st);
    return 0;
  } catch (const std::exception& e) {
    return fail(ctx, ZKA_E_CUDA, e.what());
  }
}


int zka_field_op_batch(zka_ctx* ctx, int field, int op, uint32_t count, const uint8_t* a, const uint8_t* b,
                       uint8_t* out) {
  if (!ctx || !a || !out || field < 0 || field > 2 || op < 0 || op > 4 || (op < 3 && !b)) return ZKA_E_ARG;
  if (count == 0) return 0;
  try {
    Stream& st = ctx->st;
    const int nb = field == 2 ? WCB : 32;
    const uint8_t* da = stage_in(ctx, ctx->in[0], a, (size_t)count * nb);
    const uint8_t* db = b ? stage_in(ctx, ctx->in[1], b, (size_t)count * nb) : nullptr;
    uint8_t* dout = is_device_ptr(out) ? out : ctx->out[0].get<uint8_t>((size_t)count * nb);
    if (field == 0) launch(st, count, FieldOpTask<P256p, 32>{da, db, dout, op});
    else if (field == 1) launch(st, count, FieldOpTask<P256n, 32>{da, db, dout, op});
    else launch(st, count, FieldOpTask<PGp, WCB>{da, db, dout, op});
    if (dout != out) copy_d2h(st, out, dout, (size_t)count * nb);
    sync(st);
    return 0;
  } catch (const std::exception& e) {
    return fail(ctx, ZKA_E_CUDA, e.what());
  }
}

int zka_hash80_batch(zka_ctx* ctx, uint32_t count, const uint8_t* msgs, size_t msg_stride, const uint32_t* len,
                     uint8_t* out) {
  if (!ctx || !msgs || !len || !out) return ZKA_E_ARG;
  if (count == 0) return 0;
  try {
    Stream& st = ctx->st;
    const uint8_t* dm = stage_in(ctx, ctx->in[0], msgs, (size_t)count * msg_stride);
    const uint32_t* dl = stage_in(ctx, ctx->in[1], len, (size_t)count);
    uint8_t* dout = is_device_ptr(out) ? out : ctx->out[0].get<uint8_t>((size_t)count * 10);
    launch(st, count, Hash80Task{dm, msg_stride, dl, dout});
    if (dout != out) copy_d2h(st, out, dout, (size_t)count * 10);
    sync(st);
    return 0;
  } catch (const std::exception& e) {
    return fail(ctx, ZKA_E_CUDA, e.what());
  }
}

int zka_params_generate(zka_ctx* ctx, const uint8_t rnd[64], uint8_t h_nist[65], uint8_t h_proof[67]) {
  if (!ctx || !rnd || !h_nist || !h_proof) return ZKA_E_ARG;
  try {
    // h_nist = G * rnd0 (pedersen.ts:66-67 on p256); h_proof = g * rnd1 + 0 * g
    int rc = zka_p256_mul_batch(ctx, 1, nullptr, rnd, h_nist);
    if (rc) return rc;
    Stream& st = ctx->st;
    uint32_t* jv = ctx->w[0].get<uint32_t>(8);
    uint32_t* jr = ctx->w[1].get<uint32_t>(8);
    uint32_t* proj = ctx->w[2].get<uint32_t>(TOM_PROJ_WORDS);
    uint32_t* aff = ctx->w[3].get<uint32_t>(TOM_AFF_WORDS);
    uint8_t* bytes = ctx->w[4].get<uint8_t>(BSTRIDE);
    uint8_t* d_rnd = ctx->in[0].get<uint8_t>(32);
    copy_h2d(st, d_rnd, rnd + 32, 32);
    launch(st, 1, GenConvTask{d_rnd, jv, jr});
    // v*g + 0*g: use the g table for both bases
    launch(st, 1, TomCommitTask{jv, jr, ctx->tg.tab, ctx->tg.tab, proj, ctx->tom_w, ctx->tom_nwin});
    launch_tom_norm(st, proj, aff, bytes, (long long)(1), 1);
    copy_d2h(st, h_proof, bytes, WP);
    sync(st);
    return 0;
  } catch (const std::exception& e) {
    return fail(ctx, ZKA_E_CUDA, e.what());
  }
}

int zka_key_to_int(zka_ctx* ctx, uint32_t count, const uint8_t* pk, uint8_t* x_out, int32_t* status) {
  if (!ctx || !pk || !x_out) return ZKA_E_ARG;
  if (count == 0) return 0;
  try {
    Stream& st = ctx->st;
    const uint8_t* dp = stage_in(ctx, ctx->in[0], pk, (size_t)count * 65);
    uint32_t* aff = ctx->w[0].get<uint32_t>((size_t)count * 16);
    uint8_t* bad = ctx->w[1].get<uint8_t>(count);
    uint8_t* inf = ctx->w[2].get<uint8_t>(count);
    uint8_t* dx = ctx->out[0].get<uint8_t>((size_t)count * 32);
    int32_t* ds = ctx->out[1].get<int32_t>(count);
    launch(st, count, ParsePointsTask{dp, nullptr, aff, nullptr, bad, inf});
    launch(st, count, KeyToIntTask{aff, bad, inf, dx, ds});
    if (is_device_ptr(x_out)) copy_d2d(st, x_out, dx, (size_t)count * 32); else copy_d2h(st, x_out, dx, (size_t)count * 32);
    if (status) {
      if (is_device_ptr(status)) copy_d2d(st, status, ds, (size_t)count * 4); else copy_d2h(st, status, ds, (size_t)count * 4);
    }
    sync(st);
    return 0;
  } catch (const std::exception& e) {
    return fail(ctx, ZKA_E_CUDA, e.what());
  }
}

// ------------------------------------------------------------------------------- prove
// mode 0: proveSignatureList.  mode 1: proveExp alone (exp.ts:126-231) — base / s_in / q_in are the statement,
// msg_hash / sig / which / ring are unused, the rows hold the repetitions only.
static int prove_impl(zka_ctx* ctx, const zka_params* P, uint32_t B, const uint8_t* msg_hash, const uint8_t* sig,
                      const uint8_t* pk, const uint32_t* which, const uint8_t* ring, uint32_t N, const uint8_t* tape,
                      size_t tape_stride, uint8_t* proofs, size_t proof_stride, uint32_t* proof_len_out,
                      int32_t* status, int mode, const uint8_t* base, const uint8_t* s_in, const uint8_t* q_in) {
  if (!ctx || !P || !pk || !tape || !proofs || !proof_len_out || !status) return ZKA_E_ARG;
  if (mode == 0 && (!msg_hash || !sig || !which || !ring)) return ZKA_E_ARG;
  if (mode == 1 && (!base || !s_in)) return ZKA_E_ARG;
  if (B == 0) return 0;
  // N = 1 makes hashPoints([]) throw in the reference (group.ts:223 reduce of an empty array)
  if (mode == 0 && (N < 2 || N > (1u << 20))) return fail(ctx, ZKA_E_ARG, "ring size must be in [2, 2^20]");
  const int S = (int)P->sec_level;
  const int n = mode == 0 ? ceil_log2(N) : 0;
  if (proof_stride < (mode == 0 ? zka_proof_max_len(N, S) : (size_t)S * REP0_LEN)) return fail(ctx, ZKA_E_ARG, "proof_stride < zka_proof_max_len");
  if (tape_stride < (size_t)32 * (mode == 0 ? prove_draws(0, n, S) : draws_before_items(S))) return fail(ctx, ZKA_E_ARG, "tape_stride too small");
  try {
    // ring + Lagrange matrix: once per call, on lane 0, finished before the lanes start
    if (mode == 0) {
      Stream& st0 = ctx->st;
      const uint8_t* d_ring = stage_in(st0, ctx->ring_in, ring, (size_t)N * 32);
      uint32_t* rm = ctx->ring_m.get<uint32_t>(((size_t)1 << n) * 8);
      launch(st0, 1ll << n, RingPrepTask{d_ring, rm, (int)N});
      if (ctx->lag_n != n) {   // depends only on n, cached per context
        uint32_t* l = ctx->lag.get<uint32_t>((size_t)n * n * 8);
        launch(st0, 1, GkLagrangeTask{l, n});
        ctx->lag_n = n;
      }
      sync(st0);
    }
    const uint32_t* ring_m = (const uint32_t*)ctx->ring_m.p;
    uint32_t* lag = (uint32_t*)ctx->lag.p;
    const bool out_dev = is_device_ptr(proofs);
    const bool len_dev = is_device_ptr(proof_len_out), st_dev = is_device_ptr(status);
    const int lanes = ctx->nlanes;
    const bool all_dev = out_dev && is_device_ptr(tape);
    const std::vector<uint32_t> off = chunk_schedule(B, (uint32_t)(all_dev ? ctx->chunk : std::min(ctx->chunk, ctx->host_chunk)), lanes, !all_dev);
    const uint32_t nchunks = (uint32_t)off.size() - 1;
    const int used = (int)std::min<uint32_t>((uint32_t)lanes, nchunks);
    const bool trace = getenv("ZKA_TRACE") != nullptr;
    const auto t_call = std::chrono::steady_clock::now();
    auto ms_now = [&] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_call).count(); };
    std::atomic<uint32_t> next_chunk((uint32_t)used);
    // Every lane claims chunks from a shared counter (one ahead of the one it is computing, so that its inputs are
    // already on their way).  Within a lane, chunks are software-pipelined over three streams when buffers live in
    // host memory (staging buffers double-buffered by slot).
    auto run_lane = [&](int li) {
      Lane& ln = ctx->lane(li);
      Stream& st = ln.st;
      DevBuf* W = ln.w;
      struct ChunkIn { const uint8_t *msg_hash, *sig, *pk, *tape, *base, *s_in, *q_in; const uint32_t* which; } cin[2];
      auto issue_inputs = [&](uint32_t k, int slot) {
        const uint32_t b0 = off[k];
        const size_t Bc = off[k + 1] - b0;
        Stream& ci = ln.cs_in;
        DevBuf* in = ln.in + 5 * slot;
        ev_wait(ci, ln.ev_done[slot]);   // the chunk that used these staging buffers before has finished reading them
        cin[slot].msg_hash = msg_hash ? stage_in(ci, in[0], msg_hash + (size_t)b0 * 32, Bc * 32) : nullptr;
        cin[slot].sig = sig ? stage_in(ci, in[1], sig + (size_t)b0 * 64, Bc * 64) : nullptr;
        cin[slot].pk = stage_in(ci, in[2], pk + (size_t)b0 * 65, Bc * 65);
        cin[slot].which = which ? stage_in(ci, in[3], which + b0, Bc) : nullptr;
        DevBuf* inx = ln.in + 10 + 3 * slot;
        cin[slot].base = base ? stage_in(ci, inx[0], base + (size_t)b0 * 65, Bc * 65) : nullptr;
        cin[slot].s_in = s_in ? stage_in(ci, inx[1], s_in + (size_t)b0 * 32, Bc * 32) : nullptr;
        cin[slot].q_in = q_in ? stage_in(ci, inx[2], q_in + (size_t)b0 * 65, Bc * 65) : nullptr;
        ev_record(ln.ev_small[slot], ci);
        if (is_device_ptr(tape)) {
          cin[slot].tape = tape + (size_t)b0 * tape_stride;
        } else if (!ctx->tape_split) {
          cin[slot].tape = stage_in(ci, in[4], tape + (size_t)b0 * tape_stride, Bc * tape_stride);
        } else {
          // host tape: only the draws used before the challenge (3 + 4S of up to 3 + 44S + 5n) travel now; the
          // item and GK draws of each proof follow after the challenge, when their number is known
          uint8_t* dt = in[4].get<uint8_t>(Bc * tape_stride);
          const size_t pre = std::min(tape_stride, (size_t)32 * draws_before_items(S));
          copy_d2h_2d(ci, dt, tape_stride, tape + (size_t)b0 * tape_stride, tape_stride, pre, Bc);
          cin[slot].tape = dt;
        }
        ev_record(ln.ev_tape[slot], ci);
      };
      // the first `used` chunks are dealt statically (lane threads start at slightly different times); later ones are
      // claimed from the shared counter at the mid-pipeline synchronisation point of the current chunk, when about
      // half of its kernels are queued: early enough for the next inputs to travel behind them, late enough that a
      // lane that started first does not grab the chunks of lanes that are still starting
      uint32_t k = (uint32_t)li;
      int slot = 0;
      if (k < nchunks) issue_inputs(k, slot);
      for (; k < nchunks; slot ^= 1) {
        const uint32_t b0 = off[k];
        const int Bc = (int)(off[k + 1] - b0);
        const uint32_t k_this = k;
        const double t_begin = ms_now();
        ev_wait(st, ln.ev_small[slot]);
        ev_wait(st, ln.ev_out[slot]);    // the proofs of chunk k-2 have left the output staging buffers
        ProveCtx c;
        memset(&c, 0, sizeof(c));
        c.B = Bc; c.S = S; c.N = (int)N; c.n = n; c.M = 0;
        c.mode = mode; c.head_len = mode == 0 ? HEAD_LEN : 0;
        c.base = cin[slot].base; c.s_in = cin[slot].s_in; c.q_in = cin[slot].q_in;
        c.tom_w = ctx->tom_w; c.tom_nwin = ctx->tom_nwin;
        c.msg_hash = cin[slot].msg_hash;
        c.sig = cin[slot].sig;
        c.pk = cin[slot].pk;
        c.which = cin[slot].which;
        c.tape = cin[slot].tape;
        c.tape_stride = tape_stride;
        c.tape_draws = (uint32_t)(tape_stride / 32);
        c.ring_m = ring_m;
        c.g_tab8 = ctx->g8.tab; c.h_tab8 = P->h8.tab; c.h_w = P->h_w;
        c.g_tabw = ctx->gw.tab; c.g_w = ctx->p256_hw;
        c.tg_tab = ctx->tg.tab; c.th_tab = P->th.tab;
        c.tg_bytes = (const uint8_t*)ctx->tg_bytes.p;
        const size_t S1 = (size_t)S + 1;
        const size_t nA = (size_t)Bc * S1;
        const size_t n1 = (size_t)Bc * (2 + 2 * S);
        c.s1 = W[0].get<uint32_t>((size_t)Bc * 8);
        c.pk_aff = W[1].get<uint32_t>((size_t)Bc * 16);
        c.q_aff = W[2].get<uint32_t>((size_t)Bc * 16);
        c.q_inf = W[3].get<uint8_t>(Bc);
        c.r_aff = W[4].get<uint32_t>((size_t)Bc * 16);
        c.r_bytes = W[5].get<uint8_t>((size_t)Bc * BSTRIDE);
        c.rpows = W[6].get<uint32_t>((size_t)Bc * RT_NWIN * P256_PROJ_WORDS);
        c.rrows = W[7].get<uint32_t>((size_t)Bc * KEY_CAP * P256_PROJ_WORDS);
        c.rtab = W[8].get<uint32_t>((size_t)Bc * KEY_CAP * P256_AFF_WORDS);
        c.pa_T = W[9].get<uint32_t>(nA * P256_PROJ_WORDS);
        c.pa_A = W[10].get<uint32_t>(nA * P256_PROJ_WORDS);
        c.pa_T_aff = W[11].get<uint32_t>(nA * 16);
        c.pa_T_inf = W[12].get<uint8_t>(nA);
        c.pa_A_aff = W[13].get<uint32_t>(nA * 16);
        c.pa_A_bytes = W[14].get<uint8_t>(nA * BSTRIDE);
        c.pa_A_inf = W[15].get<uint8_t>(nA);
        c.s1_jv = W[16].get<uint32_t>(n1 * 8);
        c.s1_jr = W[17].get<uint32_t>(n1 * 8);
        c.s1_proj = W[18].get<uint32_t>(n1 * TOM_PROJ_WORDS);
        c.s1_aff = W[19].get<uint32_t>(n1 * TOM_AFF_WORDS);
        c.s1_bytes = W[20].get<uint8_t>(n1 * BSTRIDE);
        c.chal = W[21].get<uint32_t>((size_t)Bc * 3);
        c.zcount = W[22].get<uint32_t>(Bc);
        c.item_base = W[23].get<uint32_t>(Bc);
        c.item_total = W[24].get<uint32_t>(2);
        c.rep_off = W[25].get<uint32_t>((size_t)Bc * S);
        c.gk_off = W[26].get<uint32_t>(Bc);
        c.gk_dv = W[27].get<uint32_t>((size_t)Bc * n * 8);
        c.gk_lag = lag;
        c.gk_x = W[28].get<uint32_t>((size_t)Bc * 3);
        c.u12 = W[46].get<uint32_t>((size_t)Bc * 16);
        c.tab_of = W[48].get<uint32_t>(Bc);
        c.tab_rep = W[49].get<uint32_t>((size_t)Bc * 2);
        c.tab_count = W[50].get<uint32_t>(2);
        c.which_s = W[52].get<uint32_t>(Bc);
        c.base_aff = mode == 0 ? c.pk_aff : W[53].get<uint32_t>((size_t)Bc * 16);
        c.proof_stride = proof_stride;
        DevBuf* ob = ln.out + 3 * slot;
        c.proofs = out_dev ? proofs + (size_t)b0 * proof_stride : ob[0].get<uint8_t>((size_t)Bc * proof_stride);
        c.proof_len = len_dev ? proof_len_out + b0 : ob[1].get<uint32_t>(Bc);
        c.status = st_dev ? status + b0 : ob[2].get<int32_t>(Bc);
  
        // --- statement + per-proof tables of pk, then R = u1*G + u2*pk on the tables
        launch(st, Bc, PreKeyTask{c});
        // one table per DISTINCT key of the chunk (grids are sized for Bc tables, surplus threads return)
        launch(st, Bc, KeyDedupTask{c});
        launch(st, Bc, KeyRankTask{c});
        launch(st, Bc, KeyAssignTask{c});
        {
          const int Bp = (Bc + 31) & ~31;
          launch(st, (long long)Bp + Bc,
                 PowsAndPreTask{P256PowsTask{c.base_aff, nullptr, c.rpows, Bc, RT_NWIN, RT_W, c.tab_rep, c.tab_count, c.tab_count + 1}, PreTask{c}, Bp});
        }
        // the window bits of these tables are chosen on the device from the number of distinct keys (tab_count[1]);
        // grids are sized for the worst case, surplus threads return
        // (one thread per (key, window, block of 16 entries): keys x windows x blocks <= Bc x KEY_CAP / 16 by the memory
        // rule of key_window_bits, e.g. 0.2 Bc keys x 33 x 8 at w = 8 or Bc x 52 x 1 at w = 5)
        launch(st, (long long)Bc * ((KEY_CAP + 15) / 16), P256RowsBlockTask{c.rpows, c.rrows, KEY_W_MIN, c.tab_count, c.tab_count + 1});
        {
          const long long np = (long long)Bc * KEY_CAP;
          // points per thread from the EXPECTED table volume (at most min(N, Bc) distinct keys when every key is a ring
          // member); the grid still covers the worst case
          const uint32_t kest = std::min<uint32_t>(N, (uint32_t)Bc);
          const int west = key_window_bits(kest, (uint32_t)Bc, (uint32_t)S + 2);
          const int ch = norm_chunk_for((long long)kest * fb_windows(west) * fb_entries(west), 5);
          launch(st, (np + ch - 1) / ch, P256NormTask{c.rrows, c.rtab, nullptr, nullptr, (int)np, ch, c.tab_count, 0, c.tab_count + 1});
        }
        // --- phase A (first consumer of the tape) and R = u1*G + u2*pk side by side
        ev_wait(st, ln.ev_tape[slot]);
        {
          const int nAp = (int)((nA + 31) & ~(size_t)31);
          launch(st, (long long)nAp + Bc, PhaseAAndRPointTask{PhaseAP256Task{c}, RPointTask{c}, (int)nA, nAp});
        }
        launch_p256_norm(st, c.pa_T, c.pa_T_aff, nullptr, c.pa_T_inf, (long long)(nA));
        launch_p256_norm(st, c.pa_A, c.pa_A_aff, c.pa_A_bytes, c.pa_A_inf, (long long)(nA));
        if (mode == 1) launch(st, Bc, ExpStatementTask{c});
        launch(st, (long long)n1, JobsATask{c});
        launch(st, (long long)n1, TomCommitTask{c.s1_jv, c.s1_jr, c.tg_tab, c.th_tab, c.s1_proj, c.tom_w, c.tom_nwin});
        launch_tom_norm(st, c.s1_proj, c.s1_aff, c.s1_bytes, (long long)(n1), 1);
        // --- challenge, layout
        launch(st, Bc, ExpChallengeTask{c});
        launch(st, 1, ScanTask{c});
        uint32_t tot2[2] = {0, 0};
        copy_d2h(st, tot2, c.item_total, 8);
        const bool tape_host = ctx->tape_split && !is_device_ptr(tape);
        sync(st);
        if (tape_host) {
          // second part of the tape: draws [3 + 4S, 3 + 4S + 40 zmax + 5n) of every row in one strided copy
          // (zmax = the largest zero-bit count of the chunk)
          const size_t o0 = (size_t)32 * draws_before_items(S);
          const size_t o1 = std::min(tape_stride, (size_t)32 * prove_draws((int)tot2[1], n, S));
          if (o1 > o0)
            copy_d2h_2d(st, const_cast<uint8_t*>(c.tape) + o0, tape_stride, tape + (size_t)b0 * tape_stride + o0, tape_stride, o1 - o0, Bc);
        }
        const double t_mid = ms_now();
        {
          const uint32_t kn = next_chunk.fetch_add(1);
          if (kn < nchunks) issue_inputs(kn, slot ^ 1);
          k = kn;
        }
        const uint32_t M = tot2[0];
        const size_t max_len = mode == 0 ? (size_t)proof_len((int)tot2[1], n, S)
                                         : (size_t)tot2[1] * REP0_LEN + (size_t)(S - (int)tot2[1]) * REP1_LEN;
        c.M = (int)M;
        c.item_b = W[29].get<uint32_t>(M);
        c.item_i = W[30].get<uint32_t>(M);
        c.item_k = W[31].get<uint32_t>(M);
        c.pb_T1 = W[32].get<uint32_t>((size_t)M * P256_PROJ_WORDS);
        c.pb_T1_aff = W[33].get<uint32_t>((size_t)M * 16);
        c.pb_T1_inf = W[34].get<uint8_t>(M);
        const size_t n2 = c.s2_count();
        c.s2_jv = W[35].get<uint32_t>(n2 * 8);
        c.s2_jr = W[36].get<uint32_t>(n2 * 8);
        c.s2_proj = W[37].get<uint32_t>(n2 * TOM_PROJ_WORDS);
        c.s2_aff = W[38].get<uint32_t>(n2 * TOM_AFF_WORDS);
        c.s2_bytes = W[39].get<uint8_t>(n2 * BSTRIDE);
        c.secrets = W[42].get<uint32_t>((size_t)M * SECRETS_PER_ITEM * 8);
        c.item_inv = W[47].get<uint32_t>((size_t)M * 8);
        c.item_chal = W[43].get<uint32_t>((size_t)M * HASHES_PER_ITEM * 3);
        launch(st, Bc, ItemsTask{c});
        // --- phase B
        launch(st, M, PhaseBP256Task{c});
        launch_p256_norm(st, c.pb_T1, c.pb_T1_aff, nullptr, c.pb_T1_inf, (long long)(M));
        launch(st, ((long long)M + ITEM_INV_CHUNK - 1) / ITEM_INV_CHUNK, ItemInvTask{c});
        launch(st, M, ItemScalarsTask{c});
        launch(st, (long long)Bc * n, GkJobsTask{c});
        {
          const int nblk = 1 << (n - gk_block_bits(n));
          c.gk_part = nblk > 1 ? W[51].get<uint32_t>((size_t)Bc * n * nblk * 8) : nullptr;
          launch(st, (long long)Bc * n * nblk, GkPolyTask{c});
          if (nblk > 1) launch(st, (long long)Bc * n, GkPolyReduceTask{c});
        }
        launch(st, (long long)Bc * n, GkCdJobsTask{c});
        {
          const size_t nj = (size_t)M * JOBS_PER_ITEM, nd = (size_t)M * DERS_PER_ITEM, ng = (size_t)Bc * 4 * n;
          const size_t g0 = nj + nd;
          // item jobs: g-parts once per distinct committed value, then r*h on top (TomCommitG/HTask)
          uint32_t* gext = W[45].get<uint32_t>((size_t)M * GJOBS_PER_ITEM * TOM_EXT_WORDS);
          launch(st, (long long)M * GJOBS_PER_ITEM, TomCommitGTask{c.s2_jv, c.tg_tab, gext, c.tom_w, c.tom_nwin});
          launch(st, (long long)nj, TomCommitHTask{c.s2_jr, c.th_tab, gext, c.s2_proj, c.tom_w, c.tom_nwin});
          launch(st, (long long)ng, TomCommitTask{c.s2_jv + g0 * 8, c.s2_jr + g0 * 8, c.tg_tab, c.th_tab,
                                                   c.s2_proj + g0 * TOM_PROJ_WORDS, c.tom_w, c.tom_nwin});
          // only T1x, T1y (jobs 0, 1 of each item) are needed again as points (DerivedTask)
          launch_tom_norm(st, c.s2_proj, c.s2_aff, c.s2_bytes, (long long)(nj), 1, JOBS_PER_ITEM, 2);
          launch(st, M, DerivedTask{c});
          // derived points come from complete E1 additions, the GK commitments from the commit kernel (E2)
          launch_tom_norm(st, c.s2_proj + nj * TOM_PROJ_WORDS, nullptr, c.s2_bytes + nj * BSTRIDE, (long long)nd, 0);
          launch_tom_norm(st, c.s2_proj + g0 * TOM_PROJ_WORDS, nullptr, c.s2_bytes + g0 * BSTRIDE, (long long)ng, 1);
        }
        launch(st, (long long)M * HASHES_PER_ITEM, ItemHashTask{c});
        launch(st, (long long)M * 7, ItemEmitTask{c});
        launch(st, (long long)nA, RepEmitTask{c});
        if (mode == 0) launch(st, Bc, GkEmitTask{c});
        launch(st, (long long)Bc * FIN_PARTS, FinalizeTask{c});
        // --- results: on the output stream, behind this chunk's last kernel
        ev_record(ln.ev_done[slot], st);
        if (ctx->progress && k_this < ctx->progress_cap) notify_progress(st, ctx->progress + k_this);
        if (!out_dev || !len_dev || !st_dev) {
          Stream& co = ln.cs_out;
          ev_wait(co, ln.ev_done[slot]);
          // only the bytes up to the longest proof of the chunk are copied back (rows are stride-padded)
          // (one cudaMemcpyAsync per row with its exact length was measured: 8192 driver calls per step cost more than
          // the ~25 % of padding they save — config2 e2e 89.9k -> 38.5k proofs/s, gpurun_out/bench_c2_r2g.json)
          if (!out_dev) copy_d2h_2d(co, proofs + (size_t)b0 * proof_stride, proof_stride, c.proofs, proof_stride, max_len, Bc);
          if (!len_dev) copy_d2h(co, proof_len_out + b0, c.proof_len, (size_t)Bc * 4);
          if (!st_dev) copy_d2h(co, status + b0, c.status, (size_t)Bc * 4);
          ev_record(ln.ev_out[slot], co);
        }
        if (trace) {
          const double t_enq = ms_now();
          sync(st);
          const double t_comp = ms_now();
          sync(ln.cs_out);
          fprintf(stderr, "TRACE lane %d chunk %u rows %d begin %.2f mid %.2f enqueued %.2f computed %.2f copied %.2f\n", li, k_this, Bc,
                  t_begin, t_mid, t_enq, t_comp, ms_now());
        }
      }
      sync(ln.cs_in);
      sync(ln.cs_out);
      sync(st);
    };
    run_lanes(ctx, used, run_lane);
    return 0;
  } catch (const std::exception& e) {
    return fail(ctx, ZKA_E_CUDA, e.what());
  }
}

int zka_prove_batch(zka_ctx* ctx, const zka_params* P, uint32_t B, const uint8_t* msg_hash, const uint8_t* sig,
                    const uint8_t* pk, const uint32_t* which, const uint8_t* ring, uint32_t N, const uint8_t* tape,
                    size_t tape_stride, uint8_t* proofs, size_t proof_stride, uint32_t* proof_len_out,
                    int32_t* status) {
  return prove_impl(ctx, P, B, msg_hash, sig, pk, which, ring, N, tape, tape_stride, proofs, proof_stride, proof_len_out, status, 0,
                    nullptr, nullptr, nullptr);
}

// proveExp(paramsNIST = (p256, base, NistGroup.h), paramsWario = ProofGroup, s, Cs, P = pk, Px, Py, secparam = sec_level, Q?)
// (exp.ts:126-231).  Tape layout = zka_prove_batch's: draws 0..2 are the blinders of Cs, Px, Py (drawn when those
// commitments were made), then 4 per repetition, then 40 per 0-bit repetition.  Rows: the repetitions only.
int zka_prove_exp_batch(zka_ctx* ctx, const zka_params* P, uint32_t B, const uint8_t* base, const uint8_t* s, const uint8_t* pk,
                        const uint8_t* q, const uint8_t* tape, size_t tape_stride, uint8_t* proofs, size_t proof_stride,
                        uint32_t* proof_len, int32_t* status) {
  return prove_impl(ctx, P, B, nullptr, nullptr, pk, nullptr, nullptr, 2, tape, tape_stride, proofs, proof_stride, proof_len, status, 1,
                    base, s, q);
}

// proveMembership(params = ProofGroup, com, index, ring) (gk.ts:94-195) for B commitments over one ring.
// com_r: the blinder of com = commit(ring[index]) (B x 32); tape: the 5n draws (r_i, a_i, s_i, t_i, rho_i per round).
// Rows: the GK block of the flat layout.
int zka_prove_membership_batch(zka_ctx* ctx, const zka_params* P, uint32_t B, const uint8_t* com_r, const uint32_t* index,
                               const uint8_t* ring, uint32_t N, const uint8_t* tape, size_t tape_stride, uint8_t* proofs,
                               size_t proof_stride, uint32_t* proof_len, int32_t* status) {
  if (!ctx || !P || !com_r || !index || !ring || !tape || !proofs || !proof_len || !status) return ZKA_E_ARG;
  if (B == 0) return 0;
  if (N < 2 || N > (1u << 20)) return fail(ctx, ZKA_E_ARG, "ring size must be in [2, 2^20]");
  const int n = ceil_log2(N);
  if (proof_stride < (size_t)gk_len(n)) return fail(ctx, ZKA_E_ARG, "proof_stride < GK block length");
  if (tape_stride < (size_t)32 * 5 * n) return fail(ctx, ZKA_E_ARG, "tape_stride < 32 * 5n");
  try {
    Stream& st = ctx->st;
    DevBuf* W = ctx->w;
    const uint8_t* d_ring = stage_in(st, ctx->ring_in, ring, (size_t)N * 32);
    uint32_t* ring_m = ctx->ring_m.get<uint32_t>(((size_t)1 << n) * 8);
    launch(st, 1ll << n, RingPrepTask{d_ring, ring_m, (int)N});
    if (ctx->lag_n != n) {
      uint32_t* l = ctx->lag.get<uint32_t>((size_t)n * n * 8);
      launch(st, 1, GkLagrangeTask{l, n});
      ctx->lag_n = n;
    }
    const size_t it_stride = 96 + tape_stride;   // internal tape: [pad, com.r, pad] then the caller's draws (S = 0: GK draws start at 3)
    const uint32_t chunk = 8192;
    for (uint32_t b0 = 0; b0 < B; b0 += chunk) {
      const int Bc = (int)std::min<uint32_t>(chunk, B - b0);
      const uint8_t* d_cr = stage_in(st, ctx->in[0], com_r + (size_t)b0 * 32, (size_t)Bc * 32);
      const uint32_t* d_idx = stage_in(st, ctx->in[1], index + b0, (size_t)Bc);
      const uint8_t* d_tape = stage_in(st, ctx->in[2], tape + (size_t)b0 * tape_stride, (size_t)Bc * tape_stride);
      ProveCtx c;
      memset(&c, 0, sizeof(c));
      c.B = Bc; c.S = 0; c.N = (int)N; c.n = n; c.M = 0;
      c.tom_w = ctx->tom_w; c.tom_nwin = ctx->tom_nwin;
      c.which = d_idx;
      c.ring_m = ring_m;
      c.tg_tab = ctx->tg.tab; c.th_tab = P->th.tab;
      uint8_t* itape = W[0].get<uint8_t>((size_t)Bc * it_stride);
      c.tape = itape; c.tape_stride = it_stride; c.tape_draws = (uint32_t)(it_stride / 32);
      c.which_s = W[1].get<uint32_t>(Bc);
      c.zcount = W[2].get<uint32_t>(Bc);
      c.gk_off = W[3].get<uint32_t>(Bc);
      c.gk_dv = W[4].get<uint32_t>((size_t)Bc * n * 8);
      c.gk_lag = (uint32_t*)ctx->lag.p;
      c.gk_x = W[5].get<uint32_t>((size_t)Bc * 3);
      const size_t ng = (size_t)Bc * 4 * n;
      c.s2_jv = W[6].get<uint32_t>(ng * 8);
      c.s2_jr = W[7].get<uint32_t>(ng * 8);
      c.s2_proj = W[8].get<uint32_t>(ng * TOM_PROJ_WORDS);
      c.s2_bytes = W[9].get<uint8_t>(ng * BSTRIDE);
      c.proof_stride = proof_stride;
      c.proofs = is_device_ptr(proofs) ? proofs + (size_t)b0 * proof_stride : ctx->out[0].get<uint8_t>((size_t)Bc * proof_stride);
      c.proof_len = is_device_ptr(proof_len) ? proof_len + b0 : ctx->out[1].get<uint32_t>(Bc);
      c.status = is_device_ptr(status) ? status + b0 : ctx->out[2].get<int32_t>(Bc);
      launch(st, Bc, GkAloneSetupTask{c, d_cr, d_tape, tape_stride, itape});
      launch(st, (long long)Bc * n, GkJobsTask{c});
      {
        const int nblk = 1 << (n - gk_block_bits(n));
        c.gk_part = nblk > 1 ? W[10].get<uint32_t>((size_t)Bc * n * nblk * 8) : nullptr;
        launch(st, (long long)Bc * n * nblk, GkPolyTask{c});
        if (nblk > 1) launch(st, (long long)Bc * n, GkPolyReduceTask{c});
      }
      launch(st, (long long)Bc * n, GkCdJobsTask{c});
      launch(st, (long long)ng, TomCommitTask{c.s2_jv, c.s2_jr, c.tg_tab, c.th_tab, c.s2_proj, c.tom_w, c.tom_nwin});
      launch_tom_norm(st, c.s2_proj, nullptr, c.s2_bytes, (long long)ng, 1);
      launch(st, Bc, GkEmitTask{c});
      launch(st, (long long)Bc * FIN_PARTS, FinalizeTask{c});
      if (!is_device_ptr(proofs)) copy_d2h_2d(st, proofs + (size_t)b0 * proof_stride, proof_stride, c.proofs, proof_stride, (size_t)gk_len(n), Bc);
      if (!is_device_ptr(proof_len)) copy_d2h(st, proof_len + b0, c.proof_len, (size_t)Bc * 4);
      if (!is_device_ptr(status)) copy_d2h(st, status + b0, c.status, (size_t)Bc * 4);
      sync(st);
    }
    return 0;
  } catch (const std::exception& e) {
    return fail(ctx, ZKA_E_CUDA, e.what());
  }
}

// proveEquality / proveMult alone (kind 0 / 1): see SubProveJobsTask
static int prove_sub(zka_ctx* ctx, const zka_params* P, int kind, uint32_t B, const uint8_t* scalars, const uint8_t* tape,
                     size_t tape_stride, uint8_t* commitments, uint8_t* proofs, int32_t* status) {
  if (!ctx || !P || !scalars || !tape || !commitments || !proofs || !status) return ZKA_E_ARG;
  if (B == 0) return 0;
  const int ns = kind == 0 ? 3 : 6, nd = kind == 0 ? 3 : 7, J = kind == 0 ? SUBP_EQ_JOBS : SUBP_MULT_JOBS;
  const int nc = kind == 0 ? 2 : 3, plen = kind == 0 ? EQ_LEN : MULT_LEN;
  if (tape_stride < (size_t)32 * nd) return fail(ctx, ZKA_E_ARG, "tape_stride too small for this sub-proof");
  try {
    Stream& st = ctx->st;
    DevBuf* W = ctx->w;
    const uint32_t chunk = 16384;
    for (uint32_t b0 = 0; b0 < B; b0 += chunk) {
      const int Bc = (int)std::min<uint32_t>(chunk, B - b0);
      const uint8_t* d_sc = stage_in(st, ctx->in[0], scalars + (size_t)b0 * ns * 32, (size_t)Bc * ns * 32);
      const uint8_t* d_tape = stage_in(st, ctx->in[1], tape + (size_t)b0 * tape_stride, (size_t)Bc * tape_stride);
      const size_t nj = (size_t)Bc * J;
      uint32_t* jv = W[0].get<uint32_t>(nj * 8);
      uint32_t* jr = W[1].get<uint32_t>(nj * 8);
      uint32_t* proj = W[2].get<uint32_t>(nj * TOM_PROJ_WORDS);
      uint8_t* bytes = W[3].get<uint8_t>(nj * BSTRIDE);
      const bool cd = is_device_ptr(commitments), pd = is_device_ptr(proofs), sd = is_device_ptr(status);
      uint8_t* d_com = cd ? commitments + (size_t)b0 * nc * WP : ctx->out[0].get<uint8_t>((size_t)Bc * nc * WP);
      uint8_t* d_prf = pd ? proofs + (size_t)b0 * plen : ctx->out[1].get<uint8_t>((size_t)Bc * plen);
      int32_t* d_st = sd ? status + b0 : ctx->out[2].get<int32_t>(Bc);
      launch(st, Bc, SubProveJobsTask{kind, d_sc, d_tape, tape_stride, jv, jr, d_st});
      launch(st, (long long)nj, TomCommitTask{jv, jr, ctx->tg.tab, P->th.tab, proj, ctx->tom_w, ctx->tom_nwin});
      launch_tom_norm(st, proj, nullptr, bytes, (long long)nj, 1);
      launch(st, Bc, SubProveEmitTask{kind, d_sc, d_tape, tape_stride, jr, bytes, d_com, d_prf, d_st});
      if (!cd) copy_d2h(st, commitments + (size_t)b0 * nc * WP, d_com, (size_t)Bc * nc * WP);
      if (!pd) copy_d2h(st, proofs + (size_t)b0 * plen, d_prf, (size_t)Bc * plen);
      if (!sd) copy_d2h(st, status + b0, d_st, (size_t)Bc * 4);
      sync(st);
    }
    return 0;
  } catch (const std::exception& e) {
    return fail(ctx, ZKA_E_CUDA, e.what());
  }
}
int zka_prove_equality_batch(zka_ctx* ctx, const zka_params* P, uint32_t B, const uint8_t* scalars, const uint8_t* tape,
                             size_t tape_stride, uint8_t* commitments, uint8_t* proofs, int32_t* status) {
  return prove_sub(ctx, P, 0, B, scalars, tape, tape_stride, commitments, proofs, status);
}
int zka_prove_mult_batch(zka_ctx* ctx, const zka_params* P, uint32_t B, const uint8_t* scalars, const uint8_t* tape,
                         size_t tape_stride, uint8_t* commitments, uint8_t* proofs, int32_t* status) {
  return prove_sub(ctx, P, 1, B, scalars, tape, tape_stride, commitments, proofs, status);
}

// provePointAdd alone (pointAdd.ts:92-163): the item stages of the batched prover with S = 1 (see PaddSetupTask)
int zka_prove_pointadd_batch(zka_ctx* ctx, const zka_params* P, uint32_t B, const uint8_t* points, const uint8_t* blinders,
                             const uint8_t* tape, size_t tape_stride, uint8_t* commitments, uint8_t* proofs, int32_t* status) {
  if (!ctx || !P || !points || !blinders || !tape || !commitments || !proofs || !status) return ZKA_E_ARG;
  if (B == 0) return 0;
  if (tape_stride < (size_t)32 * 38) return fail(ctx, ZKA_E_ARG, "tape_stride < 32 * 38");
  try {
    Stream& st = ctx->st;
    DevBuf* W = ctx->w;
    const size_t it_stride = (size_t)32 * (9 + 38);
    const size_t row_stride = REP0_LEN;
    const uint32_t chunk = 8192;
    for (uint32_t b0 = 0; b0 < B; b0 += chunk) {
      const int Bc = (int)std::min<uint32_t>(chunk, B - b0);
      const uint8_t* d_pts = stage_in(st, ctx->in[0], points + (size_t)b0 * 195, (size_t)Bc * 195);
      const uint8_t* d_bl = stage_in(st, ctx->in[1], blinders + (size_t)b0 * 192, (size_t)Bc * 192);
      const uint8_t* d_tape = stage_in(st, ctx->in[2], tape + (size_t)b0 * tape_stride, (size_t)Bc * tape_stride);
      ProveCtx c;
      memset(&c, 0, sizeof(c));
      c.B = Bc; c.S = 1; c.N = 2; c.n = 0; c.M = Bc; c.mode = 1; c.head_len = 0;
      c.tom_w = ctx->tom_w; c.tom_nwin = ctx->tom_nwin;
      c.tg_tab = ctx->tg.tab; c.th_tab = P->th.tab;
      c.tg_bytes = (const uint8_t*)ctx->tg_bytes.p;
      uint8_t* itape = W[0].get<uint8_t>((size_t)Bc * it_stride);
      c.tape = itape; c.tape_stride = it_stride; c.tape_draws = (uint32_t)(it_stride / 32);
      const size_t n1 = (size_t)Bc * 4, n2 = (size_t)Bc * (JOBS_PER_ITEM + DERS_PER_ITEM);
      c.s1 = W[1].get<uint32_t>((size_t)Bc * 8);
      c.pk_aff = W[2].get<uint32_t>((size_t)Bc * 16);
      c.pa_T_aff = W[3].get<uint32_t>((size_t)Bc * 2 * 16);
      c.pa_T_inf = W[4].get<uint8_t>((size_t)Bc * 2);
      c.pa_A_inf = W[5].get<uint8_t>((size_t)Bc * 2);
      c.pb_T1_aff = W[6].get<uint32_t>((size_t)Bc * 16);
      c.pb_T1_inf = W[7].get<uint8_t>(Bc);
      c.chal = W[8].get<uint32_t>((size_t)Bc * 3);
      c.zcount = W[9].get<uint32_t>(Bc);
      c.item_base = W[10].get<uint32_t>(Bc);
      c.item_b = W[11].get<uint32_t>(Bc);
      c.item_i = W[12].get<uint32_t>(Bc);
      c.item_k = W[13].get<uint32_t>(Bc);
      c.rep_off = W[14].get<uint32_t>(Bc);
      c.s1_jv = W[15].get<uint32_t>(n1 * 8);
      c.s1_jr = W[16].get<uint32_t>(n1 * 8);
      c.s1_proj = W[17].get<uint32_t>(n1 * TOM_PROJ_WORDS);
      c.s1_aff = W[18].get<uint32_t>(n1 * TOM_AFF_WORDS);
      c.s1_bytes = W[19].get<uint8_t>(n1 * BSTRIDE);
      c.s2_jv = W[20].get<uint32_t>(n2 * 8);
      c.s2_jr = W[21].get<uint32_t>(n2 * 8);
      c.s2_proj = W[22].get<uint32_t>(n2 * TOM_PROJ_WORDS);
      c.s2_aff = W[23].get<uint32_t>(n2 * TOM_AFF_WORDS);
      c.s2_bytes = W[24].get<uint8_t>(n2 * BSTRIDE);
      c.secrets = W[25].get<uint32_t>((size_t)Bc * SECRETS_PER_ITEM * 8);
      c.item_inv = W[26].get<uint32_t>((size_t)Bc * 8);
      c.item_chal = W[27].get<uint32_t>((size_t)Bc * HASHES_PER_ITEM * 3);
      uint32_t* gext = W[28].get<uint32_t>((size_t)Bc * GJOBS_PER_ITEM * TOM_EXT_WORDS);
      c.proof_stride = row_stride;
      c.proofs = W[29].get<uint8_t>((size_t)Bc * row_stride);
      c.proof_len = W[30].get<uint32_t>(Bc);
      const bool cd = is_device_ptr(commitments), pd = is_device_ptr(proofs), sd = is_device_ptr(status);
      c.status = sd ? status + b0 : ctx->out[2].get<int32_t>(Bc);
      uint8_t* d_com = cd ? commitments + (size_t)b0 * 6 * WP : ctx->out[0].get<uint8_t>((size_t)Bc * 6 * WP);
      uint8_t* d_prf = pd ? proofs + (size_t)b0 * PA_LEN : ctx->out[1].get<uint8_t>((size_t)Bc * PA_LEN);
      launch(st, Bc, PaddSetupTask{c, d_pts, d_bl, d_tape, tape_stride, itape});
      launch(st, (long long)n1, JobsATask{c});
      launch(st, (long long)n1, TomCommitTask{c.s1_jv, c.s1_jr, c.tg_tab, c.th_tab, c.s1_proj, c.tom_w, c.tom_nwin});
      launch_tom_norm(st, c.s1_proj, c.s1_aff, c.s1_bytes, (long long)n1, 1);
      launch(st, ((long long)Bc + ITEM_INV_CHUNK - 1) / ITEM_INV_CHUNK, ItemInvTask{c});
      launch(st, Bc, ItemScalarsTask{c});
      const size_t nj = (size_t)Bc * JOBS_PER_ITEM, nd = (size_t)Bc * DERS_PER_ITEM;
      launch(st, (long long)Bc * GJOBS_PER_ITEM, TomCommitGTask{c.s2_jv, c.tg_tab, gext, c.tom_w, c.tom_nwin});
      launch(st, (long long)nj, TomCommitHTask{c.s2_jr, c.th_tab, gext, c.s2_proj, c.tom_w, c.tom_nwin});
      launch_tom_norm(st, c.s2_proj, c.s2_aff, c.s2_bytes, (long long)nj, 1, JOBS_PER_ITEM, 2);
      launch(st, Bc, DerivedTask{c});
      launch_tom_norm(st, c.s2_proj + nj * TOM_PROJ_WORDS, nullptr, c.s2_bytes + nj * BSTRIDE, (long long)nd, 0);
      launch(st, (long long)Bc * HASHES_PER_ITEM, ItemHashTask{c});
      launch(st, (long long)Bc * 7, ItemEmitTask{c});
      launch(st, Bc, PaddExtractTask{c, d_com, d_prf});
      if (!cd) copy_d2h(st, commitments + (size_t)b0 * 6 * WP, d_com, (size_t)Bc * 6 * WP);
      if (!pd) copy_d2h(st, proofs + (size_t)b0 * PA_LEN, d_prf, (size_t)Bc * PA_LEN);
      if (!sd) copy_d2h(st, status + b0, c.status, (size_t)Bc * 4);
      sync(st);
    }
    return 0;
  } catch (const std::exception& e) {
    return fail(ctx, ZKA_E_CUDA, e.what());
  }
}

size_t zka_verify_tape_len_ex(uint32_t ring_size, uint32_t sec_level, uint32_t samples) {
  return verify_tape_len(ceil_log2(ring_size), (int)sec_level, (int)samples);
}

int zka_verify_batch(zka_ctx* ctx, const zka_params* P, uint32_t B, const uint8_t* msg_hash, const uint8_t* ring,
                     uint32_t N, const uint8_t* proofs, size_t proof_stride, const uint32_t* proof_len,
                     const uint8_t* tape, size_t tape_stride, uint8_t* ok, int32_t* status) {
  // verifySignatureList hard-codes secparam = 20 (zkpAttestList.ts:177)
  return zka_verify_batch_ex(ctx, P, B, msg_hash, ring, N, proofs, proof_stride, proof_len, tape, tape_stride, ok, status, V_SAMPLES);
}

static int verify_impl(zka_ctx* ctx, const zka_params* P, uint32_t B, const uint8_t* msg_hash, const uint8_t* ring,
                       uint32_t N, const uint8_t* proofs, size_t proof_stride, const uint32_t* proof_len,
                       const uint8_t* tape, size_t tape_stride, uint8_t* ok, int32_t* status, uint32_t samples, int mode,
                       const uint8_t* q_ext);

int zka_verify_batch_ex(zka_ctx* ctx, const zka_params* P, uint32_t B, const uint8_t* msg_hash, const uint8_t* ring,
                        uint32_t N, const uint8_t* proofs, size_t proof_stride, const uint32_t* proof_len,
                        const uint8_t* tape, size_t tape_stride, uint8_t* ok, int32_t* status, uint32_t samples) {
  if (!msg_hash || !ring) return ZKA_E_ARG;
  return verify_impl(ctx, P, B, msg_hash, ring, N, proofs, proof_stride, proof_len, tape, tape_stride, ok, status, samples, 0, nullptr);
}

// mode 0: verifySignatureList; mode 1: verifyExp alone on assembled rows (msg_hash / ring unused, Q from q_ext)
static int verify_impl(zka_ctx* ctx, const zka_params* P, uint32_t B, const uint8_t* msg_hash, const uint8_t* ring,
                       uint32_t N, const uint8_t* proofs, size_t proof_stride, const uint32_t* proof_len,
                       const uint8_t* tape, size_t tape_stride, uint8_t* ok, int32_t* status, uint32_t samples, int mode,
                       const uint8_t* q_ext) {
  if (!ctx || !P || !proofs || !proof_len || !tape || !ok || !status) return ZKA_E_ARG;
  if (B == 0) return 0;
  if (N < 2 || N > (1u << 20)) return fail(ctx, ZKA_E_ARG, "ring size must be in [2, 2^20]");
  const int S = (int)P->sec_level;
  const int K = (int)samples;
  if (K < 1) return fail(ctx, ZKA_E_ARG, "samples must be >= 1");
  // verifyExp throws 'security level not achieved' when secparam > pi.length (exp.ts:243-245)
  if (S < K) return fail(ctx, ZKA_E_ARG, "security level not achieved");
  const int n = ceil_log2(N);
  if (tape_stride < (mode == 1 ? (size_t)V_IDX_PAD + (size_t)32 * 25 * K : verify_tape_len(n, S, K)))
    return fail(ctx, ZKA_E_ARG, "tape_stride < zka_verify_tape_len");
  try {
    if (mode == 0) {
      Stream& st0 = ctx->st;
      const uint8_t* d_ring = stage_in(st0, ctx->ring_in, ring, (size_t)N * 32);
      uint32_t* rm = ctx->ring_m.get<uint32_t>(((size_t)1 << n) * 8);
      launch(st0, 1ll << n, RingPrepTask{d_ring, rm, (int)N});
      sync(st0);
    }
    const uint32_t* ring_m = (const uint32_t*)ctx->ring_m.p;
    const int lanes = ctx->nlanes;
    const bool all_dev = is_device_ptr(proofs) && is_device_ptr(tape);
    // (two equal chunks per lane instead of the tapered host schedule were measured: config2 e2e unchanged at ~105 k
    // verifies/s, config1 57 k -> 36 k; gpurun_out/bench_c{1,2}_r2m.json)
    const std::vector<uint32_t> off = chunk_schedule(B, (uint32_t)std::min(ctx->chunk, all_dev ? 4096 : std::min(4096, ctx->host_chunk)), lanes, !all_dev);
    const uint32_t nchunks = (uint32_t)off.size() - 1;
    const int used = (int)std::min<uint32_t>((uint32_t)lanes, nchunks);
    std::atomic<uint32_t> next_chunk((uint32_t)used);
    const bool trace = getenv("ZKA_TRACE") != nullptr;
    const auto t_call = std::chrono::steady_clock::now();
    auto ms_now = [&] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_call).count(); };
    // every lane starts with chunk `li` and then claims chunks from a shared counter: copy-in, kernels and copy-out of a chunk are sequential on the
    // lane's stream; the copies of one lane overlap the kernels of the others
    // the inputs of a lane's NEXT chunk travel on the copy-in stream (second set of staging buffers) while the current
    // chunk computes; a chunk's kernels wait for its event only
    struct VIn { const uint8_t* msg; const uint8_t* proofs; const uint32_t* plen; const uint8_t* tape; };
    auto stage_chunk = [&](Lane& ln, int slot, uint32_t kk) {
      // ONE copy-in stream for all lanes of the call: the chunks' inputs cross PCIe in the order they were queued, each at
      // full bandwidth (with a copy stream per lane the first chunks and the prefetched ones were all in flight at once).
      std::lock_guard<std::mutex> copy_lock(ctx->copy_mu);
      Stream& ci = ctx->cs_in;
      DevBuf* in = ln.in + 8 * slot;
      const uint32_t b0 = off[kk];
      const size_t Bc = off[kk + 1] - b0;
      VIn v;
      v.msg = msg_hash ? stage_in(ci, in[0], msg_hash + (size_t)b0 * 32, Bc * 32) : nullptr;
      if (is_device_ptr(proofs) || is_device_ptr(proof_len)) {
        v.proofs = stage_in(ci, in[1], proofs + (size_t)b0 * proof_stride, Bc * proof_stride);
      } else {
        // host rows: only the bytes up to the longest proof of the chunk cross PCIe (rows are stride-padded;
        // a length above the stride is rejected by VLayoutTask without reading the row)
        size_t w = 0;
        for (size_t i = 0; i < Bc; i++) w = std::max<size_t>(w, proof_len[b0 + i]);
        w = std::min(proof_stride, (w + 15) & ~(size_t)15);
        uint8_t* dp = in[1].get<uint8_t>(Bc * proof_stride);
        copy_d2h_2d(ci, dp, proof_stride, proofs + (size_t)b0 * proof_stride, proof_stride, w, Bc);
        v.proofs = dp;
      }
      v.plen = stage_in(ci, in[2], proof_len + b0, Bc);
      v.tape = stage_in(ci, in[4], tape + (size_t)b0 * tape_stride, Bc * tape_stride);
      ev_record(ln.ev_small[slot], ci);
      return v;
    };
    // the first chunk of every lane is queued here, in chunk order, before any lane prefetches its second one (the trace of
    // the first version: a lane queued its first chunk and its prefetch back to back, another lane's first inputs landed
    // after 27 ms)
    std::vector<VIn> first((size_t)used);
    for (int li = 0; li < used; li++) first[(size_t)li] = stage_chunk(ctx->lane(li), 0, (uint32_t)li);
    auto run_lane = [&](int li) {
      Lane& ln = ctx->lane(li);
      Stream& st = ln.st;
      DevBuf* W = ln.w;
      uint32_t k = (uint32_t)li;
      if (k >= nchunks) return;
      int slot = 0;
      VIn cur = first[(size_t)li];
      for (;;) {
      // claim the next chunk now and send its inputs on their way (the other slot's buffers were last read by the chunk
      // before this one, which ended with a stream synchronisation)
      const uint32_t kn = next_chunk.fetch_add(1);
      VIn nxt{};
      if (kn < nchunks) nxt = stage_chunk(ln, slot ^ 1, kn);
      ev_wait(st, ln.ev_small[slot]);
      const uint32_t b0 = off[k];
      const int Bc = (int)(off[k + 1] - b0);
      const double t_begin = trace ? ms_now() : 0.0;
      VerifyCtx c;
      memset(&c, 0, sizeof(c));
      c.B = Bc; c.S = S; c.N = (int)N; c.n = n; c.K = K; c.mode = mode;
      c.q_ext = q_ext ? q_ext + (size_t)b0 * NP : nullptr;
      c.tom_w = ctx->tom_w; c.tom_nwin = ctx->tom_nwin;
      c.msg_hash = cur.msg;
      c.proofs = cur.proofs;
      c.proof_stride = proof_stride;
      c.proof_len = cur.plen;
      c.tape = cur.tape;
      c.tape_stride = tape_stride;
      c.ring_m = ring_m;
      c.g_tab8 = ctx->g8.tab; c.h_tab8 = P->h8.tab; c.h_w = P->h_w;
      c.tg_tab = ctx->tg.tab; c.th_tab = P->th.tab;
      c.tg_bytes = (const uint8_t*)ctx->tg_bytes.p;
      double t_in = 0.0;
      if (trace) { sync(st); t_in = ms_now(); }
      const size_t ns = (size_t)Bc * K;
      const int ET = c.ent_tom(), EN = c.ent_nist(), SG = c.segs();
      const int ngk = 4 * n + 1;
      c.rep_off = W[0].get<uint32_t>((size_t)Bc * S);
      c.gk_off = W[1].get<uint32_t>(Bc);
      c.tagbits = W[2].get<uint32_t>((size_t)Bc * 3);
      c.chal = W[3].get<uint32_t>((size_t)Bc * 3);
      c.gk_ok_len = W[4].get<uint8_t>(Bc);
      c.r_aff = W[5].get<uint32_t>((size_t)Bc * 16);
      c.q_aff = W[6].get<uint32_t>((size_t)Bc * 16);
      c.q_inf = W[7].get<uint8_t>(Bc);
      c.rpows = W[8].get<uint32_t>((size_t)Bc * RT_NWIN * P256_PROJ_WORDS);
      c.rrows = W[9].get<uint32_t>((size_t)Bc * RT_ENTRIES * P256_PROJ_WORDS);
      c.rtab = W[10].get<uint32_t>((size_t)Bc * RT_ENTRIES * P256_AFF_WORDS);
      c.samp_idx = W[11].get<uint32_t>(ns);
      c.samp_draw = W[12].get<uint32_t>(ns);
      c.sp_T = W[13].get<uint32_t>(ns * P256_PROJ_WORDS);
      c.sp_T_aff = W[14].get<uint32_t>(ns * 16);
      c.sp_T_inf = W[15].get<uint8_t>(ns);
      c.ta_jv = W[16].get<uint32_t>(ns * 2 * 8);
      c.ta_jr = W[17].get<uint32_t>(ns * 2 * 8);
      c.ta_proj = W[18].get<uint32_t>(ns * 2 * TOM_PROJ_WORDS);
      c.ta_aff = W[19].get<uint32_t>(ns * 2 * TOM_AFF_WORDS);
      c.td_proj = W[20].get<uint32_t>(ns * DERS_PER_ITEM * TOM_PROJ_WORDS);
      c.td_aff = W[21].get<uint32_t>(ns * DERS_PER_ITEM * TOM_AFF_WORDS);
      c.td_bytes = W[22].get<uint8_t>(ns * DERS_PER_ITEM * BSTRIDE);
      c.item_chal = W[23].get<uint32_t>(ns * HASHES_PER_ITEM * 3);
      c.ent_scalar = W[24].get<uint32_t>((size_t)Bc * ET * 8);
      c.ent_off = W[25].get<uint32_t>((size_t)Bc * ET);
      c.ent_pre = W[26].get<uint32_t>((size_t)Bc * ET * TOM_PRE_WORDS);
      c.ent_cnt = W[27].get<uint32_t>(ns);
      c.part = W[28].get<uint32_t>(ns * V_PART_WORDS);
      c.nent_scalar = W[29].get<uint32_t>((size_t)Bc * EN * 8);
      c.nent_aff = W[30].get<uint32_t>((size_t)Bc * EN * 16);
      c.nent_skip = W[31].get<uint8_t>((size_t)Bc * EN);
      c.gk_scalar = W[32].get<uint32_t>((size_t)Bc * ngk * 8);
      c.gk_pre = W[33].get<uint32_t>((size_t)Bc * ngk * TOM_PRE_WORDS);
      uint32_t* gk_offs = W[34].get<uint32_t>((size_t)Bc * ngk);
      c.fx_jv = W[35].get<uint32_t>((size_t)Bc * 2 * 8);
      c.fx_jr = W[36].get<uint32_t>((size_t)Bc * 2 * 8);
      c.fx_proj = W[37].get<uint32_t>((size_t)Bc * 2 * TOM_PROJ_WORDS);
      c.nfix = W[38].get<uint32_t>((size_t)Bc * P256_PROJ_WORDS);
      c.win_w = W[39].get<uint32_t>((size_t)Bc * SG * MSM_NWIN * 36);
      c.win_g = W[42].get<uint32_t>((size_t)Bc * MSM_NWIN * 36);
      c.win_n = W[43].get<uint32_t>((size_t)Bc * MSM_NWIN_N * P256_PROJ_WORDS);
      c.id_flags = W[44].get<uint8_t>((size_t)Bc * 3);
      c.ok = is_device_ptr(ok) ? ok + b0 : ln.out[0].get<uint8_t>(Bc);
      c.status = is_device_ptr(status) ? status + b0 : ln.out[1].get<int32_t>(Bc);

      launch(st, Bc, VLayoutTask{c});
      launch(st, (long long)Bc * (S + 1), VValidateTask{c});
      {
        // the per-proof tables of R (a 255-doubling chain per proof, rows, normalisation: no status writes) run beside the
        // Fiat-Shamir hash of the repetitions (one thread per proof, 16 KB) unless per-kernel profiling is on
        const bool fork = !st.profiling;
        Stream& sr = fork ? ln.aux[0] : st;
        if (fork) { ev_record(ln.ev_fork, st); ev_wait(sr, ln.ev_fork); }
        launch(sr, Bc, P256PowsTask{c.r_aff, nullptr, c.rpows, Bc, RT_NWIN, RT_W});
        launch(sr, (long long)Bc * RT_NWIN, P256RowsSignedTask{c.rpows, c.rrows});
        launch_p256_norm(sr, c.rrows, c.rtab, nullptr, nullptr, (long long)Bc * RT_ENTRIES);
        launch(st, Bc, VChallengeTask{c});
        if (fork) { ev_record(ln.ev_join[0], sr); ev_wait(st, ln.ev_join[0]); }
      }
      // the Groth-Kohlweiss chain (ring polynomial, relations, offsets) only needs the layout: it runs on a side stream
      // beside the sampled-repetition chain; its tape-range status is folded in by VReduceTask (same precedence)
      const bool gk_fork = mode == 0 && !st.profiling;
      Stream& sg = gk_fork ? ln.aux[1] : st;
      auto gk_chain = [&] {
        const int nblk = 1 << (n - gk_block_bits(n));
        c.gk_part = nblk > 1 ? W[51].get<uint32_t>((size_t)Bc * nblk * 8) : nullptr;
        if (nblk > 1) launch(sg, (long long)Bc * nblk, VGkSumTask{c});
        launch(sg, Bc, VGkTask{c});
        launch(sg, (long long)Bc * ngk, VGkOffsetsTask{c, gk_offs});
      };
      if (mode == 0) c.gk_tape_bad = W[54].get<uint8_t>(Bc);
      if (gk_fork) {
        ev_record(ln.ev_fork, st);
        ev_wait(sg, ln.ev_fork);
        gk_chain();
        ev_record(ln.ev_join[1], sg);
      }
      launch(st, (long long)ns, VSampleP256Task{c});
      launch_p256_norm(st, c.sp_T, c.sp_T_aff, nullptr, c.sp_T_inf, (long long)(ns));
      launch(st, (long long)ns, VSampleJobsTask{c});
      launch(st, (long long)ns * 2, TomCommitTask{c.ta_jv, c.ta_jr, c.tg_tab, c.th_tab, c.ta_proj, c.tom_w, c.tom_nwin});
      launch_tom_norm(st, c.ta_proj, c.ta_aff, nullptr, (long long)(ns * 2), 1);
      launch(st, (long long)ns, VDerivedTask{c});
      launch_tom_norm(st, c.td_proj, nullptr, c.td_bytes, (long long)(ns * DERS_PER_ITEM), 0);
      launch(st, (long long)ns * HASHES_PER_ITEM, VItemHashTask{c});
      dev_memset(st, c.ent_off, 0, (size_t)Bc * ET * 4);
      launch(st, (long long)ns, VRelationsTask{c});
      if (mode == 0) {
        if (gk_fork) ev_wait(st, ln.ev_join[1]);
        else gk_chain();
      }
      launch(st, Bc, VReduceTask{c});
      launch(st, (long long)Bc * ET, VParseEntriesTask{c.proofs, proof_stride, c.ent_off, c.ent_pre, ET});
      if (mode == 0) launch(st, (long long)Bc * ngk, VParseEntriesTask{c.proofs, proof_stride, gk_offs, c.gk_pre, ngk});
      launch(st, (long long)Bc * 2, TomCommitTask{c.fx_jv, c.fx_jr, c.tg_tab, c.th_tab, c.fx_proj, c.tom_w, c.tom_nwin});
      // chunk-wide aggregate check (zk_verify_agg.cuh): the sum over all proofs of the chunk of the three linear
      // combinations, as ONE wide-window MSM per group; when both sums are the identity the per-proof MSMs below
      // return at once
      uint32_t* ctl = nullptr;
      if (ctx->agg && mode == 0) {
        DevBuf* A = ln.agg;
        ctl = A[0].get<uint32_t>(AGG_CTL_WORDS);
        dev_memset(st, ctl, 0, AGG_CTL_WORDS * 4);
        launch(st, Bc, AggGateTask{c, ctl});
        const AggTomSrc tsrc{c.ent_scalar, c.ent_pre, c.ent_cnt, c.gk_scalar, c.gk_pre, Bc, ET, K, ngk};
        const AggNistSrc nsrc{c.nent_scalar, c.nent_aff, c.nent_skip, Bc, EN};
        // three independent chains from here to AggFinalTask: the tomEdwards256 MSM (this stream), the torsion guard and the
        // P-256 MSM with its fixed parts (two side streams; with per-kernel profiling on, everything stays on one stream
        // so that the event pairs time one kernel at a time).  A skip flag raised by the torsion guard may reach the MSM
        // kernels late — they then only do work AggFinalTask discards.
        const bool fork = !st.profiling;
        Stream& sa = fork ? ln.aux[0] : st;
        Stream& sb = fork ? ln.aux[1] : st;
        if (fork) {
          ev_record(ln.ev_fork, st);
          ev_wait(sa, ln.ev_fork);
          ev_wait(sb, ln.ev_fork);
        }
#if !defined(ZKA_PG_WAR256)
        {   // cofactor 4: no small-order components, or the per-proof path decides
          uint32_t* tpart = A[46].get<uint32_t>((size_t)Bc * (K + 1) * 2 * PG_EXT_WORDS);
          launch(sa, (long long)Bc * (K + 1) * 2, AggTorsionPartTask{tsrc, ctl, tpart});
          launch(sa, Bc, AggTorsionTask{tpart, ctl, K});
        }
#endif
        const AggPlan tp = agg_plan((double)Bc * (0.5 * K * V_ENT_PER_SAMPLE + 2 + ngk), ctx->agg_c);
        const AggPlan np = agg_plan((double)Bc * EN, 0);
        ctx->agg_c_last = tp.D.c;
        const uint32_t *tA, *tB, *nA, *nB;
        agg_msm(st, A + 1, tsrc, tp, ctl, &tA, &tB);
        agg_msm(sb, A + 20, nsrc, np, ctl, &nA, &nB);
        // fixed-base parts: one commitment for the summed tomEdwards256 scalars, a two-level sum of the P-256 points
        const int fgroups = (Bc * 2 + 63) / 64, ngroups = (Bc + 31) / 32;
        uint32_t* fpart = A[40].get<uint32_t>((size_t)fgroups * 16);
        uint32_t* fjv = A[41].get<uint32_t>(8);
        uint32_t* fjr = A[42].get<uint32_t>(8);
        uint32_t* fproj = A[43].get<uint32_t>(TOM_PROJ_WORDS);
        uint32_t* npart = A[44].get<uint32_t>((size_t)ngroups * P256_PROJ_WORDS);
        launch(st, fgroups, AggFixPartTask{ctl, c.fx_jv, c.fx_jr, fpart, Bc});
        launch(st, 1, AggFixSumTask{ctl, fpart, fjv, fjr, fgroups});
        launch(st, 1, TomCommitTask{fjv, fjr, c.tg_tab, c.th_tab, fproj, c.tom_w, c.tom_nwin});
        launch(sb, ngroups, AggNistFixPartTask{ctl, c.nfix, npart, Bc});
        int nleft = ngroups;            // second level: at most Bc / 1024 partial sums reach the final thread
        const uint32_t* nsum = npart;
        if (nleft > 32) {
          uint32_t* npart2 = A[45].get<uint32_t>((size_t)((nleft + 31) / 32) * P256_PROJ_WORDS);
          launch(sb, (nleft + 31) / 32, AggNistFixPartTask{ctl, npart, npart2, nleft});
          nsum = npart2;
          nleft = (nleft + 31) / 32;
        }
        if (fork) {
          ev_record(ln.ev_join[0], sa);
          ev_record(ln.ev_join[1], sb);
          ev_wait(st, ln.ev_join[0]);
          ev_wait(st, ln.ev_join[1]);
        }
        launch(st, 33, AggFinalTask{ctl, tA, tB, fproj, tp.D.nwin, tp.D.c, nA, nB, nsum, np.D.nwin, np.D.c, nleft});
        c.agg_ctl = ctl;
      }
      {
        const int nW = Bc * SG * MSM_NWIN, nWp = (nW + 31) & ~31, nG = Bc * MSM_NWIN;
        launch(st, (long long)nWp + nG,
               MsmTomWindowBothTask{MsmTomWindowTask{c.ent_scalar, c.ent_pre, c.ent_cnt, ET, K, V_ENT_PER_SAMPLE, 2, V_SEG, SG, c.win_w},
                                    MsmTomWindowTask{c.gk_scalar, c.gk_pre, nullptr, ngk, 0, 0, mode == 0 ? ngk : 0, V_SEG, 1, c.win_g}, nW, nWp, nG, ctl});
      }
      launch(st, (long long)Bc * MSM_NWIN_N, MsmP256WindowTask{c.nent_scalar, c.nent_aff, c.nent_skip, c.win_n, EN, ctl});
      {
        const int Bp = (Bc + 31) & ~31;
        launch(st, 3ll * Bp, MsmCombineAllTask{MsmTomCombineTask{c.win_g, c.fx_proj, c.id_flags, 2, 0, 0},
                                               MsmTomCombineTask{c.win_w, c.fx_proj, c.id_flags, 2, 1, 1, SG},
                                               MsmP256CombineTask{c.win_n, c.nfix, c.id_flags}, Bc, Bp, ctl});
      }
      launch(st, Bc, VFinalTask{c});
      if (!is_device_ptr(ok)) copy_d2h(st, ok + b0, c.ok, (size_t)Bc);
      if (!is_device_ptr(status)) copy_d2h(st, status + b0, c.status, (size_t)Bc * 4);
      uint32_t hctl[AGG_CTL_WORDS] = {0, 0, 0, 0};
      if (ctl) copy_d2h(st, hctl, ctl, sizeof(hctl));
      const double t_enq = trace ? ms_now() : 0.0;
      sync(st);
      if (trace)
        fprintf(stderr, "VTRACE lane %d chunk %u rows %d begin %.2f inputs_on_device %.2f enqueued %.2f done %.2f\n", li, k, Bc, t_begin, t_in,
                t_enq, ms_now());
      if (ctl) {
        std::lock_guard<std::mutex> g(ctx->stat_mu);
        if (hctl[AGG_TOM_PASS] && hctl[AGG_NIST_PASS]) ctx->agg_pass++;
        else ctx->agg_fail++;
      }
      if (kn >= nchunks) break;
      k = kn;
      cur = nxt;
      slot ^= 1;
    }
    };
    run_lanes(ctx, used, run_lane);
    sync(ctx->cs_in);
    return 0;
  } catch (const std::exception& e) {
    return fail(ctx, ZKA_E_CUDA, e.what());
  }
}

// ------------------------------------------------------------------ stand-alone sub-proof verifiers
// verifyExp(paramsNIST = (p256, base, NistGroup.h), paramsWario = ProofGroup, Clambda, Px, Py, pi, secparam, Q?)
// (exp.ts:233-349) for B independent statements.  The repetitions arrive in the flat layout of include/zkattest.h.
int zka_verify_exp_batch(zka_ctx* ctx, const zka_params* P, uint32_t B, const uint8_t* base, const uint8_t* com,
                         const uint8_t* px, const uint8_t* py, const uint8_t* q, const uint8_t* proofs, size_t proof_stride,
                         const uint32_t* proof_len, const uint8_t* tape, size_t tape_stride, uint32_t samples, uint8_t* ok,
                         int32_t* status) {
  if (!ctx || !P || !base || !com || !px || !py || !proofs || !proof_len || !tape || !ok || !status || proof_stride == 0) return ZKA_E_ARG;
  if (B == 0) return 0;
  try {
    Stream& st = ctx->st;
    DevBuf bufs[8];
    const uint8_t* d_base = stage_in(st, bufs[0], base, (size_t)B * NP);
    const uint8_t* d_com = stage_in(st, bufs[1], com, (size_t)B * NP);
    const uint8_t* d_px = stage_in(st, bufs[2], px, (size_t)B * WP);
    const uint8_t* d_py = stage_in(st, bufs[3], py, (size_t)B * WP);
    const uint8_t* d_q = q ? stage_in(st, bufs[4], q, (size_t)B * NP) : nullptr;
    const uint8_t* d_body = stage_in(st, bufs[5], proofs, (size_t)B * proof_stride);
    const uint32_t* d_len = stage_in(st, bufs[6], proof_len, (size_t)B);
    const uint8_t* d_tape = stage_in(st, bufs[7], tape, (size_t)B * tape_stride);
    const size_t row_stride = (HEAD_LEN + proof_stride + 15) & ~(size_t)15;
    DevBuf rows, rlen;
    uint8_t* d_rows = rows.get<uint8_t>((size_t)B * row_stride);
    uint32_t* d_rlen = rlen.get<uint32_t>(B);
    const int pieces = (int)((row_stride + 63) / 64);
    launch(st, (long long)B * pieces, VAssembleTask{d_base, d_com, d_px, d_py, d_body, proof_stride, d_len, d_rows, row_stride, d_rlen, pieces});
    sync(st);
    const int rc = verify_impl(ctx, P, B, nullptr, nullptr, 2, d_rows, row_stride, d_rlen, d_tape, tape_stride, ok, status, samples, 1, d_q);
    for (auto& b : bufs) b.release();
    rows.release();
    rlen.release();
    return rc;
  } catch (const std::exception& e) {
    return fail(ctx, ZKA_E_CUDA, e.what());
  }
}

// verifyMembership(ProofGroup params, com, ring, proof) (gk.ts:197-262) for B commitments over one ring.
// tape: the 2n+1 Relation.drain scalars per proof, in call order.
int zka_verify_membership_batch(zka_ctx* ctx, const zka_params* P, uint32_t B, const uint8_t* com, const uint8_t* ring, uint32_t N,
                                const uint8_t* proofs, size_t proof_stride, const uint32_t* proof_len, const uint8_t* tape,
                                size_t tape_stride, uint8_t* ok, int32_t* status) {
  if (!ctx || !P || !com || !ring || !proofs || !proof_len || !tape || !ok || !status || proof_stride == 0) return ZKA_E_ARG;
  if (B == 0) return 0;
  if (N < 2 || N > (1u << 20)) return fail(ctx, ZKA_E_ARG, "ring size must be in [2, 2^20]");
  const int n = ceil_log2(N);
  if (tape_stride < (size_t)32 * (2 * n + 1)) return fail(ctx, ZKA_E_ARG, "tape_stride < 32 * (2n + 1)");
  try {
    Stream& st = ctx->st;
    DevBuf* W = ctx->w;
    DevBuf bufs[4];
    const uint8_t* d_ring = stage_in(st, ctx->ring_in, ring, (size_t)N * 32);
    uint32_t* ring_m = ctx->ring_m.get<uint32_t>(((size_t)1 << n) * 8);
    launch(st, 1ll << n, RingPrepTask{d_ring, ring_m, (int)N});
    const size_t row_stride = (HEAD_LEN + proof_stride + 15) & ~(size_t)15;
    const int pieces = (int)((row_stride + 63) / 64);
    const int ngk = 4 * n + 1;
    const uint32_t chunk = 4096;
    for (uint32_t b0 = 0; b0 < B; b0 += chunk) {
      const int Bc = (int)std::min<uint32_t>(chunk, B - b0);
      const uint8_t* d_com = stage_in(st, bufs[0], com + (size_t)b0 * WP, (size_t)Bc * WP);
      const uint8_t* d_body = stage_in(st, bufs[1], proofs + (size_t)b0 * proof_stride, (size_t)Bc * proof_stride);
      const uint32_t* d_len = stage_in(st, bufs[2], proof_len + b0, (size_t)Bc);
      VerifyCtx c;
      memset(&c, 0, sizeof(c));
      c.B = Bc; c.S = (int)P->sec_level; c.N = (int)N; c.n = n; c.K = 1; c.mode = 2;
      c.tom_w = ctx->tom_w; c.tom_nwin = ctx->tom_nwin;
      c.tape = stage_in(st, bufs[3], tape + (size_t)b0 * tape_stride, (size_t)Bc * tape_stride);
      c.tape_stride = tape_stride;
      c.ring_m = ring_m;
      c.tg_tab = ctx->tg.tab; c.th_tab = P->th.tab;
      uint8_t* d_rows = W[0].get<uint8_t>((size_t)Bc * row_stride);
      uint32_t* d_rlen = W[1].get<uint32_t>(Bc);
      c.proofs = d_rows; c.proof_stride = row_stride; c.proof_len = d_rlen;
      c.gk_off = W[2].get<uint32_t>(Bc);
      c.gk_ok_len = W[3].get<uint8_t>(Bc);
      c.gk_scalar = W[4].get<uint32_t>((size_t)Bc * ngk * 8);
      c.gk_pre = W[5].get<uint32_t>((size_t)Bc * ngk * TOM_PRE_WORDS);
      uint32_t* gk_offs = W[6].get<uint32_t>((size_t)Bc * ngk);
      c.fx_jv = W[7].get<uint32_t>((size_t)Bc * 2 * 8);
      c.fx_jr = W[8].get<uint32_t>((size_t)Bc * 2 * 8);
      c.fx_proj = W[9].get<uint32_t>((size_t)Bc * 2 * TOM_PROJ_WORDS);
      c.win_g = W[10].get<uint32_t>((size_t)Bc * MSM_NWIN * 36);
      c.id_flags = W[11].get<uint8_t>((size_t)Bc * 3);
      c.ok = is_device_ptr(ok) ? ok + b0 : ctx->out[0].get<uint8_t>(Bc);
      c.status = is_device_ptr(status) ? status + b0 : ctx->out[1].get<int32_t>(Bc);
      launch(st, (long long)Bc * pieces, VAssembleTask{nullptr, nullptr, d_com, nullptr, d_body, proof_stride, d_len, d_rows, row_stride, d_rlen, pieces});
      launch(st, Bc, VGkOnlyLayoutTask{c});
      dev_memset(st, c.fx_jv, 0, (size_t)Bc * 2 * 8 * 4);
      dev_memset(st, c.fx_jr, 0, (size_t)Bc * 2 * 8 * 4);
      {
        const int nblk = 1 << (n - gk_block_bits(n));
        c.gk_part = nblk > 1 ? W[12].get<uint32_t>((size_t)Bc * nblk * 8) : nullptr;
        if (nblk > 1) launch(st, (long long)Bc * nblk, VGkSumTask{c});
      }
      launch(st, Bc, VGkTask{c});
      launch(st, (long long)Bc * ngk, VGkOffsetsTask{c, gk_offs});
      launch(st, (long long)Bc * ngk, VParseEntriesTask{c.proofs, row_stride, gk_offs, c.gk_pre, ngk});
      launch(st, (long long)Bc * 2, TomCommitTask{c.fx_jv, c.fx_jr, c.tg_tab, c.th_tab, c.fx_proj, c.tom_w, c.tom_nwin});
      launch(st, (long long)Bc * MSM_NWIN, MsmTomWindowTask{c.gk_scalar, c.gk_pre, nullptr, ngk, 0, 0, ngk, V_SEG, 1, c.win_g});
      launch(st, Bc, MsmTomCombineTask{c.win_g, c.fx_proj, c.id_flags, 2, 0, 0});
      launch(st, Bc, VGkOnlyFinalTask{c});
      if (!is_device_ptr(ok)) copy_d2h(st, ok + b0, c.ok, (size_t)Bc);
      if (!is_device_ptr(status)) copy_d2h(st, status + b0, c.status, (size_t)Bc * 4);
      sync(st);
    }
    for (auto& b : bufs) b.release();
    return 0;
  } catch (const std::exception& e) {
    return fail(ctx, ZKA_E_CUDA, e.what());
  }
}

// verifyEquality / verifyMult / verifyPointAdd alone (kind = 0 / 1 / 2): fixed-size inputs and proofs
static int verify_sub(zka_ctx* ctx, const zka_params* P, int kind, uint32_t B, const uint8_t* points, const uint8_t* proofs,
                      const uint8_t* tape, size_t tape_stride, uint8_t* ok, int32_t* status) {
  if (!ctx || !P || !points || !proofs || !tape || !ok || !status) return ZKA_E_ARG;
  if (B == 0) return 0;
  if (tape_stride < (size_t)32 * sub_draws(kind)) return fail(ctx, ZKA_E_ARG, "tape_stride too small for this sub-proof");
  try {
    Stream& st = ctx->st;
    DevBuf* W = ctx->w;
    const int la = sub_points(kind) * WP, lc = sub_proof_len(kind), ne = sub_entries(kind);
    const size_t stride = (size_t)(la + lc + 15) & ~(size_t)15;
    const uint32_t chunk = 8192;
    for (uint32_t b0 = 0; b0 < B; b0 += chunk) {
      const int Bc = (int)std::min<uint32_t>(chunk, B - b0);
      const uint8_t* d_pts = stage_in(st, ctx->in[0], points + (size_t)b0 * la, (size_t)Bc * la);
      const uint8_t* d_prf = stage_in(st, ctx->in[1], proofs + (size_t)b0 * lc, (size_t)Bc * lc);
      const uint8_t* d_tape = stage_in(st, ctx->in[2], tape + (size_t)b0 * tape_stride, (size_t)Bc * tape_stride);
      uint8_t* rows = W[0].get<uint8_t>((size_t)Bc * stride);
      uint32_t* ent_scalar = W[1].get<uint32_t>((size_t)Bc * SUB_ENT_MAX * 8);
      uint32_t* ent_off = W[2].get<uint32_t>((size_t)Bc * SUB_ENT_MAX);
      uint32_t* ent_pre = W[3].get<uint32_t>((size_t)Bc * SUB_ENT_MAX * TOM_PRE_WORDS);
      uint32_t* fx_jv = W[4].get<uint32_t>((size_t)Bc * 2 * 8);
      uint32_t* fx_jr = W[5].get<uint32_t>((size_t)Bc * 2 * 8);
      uint32_t* fx_proj = W[6].get<uint32_t>((size_t)Bc * 2 * TOM_PROJ_WORDS);
      uint32_t* win = W[7].get<uint32_t>((size_t)Bc * MSM_NWIN * 36);
      uint8_t* flags = W[8].get<uint8_t>((size_t)Bc * 3);
      uint8_t* d_ok = is_device_ptr(ok) ? ok + b0 : ctx->out[0].get<uint8_t>(Bc);
      int32_t* d_st = is_device_ptr(status) ? status + b0 : ctx->out[1].get<int32_t>(Bc);
      launch(st, (long long)Bc * (la + lc), VConcatTask{d_pts, d_prf, la, lc, rows, stride});
      launch(st, Bc, VSubProofTask{kind, rows, stride, d_tape, tape_stride, (const uint8_t*)ctx->tg_bytes.p, ent_scalar, ent_off,
                                   fx_jv, fx_jr, d_st, d_ok});
      launch(st, (long long)Bc * SUB_ENT_MAX, VParseEntriesTask{rows, stride, ent_off, ent_pre, SUB_ENT_MAX});
      launch(st, (long long)Bc * 2, TomCommitTask{fx_jv, fx_jr, ctx->tg.tab, P->th.tab, fx_proj, ctx->tom_w, ctx->tom_nwin});
      launch(st, (long long)Bc * MSM_NWIN, MsmTomWindowTask{ent_scalar, ent_pre, nullptr, SUB_ENT_MAX, 0, 0, ne, V_SEG, 1, win});
      launch(st, Bc, MsmTomCombineTask{win, fx_proj, flags, 2, 1, 1});
      launch(st, Bc, VSubFinalTask{d_st, flags, d_ok});
      if (!is_device_ptr(ok)) copy_d2h(st, ok + b0, d_ok, (size_t)Bc);
      if (!is_device_ptr(status)) copy_d2h(st, status + b0, d_st, (size_t)Bc * 4);
      sync(st);
    }
    return 0;
  } catch (const std::exception& e) {
    return fail(ctx, ZKA_E_CUDA, e.what());
  }
}
int zka_verify_equality_batch(zka_ctx* ctx, const zka_params* P, uint32_t B, const uint8_t* points, const uint8_t* proofs,
                              const uint8_t* tape, size_t tape_stride, uint8_t* ok, int32_t* status) {
  return verify_sub(ctx, P, SUB_EQ, B, points, proofs, tape, tape_stride, ok, status);
}
int zka_verify_mult_batch(zka_ctx* ctx, const zka_params* P, uint32_t B, const uint8_t* points, const uint8_t* proofs,
                          const uint8_t* tape, size_t tape_stride, uint8_t* ok, int32_t* status) {
  return verify_sub(ctx, P, SUB_MULT, B, points, proofs, tape, tape_stride, ok, status);
}
int zka_verify_pointadd_batch(zka_ctx* ctx, const zka_params* P, uint32_t B, const uint8_t* points, const uint8_t* proofs,
                              const uint8_t* tape, size_t tape_stride, uint8_t* ok, int32_t* status) {
  return verify_sub(ctx, P, SUB_PADD, B, points, proofs, tape, tape_stride, ok, status);
}

// ---------------------------------------------------------------------------- multi-GPU helpers
namespace {
int pack_common(zka_ctx* ctx, uint32_t B, uint8_t* rows, size_t stride, const uint32_t* len, uint8_t* packed, size_t cap,
                uint64_t* offsets, void* stream, int dir) {
  if (!ctx || !rows || !len || !packed || !offsets || stride == 0) return ZKA_E_ARG;
  if (B == 0) return 0;
#if !defined(ZKA_HOSTSIM)
  if (!is_device_ptr(rows) || !is_device_ptr(len) || !is_device_ptr(packed) || !is_device_ptr(offsets))
    return fail(ctx, ZKA_E_ARG, "zka_proofs_pack/unpack take device pointers");
#endif
  try {
    Stream tmp;          // borrowed stream (not owned, never destroyed here); launch counters go to lane 0
    Stream* st = &ctx->st;
#if !defined(ZKA_HOSTSIM)
    if (stream) { tmp.s = (cudaStream_t)stream; st = &tmp; }
#endif
    const int pieces = (int)((stride + 15) / 16);
    launch(*st, 1, PackScanTask{len, offsets, (int)B});
    const bool al = (stride % 16 == 0) && (((size_t)rows | (size_t)packed) % 16 == 0);
    if (al) launch(*st, (long long)B * pieces, PackCopy16Task{rows, stride, len, offsets, packed, cap, pieces, dir});
    else launch(*st, (long long)B * pieces, PackCopyTask{rows, stride, len, offsets, packed, cap, pieces, dir});
    if (st == &tmp) ctx->st.launches += tmp.launches; else sync(*st);
    return 0;
  } catch (const std::exception& e) {
    return fail(ctx, ZKA_E_CUDA, e.what());
  }
}
}  // namespace

int zka_proofs_pack(zka_ctx* ctx, uint32_t B, const uint8_t* proofs, size_t proof_stride, const uint32_t* proof_len,
                    uint8_t* packed, size_t cap, uint64_t* offsets, void* stream) {
  return pack_common(ctx, B, const_cast<uint8_t*>(proofs), proof_stride, proof_len, packed, cap, offsets, stream, 0);
}
int zka_proofs_unpack(zka_ctx* ctx, uint32_t B, const uint8_t* packed, size_t cap, const uint32_t* proof_len, uint8_t* proofs,
                      size_t proof_stride, uint64_t* offsets, void* stream) {
  return pack_common(ctx, B, proofs, proof_stride, proof_len, const_cast<uint8_t*>(packed), cap, offsets, stream, 1);
}

}  // extern "C"
