// zk_ops.cuh — building-block tasks: precomputed tables, scalar multiplication, batched
// affine normalisation + serialisation, SHA-256 over point encodings.
//
// A "task" is a functor `void operator()(int t) const` executed once per work item; the
// CUDA build runs it as one thread of a grid (zk_launch.cuh).  All tables and staging
// buffers live in HBM/L2: the path is integer-pipe bound, the staging
// traffic is < 1 % of HBM bandwidth.
//
// Reference mapping:
//   Point.mul / Point.dblmul (4-bit windows, /root/reference/src/curves/group.ts:97-152)
//     -> positional fixed-window tables, NO doublings at evaluation time:
//        k*P = sum_j T[j][digit_j(k)],  T[j][d] = d * 2^(w j) * P   (affine / precomputed)
//   toAffine + toBytes (weier.ts:231-255, edwards.ts:184-203; one invEuclid each)
//     -> Montgomery's trick over chunks of points, one Fermat inversion per chunk.
#pragma once
#include "zk_curves.cuh"
#include "zk_layout.h"
#include "zk_sha256.cuh"

namespace zk {

#if defined(__CUDA_ARCH__)
#define ZK_SET_STATUS(ptr, code) atomicCAS((int*)(ptr), 0, (int)(code))
// first error wins, except that `code` also replaces the later-stage error `over` (used when the two
// stages run side by side in one grid: the outcome equals running this stage first)
#define ZK_SET_STATUS_OVER(ptr, code, over) \
  do {                                      \
    atomicCAS((int*)(ptr), 0, (int)(code)); \
    atomicCAS((int*)(ptr), (int)(over), (int)(code)); \
  } while (0)
#else
#define ZK_SET_STATUS_OVER(ptr, code, over)              \
  do {                                                   \
    if (*(ptr) == 0 || *(ptr) == (over)) *(ptr) = (code); \
  } while (0)
#define ZK_SET_STATUS(ptr, code) \
  do {                           \
    if (*(ptr) == 0) *(ptr) = (code); \
  } while (0)
#endif

// ----------------------------------------------------------------------------- loads/stores
template <int N>
ZK_HD void ld(uint32_t* r, const uint32_t* p) {
#pragma unroll
  for (int i = 0; i < N; i++) r[i] = p[i];
}
template <int N>
ZK_HD void st(uint32_t* p, const uint32_t* r) {
#pragma unroll
  for (int i = 0; i < N; i++) p[i] = r[i];
}

enum : int {
  P256_PROJ_WORDS = 24,
  P256_AFF_WORDS = 16,
#if defined(ZKA_PG_WAR256)
  TOM_PROJ_WORDS = 24,   // war256 build: homogeneous (X, Y, Z), 8 limbs each
  TOM_AFF_WORDS = 16,    // affine (x, y), Montgomery
  TOM_PRE_WORDS = 16,    // table entry / parsed point = affine (x, y): one 64-byte half line
#else
  TOM_PROJ_WORDS = 28,   // X, Y, Z (T is not needed after the last addition) + 1 pad word: 7 x 16 bytes
  TOM_AFF_WORDS = 18,    // x', y on the a'=1 image curve, Montgomery
  TOM_PRE_WORDS = 32,    // x', y, k = d' x' y + 5 pad words: one 128-byte line per entry
#endif
  NORM_CHUNK_MAX = 64,   // max points per Montgomery-trick chunk (one Fermat inversion each)
};


#if defined(ZKA_PG_WAR256)
// staged war256 points: X, Y, Z at words 0, 8, 16 of a 96-byte slot; table entries are affine pairs
ZK_HD void tom_st_xyz(uint32_t* m, const uint32_t* x, const uint32_t* y, const uint32_t* z) {
  st<8>(m, x); st<8>(m + 8, y); st<8>(m + 16, z);
}
ZK_HD void tom_ld_xyz(uint32_t* x, uint32_t* y, uint32_t* z, const uint32_t* m) {
  ld<8>(x, m); ld<8>(y, m + 8); ld<8>(z, m + 16);
}
ZK_HD void tom_ld_pre(TomPre& q, const uint32_t* m) { ld<8>(q.x, m); ld<8>(q.y, m + 8); }
#else
// staged tomEdwards256 points (X, Y, Z at words 0, 9, 18 of a 112-byte, 16-byte aligned slot):
// seven 16-byte transactions instead of 27 four-byte ones (the slots are strided per thread)
ZK_HD void tom_st_xyz(uint32_t* m, const uint32_t* x, const uint32_t* y, const uint32_t* z) {
  uint32_t w[28];
#pragma unroll
  for (int i = 0; i < 9; i++) { w[i] = x[i]; w[9 + i] = y[i]; w[18 + i] = z[i]; }
  w[27] = 0;
#if defined(__CUDA_ARCH__)
  uint4* v = reinterpret_cast<uint4*>(m);
#pragma unroll
  for (int i = 0; i < 7; i++) v[i] = make_uint4(w[4 * i], w[4 * i + 1], w[4 * i + 2], w[4 * i + 3]);
#else
  for (int i = 0; i < 28; i++) m[i] = w[i];
#endif
}
ZK_HD void tom_ld_xyz(uint32_t* x, uint32_t* y, uint32_t* z, const uint32_t* m) {
  uint32_t w[28];
#if defined(__CUDA_ARCH__)
  const uint4* v = reinterpret_cast<const uint4*>(m);
#pragma unroll
  for (int i = 0; i < 7; i++) {
    const uint4 u = v[i];
    w[4 * i] = u.x; w[4 * i + 1] = u.y; w[4 * i + 2] = u.z; w[4 * i + 3] = u.w;
  }
#else
  for (int i = 0; i < 28; i++) w[i] = m[i];
#endif
#pragma unroll
  for (int i = 0; i < 9; i++) { x[i] = w[i]; y[i] = w[9 + i]; z[i] = w[18 + i]; }
}

// negate a table entry of the a = -1 image curve: -(w, v) = (-w, v) swaps v - w and v + w and negates 2 d2 w v
ZK_HD void tom2_pre_neg(TomPre& q, bool neg) {
  uint32_t nk[9];
  Tomp::neg(nk, q.k);
#pragma unroll
  for (int i = 0; i < 9; i++) {
    const uint32_t a = q.x[i], b = q.y[i];
    q.x[i] = neg ? b : a;
    q.y[i] = neg ? a : b;
    q.k[i] = neg ? nk[i] : q.k[i];
  }
}
ZK_HD void tom_ld_pre(TomPre& q, const uint32_t* m) {
#if defined(__CUDA_ARCH__)
  // one 128-byte line, eight 16-byte loads (entries are 128-byte aligned)
  const uint4* v = reinterpret_cast<const uint4*>(m);
  uint32_t w[32];
#pragma unroll
  for (int i = 0; i < 7; i++) {
    uint4 u = __ldg(v + i);
    w[4 * i] = u.x; w[4 * i + 1] = u.y; w[4 * i + 2] = u.z; w[4 * i + 3] = u.w;
  }
#pragma unroll
  for (int i = 0; i < 9; i++) { q.x[i] = w[i]; q.y[i] = w[9 + i]; q.k[i] = w[18 + i]; }
#else
  ld<9>(q.x, m); ld<9>(q.y, m + 9); ld<9>(q.k, m + 18);
#endif
}

#endif

// Encoded point (tag || x || y, big-endian coordinates of CB bytes) written as 32-bit words into a
// 4-byte aligned BSTRIDE slot: 17 word stores instead of 65/67 byte stores (the byte index of
// every output position is a compile-time constant after unrolling).
template <int N, int CB>
ZK_HD uint32_t enc_byte(int i, uint32_t tag, const uint32_t* x, const uint32_t* y) {
  if (i == 0) return tag;
  if (i >= 1 + 2 * CB) return 0u;
  const uint32_t* c = i <= CB ? x : y;
  const int pos = CB - 1 - (i <= CB ? i - 1 : i - 1 - CB);   // byte significance
  return (pos >> 2) < N ? ((c[pos >> 2] >> (8 * (pos & 3))) & 0xffu) : 0u;
}
template <int N, int CB>
ZK_HD void store_point_words(uint8_t* o, uint32_t tag, const uint32_t* x, const uint32_t* y) {
  uint32_t* ow = reinterpret_cast<uint32_t*>(o);
#pragma unroll
  for (int j = 0; j < BSTRIDE / 4; j++) {
    const uint32_t w = enc_byte<N, CB>(4 * j, tag, x, y) | (enc_byte<N, CB>(4 * j + 1, tag, x, y) << 8) |
                       (enc_byte<N, CB>(4 * j + 2, tag, x, y) << 16) | (enc_byte<N, CB>(4 * j + 3, tag, x, y) << 24);
    ow[j] = w;
  }
}

// digit j of width w (w in {4, 8}) of a canonical 256-bit scalar
ZK_HD uint32_t digit4(const uint32_t* k, int j) { return (k[j >> 3] >> (4 * (j & 7))) & 15u; }
ZK_HD uint32_t digit8(const uint32_t* k, int j) { return (k[j >> 2] >> (8 * (j & 3))) & 255u; }
// generic: bits [pos, pos+w) of a 256-bit scalar, w <= 24
ZK_HD uint32_t digit_w(const uint32_t* k, int pos, int w) {
  int wi = pos >> 5, sh = pos & 31;
  uint64_t v = k[wi];
  if (wi + 1 < 8) v |= (uint64_t)k[wi + 1] << 32;
  return (uint32_t)(v >> sh) & ((1u << w) - 1u);
}

// Fixed-base tables hold SIGNED digits: k = sum_j d_j 2^(w j), d_j in [-2^(w-1), 2^(w-1)], so a window stores the
// multiples 0 .. 2^(w-1) only (half the memory of unsigned digits at the same number of lookups); a negative
// digit negates the entry on the fly.  ceil(257 / w) windows: the carry out of the top data window needs one
// more window only when w divides 256.
ZK_HD int fb_entries(int w) { return (1 << (w - 1)) + 1; }
ZK_HD int fb_windows(int w) { return (256 + w) / w; }
// digit j of k in signed form: returns |d_j| and its sign; `carry` is the running carry (start with 0)
ZK_HD uint32_t signed_digit(const uint32_t* k, int j, int w, uint32_t& carry, bool& neg) {
  const int pos = j * w;
  uint32_t d = carry;
  if (pos < 256) d += digit_w(k, pos, (256 - pos) < w ? (256 - pos) : w);
  const uint32_t half = 1u << (w - 1);
  neg = d > half;
  carry = neg ? 1u : 0u;
  return neg ? (1u << w) - d : d;
}

// read a 32-byte big-endian tape draw into 8 limbs
ZK_HD void tape_draw(uint32_t* r, const uint8_t* tape, int draw) {
  const uint8_t* p = tape + 32 * (size_t)draw;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    const uint8_t* q = p + 28 - 4 * i;
    r[i] = ((uint32_t)q[0] << 24) | ((uint32_t)q[1] << 16) | ((uint32_t)q[2] << 8) | (uint32_t)q[3];
  }
}

// ============================================================================ P-256 tables
#define WEI_PT P256Pt
#define WEI_AFF P256Aff
#define WEI_JAC P256Jac
#define WEI_F P256p
#define WEI_FN(n) p256_##n
#define WEI_T(n) P256##n
#include "zk_weier_ops.inc"
#undef WEI_PT
#undef WEI_AFF
#undef WEI_JAC
#undef WEI_F
#undef WEI_FN
#undef WEI_T

// ---- per-KEY tables of the prover: the generic signed-digit format [fb_windows(w)][fb_entries(w)][16] with the window
// bits chosen ON THE DEVICE from the number of distinct keys of the chunk (KeyRankTask).  Grids and buffers are sized
// for the worst case (every proof its own key, w = 5: KEY_CAP entries per proof); fewer keys get wider windows inside
// the same memory: 256 keys in a chunk of 4096 proofs -> w = 8, 33 instead of 52 lookups per alpha*pk.
enum : int { KEY_W_MIN = 5, KEY_W_MAX = 8, KEY_CAP = 52 * 17 };
ZK_HD int key_window_bits(uint32_t count, uint32_t B, uint32_t uses) {   // uses = products per proof (S + 2)
  int best = KEY_W_MIN;
  uint64_t best_cost = ~0ull;
  for (int w = KEY_W_MIN; w <= KEY_W_MAX; w++) {
    const uint64_t nwin = (uint64_t)fb_windows(w), ne = (uint64_t)fb_entries(w);
    if (w > KEY_W_MIN && (uint64_t)count * nwin * ne > (uint64_t)B * KEY_CAP) break;
    // building an entry = one addition + its share of a normalisation (~2 mixed additions), used `uses` times per proof
    const uint64_t cost = nwin * ((uint64_t)count * (ne - 1) * 2 + (uint64_t)B * uses);
    if (cost < best_cost) { best_cost = cost; best = w; }
  }
  return best;
}
ZK_HD size_t key_table_words(int w) { return (size_t)fb_windows(w) * fb_entries(w) * P256_AFF_WORDS; }

// Per-base SIGNED 5-bit table (the per-proof table of R): k = sum_j d_j 32^j with d_j in [-16, 16],
// 52 windows x 16 entries (1..16 multiples); a negative digit negates y of the affine entry.
// 52 mixed additions per scalar multiplication instead of 64 with unsigned 4-bit windows, and the
// table is smaller (832 entries instead of 960).
enum : int { RT_W = 5, RT_NWIN = 52, RT_ROW = 16, RT_ENTRIES = RT_NWIN * RT_ROW };
struct P256RowsSignedTask {
  const uint32_t* pows;  // [nbase*RT_NWIN][24]
  uint32_t* rows;        // [nbase*RT_NWIN][RT_ROW][24]: entry d-1 = d * pows
  const uint32_t* count_dev = nullptr;   // optional: number of bases actually present
  ZK_HD void operator()(int t) const {
    if (count_dev && (uint32_t)(t / RT_NWIN) >= *count_dev) return;
    P256Pt p, acc;
    p256_ld_proj(p, pows + (size_t)t * P256_PROJ_WORDS);
    acc = p;
    uint32_t* out = rows + (size_t)t * RT_ROW * P256_PROJ_WORDS;
    for (int d = 1; d <= RT_ROW; d++) {
      p256_st_proj(out + (size_t)(d - 1) * P256_PROJ_WORDS, acc);
      if (d < RT_ROW) p256_add(acc, acc, p);
    }
  }
};
// acc += k * base on the signed 5-bit per-base table [RT_NWIN][RT_ROW] (P256RowsSignedTask)
ZK_HD void p256_accum_rtab(P256Pt& acc, const uint32_t* tab, const uint32_t* k) {
  uint32_t carry = 0;
  for (int j = 0; j < RT_NWIN; j++) {
    const int pos = j * RT_W;
    uint32_t d = (pos < 256 ? digit_w(k, pos, (256 - pos) < RT_W ? (256 - pos) : RT_W) : 0u) + carry;
    carry = d > 16 ? 1u : 0u;
    const bool neg = d > 16;
    if (neg) d = 32 - d;
    if (d) {
      P256Aff q;
      p256_ld_aff(q, tab + ((size_t)j * RT_ROW + (d - 1)) * P256_AFF_WORDS);
      if (neg) P256p::neg(q.y, q.y);
      p256_madd(acc, acc, q);
    }
  }
}
// Split commitments for the 34 jobs of a 0-bit repetition.  Several of them commit to the SAME value
// with different blinders (proveMult: A_z and A_4_1 both commit k_z, mult.ts:112-113; proveEquality:
// A_1 and A_2 both commit k, equality.ts:67-68), so the g-parts v*g are computed once per distinct
// value (28 per item) and every job continues from its g-part with the 16 lookups of r*h:
//   34 x 32 = 1088 lookups  ->  28 x 16 + 34 x 16 = 992.
#if defined(ZKA_PG_WAR256)
enum : int { GJOBS_PER_ITEM = 28, TOM_EXT_WORDS = 24 };
#else
enum : int { GJOBS_PER_ITEM = 28, TOM_EXT_WORDS = 36 };
#endif
// job index (0..33) -> index of its g-part (0..27)
ZK_HD int item_gpart_of_job(int j) {
  if (j < 6) return j;                       // T1x T1y C8 C10 C11 C13
  if (j < 30) {                              // MultProof m: C4 Ax Ay Az A4_1 A4_2 -> 0 1 2 3 3 4
    const int m = (j - 6) / 6, u = (j - 6) % 6;
    return 6 + 5 * m + (u < 4 ? u : u - 1);
  }
  return 26 + ((j - 30) >> 1);               // EqualityProof e: A1, A2 share k
}
// g-part index (0..27) -> a job that carries its value scalar
ZK_HD int item_job_of_gpart(int g) {
  if (g < 6) return g;
  if (g < 26) {
    const int m = (g - 6) / 5, u = (g - 6) % 5;
    return 6 + 6 * m + (u < 4 ? u : 5);
  }
  return 30 + 2 * (g - 26);
}
#if defined(ZKA_PG_WAR256)
#include "zk_ops_war.cuh"
#else
// ===================================================================== tomEdwards256 tables
struct TomPowsTask {
  const uint32_t* base_aff;  // [nbase][18] image-curve affine (x', y), Montgomery
  uint32_t* pows;            // [nbase][nwin][36] extended (X,Y,T,Z)
  int nbase, nwin, w;
  ZK_HD void operator()(int t) const {
    uint32_t x[9], y[9];
    ld<9>(x, base_aff + (size_t)t * TOM_AFF_WORDS);
    ld<9>(y, base_aff + (size_t)t * TOM_AFF_WORDS + 9);
    TomPt p;
    tom_from_affine(p, x, y);
    for (int j = 0; j < nwin; j++) {
      uint32_t* o = pows + ((size_t)t * nwin + j) * 36;
      st<9>(o, p.x); st<9>(o + 9, p.y); st<9>(o + 18, p.t); st<9>(o + 27, p.z);
      for (int k = 0; k < w; k++) tom_dbl(p, p);
    }
  }
};
// rows: proj store [(base*nwin + j) * 2^w + d][27] = d * pows[base][j]  (d = 0 is the identity)
struct TomRowsTask {
  const uint32_t* pows;
  uint32_t* rows;
  int w;
  ZK_HD void operator()(int t) const {
    TomPt p, acc;
    const uint32_t* s = pows + (size_t)t * 36;
    ld<9>(p.x, s); ld<9>(p.y, s + 9); ld<9>(p.t, s + 18); ld<9>(p.z, s + 27);
    tom_set_identity(acc);
    const int ne = fb_entries(w);
    uint32_t* out = rows + (size_t)t * ne * TOM_PROJ_WORDS;
    for (int d = 0; d < ne; d++) {
      uint32_t* o = out + (size_t)d * TOM_PROJ_WORDS;
      tom_st_xyz(o, acc.x, acc.y, acc.z);
      tom_add(acc, acc, p);
    }
  }
};
// Two-level construction of the same rows for wide windows (w > 8): first the 2^(w-8) "high"
// multiples m * 2^8 * P_j (one thread per window), then one thread per (window, m) walks the
// 256 entries below it.  2^(w-8) + 256 sequential additions instead of 2^w.
struct TomRowsHiTask {
  const uint32_t* pows;   // [nwin][36]
  uint32_t* hi;           // [nwin][2^(w-9)][36]
  uint32_t* rows;         // the top entry 2^(w-1) * pows of every window is written here directly
  int w;
  ZK_HD void operator()(int t) const {
    TomPt p, acc;
    const uint32_t* s = pows + (size_t)t * 36;
    ld<9>(p.x, s); ld<9>(p.y, s + 9); ld<9>(p.t, s + 18); ld<9>(p.z, s + 27);
    for (int k = 0; k < 8; k++) tom_dbl(p, p);
    tom_set_identity(acc);
    const int nh = 1 << (w - 9);
    for (int m = 0; m < nh; m++) {
      uint32_t* o = hi + ((size_t)t * nh + m) * 36;
      st<9>(o, acc.x); st<9>(o + 9, acc.y); st<9>(o + 18, acc.t); st<9>(o + 27, acc.z);
      tom_add(acc, acc, p);
    }
    tom_st_xyz(rows + ((size_t)t * fb_entries(w) + ((size_t)1 << (w - 1))) * TOM_PROJ_WORDS, acc.x, acc.y, acc.z);
  }
};
struct TomRowsLoTask {
  const uint32_t* pows;   // [nwin][36]
  const uint32_t* hi;     // [nwin][2^(w-9)][36]
  uint32_t* rows;         // [nwin][E][27]
  int w;
  ZK_HD void operator()(int t) const {
    const int nh = 1 << (w - 9);
    const int j = t / nh, m = t % nh;
    TomPt p, acc;
    const uint32_t* s = pows + (size_t)j * 36;
    ld<9>(p.x, s); ld<9>(p.y, s + 9); ld<9>(p.t, s + 18); ld<9>(p.z, s + 27);
    const uint32_t* h = hi + (size_t)t * 36;
    ld<9>(acc.x, h); ld<9>(acc.y, h + 9); ld<9>(acc.t, h + 18); ld<9>(acc.z, h + 27);
    uint32_t* out = rows + ((size_t)j * fb_entries(w) + ((size_t)m << 8)) * TOM_PROJ_WORDS;
    for (int d = 0; d < 256; d++) {
      uint32_t* o = out + (size_t)d * TOM_PROJ_WORDS;
      tom_st_xyz(o, acc.x, acc.y, acc.z);
      tom_add(acc, acc, p);
    }
  }
};
// Rows of a fixed-base table (E1 projective X:Y:Z) -> entries of the prover's a = -1 image curve
// E2 (zk_curves.cuh): (w, v) = (sqrt(-d1) X/Z, Z/Y), stored as (v - w, v + w, 2 d2 w v), canonical,
// one 128-byte line each.  Montgomery's trick over chunks of 16 entries on the products Y*Z.
struct TomTabE2Task {
  const uint32_t* proj;  // [count][27]
  uint32_t* pre;         // [count][32]
  int count;
  ZK_HD void operator()(int t) const {
    using F = Tomp;
    constexpr int CH = 16;
    const int lo = t * CH;
    int n = count - lo;
    if (n > CH) n = CH;
    if (n <= 0) return;
    uint32_t pf[CH][9];
    uint32_t acc[9], y[9], z[9], den[9];
    F::set_one(acc);
    for (int k = 0; k < n; k++) {
      const uint32_t* src = proj + (size_t)(lo + k) * TOM_PROJ_WORDS;
      uint32_t xx[9];
      tom_ld_xyz(xx, y, z, src);
      F::mul(den, y, z);
      F::mul(acc, acc, den);
      copy_n<9>(pf[k], acc);
    }
    uint32_t inv[9], s2[9], dd2[9];
    F::inv(inv, acc);
    tom_const(s2, TOM_SQRTND1);
    tom_const(dd2, TOM_2D2);
    for (int k = n - 1; k >= 0; k--) {
      const uint32_t* src = proj + (size_t)(lo + k) * TOM_PROJ_WORDS;
      uint32_t x[9], di[9], w[9], v[9], kk[9], ym[9], yp[9];
      tom_ld_xyz(x, y, z, src);
      F::mul(den, y, z);
      if (k > 0) F::mul(di, inv, pf[k - 1]); else copy_n<9>(di, inv);   // 1 / (Y Z)
      F::mul(inv, inv, den);
      F::mul(w, x, y);      // X Y
      F::mul(w, w, di);     // X / Z
      F::mul(w, w, s2);     // w
      F::sqr(v, z);
      F::mul(v, v, di);     // Z / Y
      F::mul(kk, w, v);
      F::mul(kk, kk, dd2);
      F::sub(ym, v, w);
      F::add(yp, v, w);
      F::reduce(ym); F::reduce(yp); F::reduce(kk);
      uint32_t* o = pre + (size_t)(lo + k) * TOM_PRE_WORDS;
      st<9>(o, ym); st<9>(o + 9, yp); st<9>(o + 18, kk);
      for (int i = 27; i < 32; i++) o[i] = 0;
    }
  }
};

// Batched normalisation of tomEdwards256 points -> E1 affine (x', y) Montgomery (+ optional 67-byte
// reference encoding of (x = x'/sqrt(a), y)).  e2 == 0: input is E1 projective (X:Y:Z);
// e2 == 1: input is an E2 point (W:V:Z) from the commitment kernel: x' = W / (Z sqrt(-d1)), y = Z / V.
struct TomNormTask {
  const uint32_t* proj;  // [count][27]
  uint32_t* aff;         // [count][18] or null
  uint8_t* bytes;        // [count][BSTRIDE] or null
  int count;
  int chunk;             // points per thread (<= NORM_CHUNK_MAX)
  int e2;
  int aff_mod, aff_lim;  // the E1 affine pair is produced only for points with (index % aff_mod) < aff_lim
  ZK_HD static void canon2p(uint32_t* r) {   // value < 4p -> [0, p)
    uint32_t t[9], p2[9];
#pragma unroll
    for (int i = 0; i < 9; i++) p2[i] = (FpTom::p(i) << 1) | (i > 0 ? (FpTom::p(i - 1) >> 31) : 0u);
    uint32_t br = sub_n<9>(t, r, p2);
    csel_n<9>(r, br == 0, t, r);
    br = sub_p<FpTom>(t, r);
    csel_n<9>(r, br == 0, t, r);
  }
  ZK_HD void operator()(int t) const {
    using F = Tomp;
    const int lo = t * chunk;
    int n = count - lo;
    if (n > chunk) n = chunk;
    if (n <= 0) return;
    uint32_t pre[NORM_CHUNK_MAX][9];
    uint32_t acc[9], z[9], v[9], den[9];
    F::set_one(acc);
    for (int k = 0; k < n; k++) {
      const uint32_t* src = proj + (size_t)(lo + k) * TOM_PROJ_WORDS;
      uint32_t xx[9];
      tom_ld_xyz(xx, v, z, src);
      if (e2) F::mul(den, z, v); else copy_n<9>(den, z);
      F::mul(acc, acc, den);   // Z != 0 (complete curve); V != 0 inside the prime-order subgroup
      copy_n<9>(pre[k], acc);
    }
    uint32_t inv[9], isa[9], is2[9], isd[9], one_plain[9];
    F::inv(inv, acc);
    tom_const(isa, TOM_INVSQRTA);
    tom_const(is2, TOM_INVSQRTND1);
    tom_const(isd, TOM_INVSQRTND);
    zero_n<9>(one_plain);
    one_plain[0] = 1;
    for (int k = n - 1; k >= 0; k--) {
      const uint32_t* src = proj + (size_t)(lo + k) * TOM_PROJ_WORDS;
      uint32_t X[9], Y[9], di[9];
      tom_ld_xyz(X, Y, z, src);
      if (e2) F::mul(den, z, Y); else copy_n<9>(den, z);
      if (k > 0) F::mul(di, inv, pre[k - 1]); else copy_n<9>(di, inv);   // Montgomery residue of 1/den
      F::mul(inv, inv, den);
      if (bytes) {
        // D = 1/den as a PLAIN integer: a Montgomery product with it leaves Montgomery form, so the
        // reference coordinates come out without separate from_mont multiplications
        uint32_t Dp[9], cx[9], cy[9];
        F::mul(Dp, di, one_plain);
        if (e2) {                       // x = W V D / sqrt(-d),  y = Z^2 D
          F::mul(cx, X, Y);
          F::mul(cx, cx, isd);
          F::mul(cx, cx, Dp);
          F::sqr(cy, z);
          F::mul(cy, cy, Dp);
        } else {                        // x = X D / sqrt(a),  y = Y D
          F::mul(cx, X, isa);
          F::mul(cx, cx, Dp);
          F::mul(cy, Y, Dp);
        }
        canon2p(cx);
        canon2p(cy);
        store_point_words<9, 33>(bytes + (size_t)(lo + k) * BSTRIDE, 0x04u, cx, cy);
      }
      if (aff && ((lo + k) % aff_mod) < aff_lim) {
        uint32_t x[9], y[9];
        if (e2) {
          uint32_t zi[9], vi[9];
          F::mul(zi, di, Y);       // 1/Z
          F::mul(vi, di, z);       // 1/V
          F::mul(x, X, zi);
          F::mul(x, x, is2);       // x' = W / (Z sqrt(-d1))
          F::mul(y, z, vi);        // y = Z / V
        } else {
          F::mul(x, X, di);
          F::mul(y, Y, di);
        }
        uint32_t* a = aff + (size_t)(lo + k) * TOM_AFF_WORDS;
        st<9>(a, x);
        st<9>(a + 9, y);
      }
    }
  }
};

// Pedersen commitment in the proof group:  C = v*g + r*h   (pedersen.ts:53-58, gk.ts:88-92),
// both bases fixed => two positional tables, 2*nwin mixed additions, no doublings.
struct TomCommitTask {
  const uint32_t* jv;    // [count][8] canonical value scalars (mod tom.order)
  const uint32_t* jr;    // [count][8] canonical blinders
  const uint32_t* gtab;  // [nwin][2^w][32]
  const uint32_t* htab;
  uint32_t* proj;        // [count][27]  E2 point (W:V:Z) -> TomNormTask{e2 = 1}
  int w, nwin;
  ZK_HD void operator()(int t) const {
    uint32_t v[8], r[8];
    ld<8>(v, jv + (size_t)t * 8);
    ld<8>(r, jr + (size_t)t * 8);
    TomPt acc;
    tom_set_identity(acc);
    const size_t ne = (size_t)fb_entries(w);
    uint32_t cv = 0, cr = 0;
    for (int j = 0; j < nwin; j++) {
      TomPre q;
      bool nv, nr;
      const uint32_t dv = signed_digit(v, j, w, cv, nv), dr = signed_digit(r, j, w, cr, nr);
      tom_ld_pre(q, gtab + ((size_t)j * ne + dv) * TOM_PRE_WORDS);
      tom2_pre_neg(q, nv);
      tom2_madd<true>(acc, acc, q);     // a = -1 image curve E2: 7M per lookup
      tom_ld_pre(q, htab + ((size_t)j * ne + dr) * TOM_PRE_WORDS);
      tom2_pre_neg(q, nr);
      tom2_madd<true>(acc, acc, q);
    }
    uint32_t* o = proj + (size_t)t * TOM_PROJ_WORDS;
    tom_st_xyz(o, acc.x, acc.y, acc.z);
  }
};

struct TomCommitGTask {   // one thread per (item, g-part): K = v*g as an extended E2 point
  const uint32_t* jv;     // [items*34][8]
  const uint32_t* gtab;
  uint32_t* ext;          // [items*28][36]
  int w, nwin;
  ZK_HD void operator()(int t) const {
    const int item = t / GJOBS_PER_ITEM, g = t % GJOBS_PER_ITEM;
    uint32_t v[8];
    ld<8>(v, jv + ((size_t)item * JOBS_PER_ITEM + item_job_of_gpart(g)) * 8);
    TomPt acc;
    tom_set_identity(acc);
    const size_t ne = (size_t)fb_entries(w);
    uint32_t carry = 0;
#pragma unroll 1
    for (int j = 0; j < nwin; j++) {
      TomPre q;
      bool neg;
      const uint32_t d = signed_digit(v, j, w, carry, neg);
      tom_ld_pre(q, gtab + ((size_t)j * ne + d) * TOM_PRE_WORDS);
      tom2_pre_neg(q, neg);
      tom2_madd<true, TompCommit>(acc, acc, q);
    }
    uint32_t* o = ext + (size_t)t * TOM_EXT_WORDS;
    st<9>(o, acc.x); st<9>(o + 9, acc.y); st<9>(o + 18, acc.t); st<9>(o + 27, acc.z);
  }
};
struct TomCommitHTask {   // one thread per job: C = K + r*h
  const uint32_t* jr;     // [items*34][8]
  const uint32_t* htab;
  const uint32_t* ext;    // [items*28][36]
  uint32_t* proj;         // [items*34][28]
  int w, nwin;
  ZK_HD void operator()(int t) const {
    const int item = t / JOBS_PER_ITEM, jb = t % JOBS_PER_ITEM;
    uint32_t r[8];
    ld<8>(r, jr + (size_t)t * 8);
    TomPt acc;
    const uint32_t* s = ext + ((size_t)item * GJOBS_PER_ITEM + item_gpart_of_job(jb)) * TOM_EXT_WORDS;
    ld<9>(acc.x, s); ld<9>(acc.y, s + 9); ld<9>(acc.t, s + 18); ld<9>(acc.z, s + 27);
    const size_t ne = (size_t)fb_entries(w);
    uint32_t carry = 0;
#pragma unroll 1
    for (int j = 0; j < nwin; j++) {
      TomPre q;
      bool neg;
      const uint32_t d = signed_digit(r, j, w, carry, neg);
      tom_ld_pre(q, htab + ((size_t)j * ne + d) * TOM_PRE_WORDS);
      tom2_pre_neg(q, neg);
      tom2_madd<true, TompCommit>(acc, acc, q);
    }
    tom_st_xyz(proj + (size_t)t * TOM_PROJ_WORDS, acc.x, acc.y, acc.z);
  }
};

#if !defined(ZKA_HOSTSIM) && defined(ZKA_COMMIT_MINBLOCKS)
template <> struct TaskMinBlocks<TomCommitHTask> { static constexpr int value = ZKA_COMMIT_MINBLOCKS; };
template <> struct TaskMinBlocks<TomCommitGTask> { static constexpr int value = ZKA_COMMIT_MINBLOCKS; };
template <> struct TaskMinBlocks<TomCommitTask> { static constexpr int value = ZKA_COMMIT_MINBLOCKS; };
#endif

#endif   // ZKA_PG_WAR256

// ============================================================================ Groth-Kohlweiss sums
// sum over the 2^k ring entries i of block `blk` of coef_i * prod_j (bit_j(i) ? fo[j] : fz[j]),
// coef_i = vw ? (vw - v_i) : v_i  (gk.ts:141-171 prover, gk.ts:239-250 verifier).  The product is walked
// incrementally over the binary counter of the low k bits (about 3 multiplications per entry); the high
// n-k bits are fixed by the block index.  k = n, blk = 0 is the whole ring in one call; large rings are
// cut into blocks of 2^GK_BLOCK_BITS entries, one thread each, and summed afterwards.
enum : int { GK_BLOCK_BITS = 10 };
ZK_HD int gk_block_bits(int n) { return n < GK_BLOCK_BITS ? n : GK_BLOCK_BITS; }
ZK_HD void gk_block_sum(uint32_t* acc, const uint32_t* ring_m, const uint32_t (*fz)[8], const uint32_t (*fo)[8], int n,
                        int k, uint32_t blk, const uint32_t* vw) {
  using F = Tomq;
  uint32_t P[21][8];
  F::set_one(P[n]);
  for (int j = n - 1; j >= k; j--) F::mul(P[j], P[j + 1], ((blk >> (j - k)) & 1u) ? fo[j] : fz[j]);
  for (int j = k - 1; j >= 0; j--) F::mul(P[j], P[j + 1], fz[j]);
  zero_n<8>(acc);
  const size_t base = (size_t)blk << k;
  const uint32_t cnt = 1u << k;
  for (uint32_t l = 0;;) {
    uint32_t vi[8], term[8];
    ld<8>(vi, ring_m + (base + l) * 8);
    if (vw) F::sub(vi, vw, vi);
    F::mul(term, vi, P[0]);
    F::add(acc, acc, term);
    l++;
    if (l == cnt) break;
    // lowest set bit of the new l: bits below it are 0, it is 1, bits above unchanged
    int tz = 0;
    while (!((l >> tz) & 1u)) tz++;
    F::mul(P[tz], P[tz + 1], fo[tz]);
    for (int j = tz - 1; j >= 0; j--) F::mul(P[j], P[j + 1], fz[j]);
  }
}

// ================================================================================== hashing
// Streams `npts` encoded points (given as (pointer, length) by a functor) through SHA-256.
// Encodings in 4-byte aligned slots are fetched one point ahead (17 independent word loads).
ZK_HD bool hash_pt_fetch(uint32_t* t, const uint8_t* p, int len) {
  if (((size_t)p & 3) || len < 64 || len > 68) return false;
  const uint32_t* q = reinterpret_cast<const uint32_t*>(p);
#pragma unroll
  for (int j = 0; j < 17; j++) t[j] = (4 * j < len) ? q[j] : 0u;
  return true;
}
template <class Src>
ZK_HD void hash_points80(uint32_t* c3, const Src& src, int npts) {
  Sha256 h;
  h.init();
  uint32_t nx[17], cu[17];
  int nlen = 0;
  const uint8_t* np = npts > 0 ? src(0, nlen) : nullptr;
  bool nfast = npts > 0 && hash_pt_fetch(nx, np, nlen);
  for (int i = 0; i < npts; i++) {
    const uint8_t* p = np;
    const int len = nlen;
    const bool fast = nfast;
#pragma unroll
    for (int j = 0; j < 17; j++) cu[j] = nx[j];
    if (i + 1 < npts) {
      np = src(i + 1, nlen);
      nfast = hash_pt_fetch(nx, np, nlen);
    }
    if (fast) h.feed17(cu, len);
    else h.update(p, len);
  }
  h.final80(c3);
}

// challenge (80 bit, c3) as a canonical 8-limb scalar
ZK_HD void challenge_to_limbs(uint32_t* r, const uint32_t* c3) {
  zero_n<8>(r);
  r[0] = c3[0]; r[1] = c3[1]; r[2] = c3[2];
}

// Store the first n bytes of a little-endian byte stream held in words w[0..NW) to an arbitrarily
// aligned destination: <= 3 head bytes, aligned 32-bit stores built with 64-bit funnel shifts,
// <= 3 tail bytes.  All register indices are static (the proof layout has odd field offsets, so
// byte-wise copies were 65/67 single-byte stores per point).
template <int NW>
ZK_HD void store_stream(uint8_t* dst, const uint32_t* w, int n) {
  const int head = (int)((4 - ((size_t)dst & 3)) & 3);
#pragma unroll
  for (int i = 0; i < 3; i++)
    if (i < head && i < n) dst[i] = (uint8_t)(w[0] >> (8 * i));
  uint32_t* dw = reinterpret_cast<uint32_t*>(dst + head);
  const int nw = (n - head) >> 2;
#pragma unroll
  for (int j = 0; j < NW; j++) {
    const uint64_t pair = ((uint64_t)(j + 1 < NW ? w[j + 1] : 0u) << 32) | w[j];
    const uint32_t v = (uint32_t)(pair >> (8 * head));
    if (j < nw) {
      dw[j] = v;
    } else if (j == nw) {
#pragma unroll
      for (int t = 0; t < 3; t++)
        if (head + 4 * nw + t < n) dst[head + 4 * nw + t] = (uint8_t)(v >> (8 * t));
    }
  }
}
// copy an encoded point (n = 65 or 67 bytes) from a 4-byte aligned BSTRIDE slot
ZK_HD void copy_point(uint8_t* dst, const uint8_t* src, int n) {
  const uint32_t* sw = reinterpret_cast<const uint32_t*>(src);
  uint32_t w[BSTRIDE / 4];
#pragma unroll
  for (int i = 0; i < BSTRIDE / 4; i++) w[i] = sw[i];
  store_stream<BSTRIDE / 4>(dst, w, n);
}
// write a canonical 8-limb scalar as a big-endian field of LEN bytes (32 or 33)
template <int LEN>
ZK_HD void put_scalar(uint8_t* o, const uint32_t* c) {
  uint32_t w[9];
#pragma unroll
  for (int j = 0; j < 9; j++) {
    uint32_t v = 0;
#pragma unroll
    for (int t = 0; t < 4; t++) {
      const int k = 4 * j + t;               // stream position
      const int pos = LEN - 1 - k;           // byte significance
      if (k < LEN && pos >= 0 && pos < 32) v |= ((c[pos >> 2] >> (8 * (pos & 3))) & 0xffu) << (8 * t);
    }
    w[j] = v;
  }
  store_stream<9>(o, w, LEN);
}

}  // namespace zk
