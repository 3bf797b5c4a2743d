// zk_verify.cuh — stage tasks of the batched verifier (verifySignatureList over B proofs).
//
// Reference call tree being replaced (per proof):
//   /root/reference/src/zkpAttestList.ts:147-184  verifySignatureList (secparam = 20 literal, :177)
//   /root/reference/src/proofGK/gk.ts:197-262     verifyMembership
//   /root/reference/src/exp/exp.ts:233-349        verifyExp  (+ generateIndices :95-109)
//   /root/reference/src/exp/pointAdd.ts:199-259   aggregatePointAdd
//   /root/reference/src/commit/mult.ts:148-175    aggregateMult      (5 relations)
//   /root/reference/src/commit/equality.ts:94-116 aggregateEquality  (2 relations)
//   /root/reference/src/curves/multimult.ts       Relation.drain (one fresh random scalar per
//                                                 relation), MultiMult.evaluate (Bos-Coster heap)
//
// The verifier's decision is `isIdentity()` of three linear combinations (GK, multiW, multiN).
// A linear combination does not depend on how it is evaluated, so the GPU
//   * folds every term on a FIXED base (g, h of both groups, R via its per-proof table, and the
//     recomputed commitments T1x = sx g + r1 h, T1y, C7, C9, C12, Cint which are known
//     combinations of g, h and proof points) into a handful of scalars per proof,
//   * evaluates the remaining variable points (<= 682 tomEdwards256 + 21 P-256 + 4n+1 GK points
//     per proof, 256-bit scalars) with a bucket (Pippenger) MSM, one thread per (proof, window),
// using exactly the reference's randomizers (tape order below), so decisions agree even on
// invalid proofs.
//
// Verifier tape (per proof; include/zkattest.h):
//   [0, 32(2n+1))            GK drains in call order: rel0_0, rel1_0, rel0_1, ... , relFinal  (mod tom.order)
//   [G, G+78)                generateIndices: byte i is rnd(80 - i) already rejection-filtered (< 80 - i)
//   [G+96, ...)              exp drains, packed in consumption order: for each sampled repetition
//                            bit 1: relA (mod p256.n), relTx, relTy (mod tom.order)            = 3 draws
//                            bit 0: relA (mod p256.n), then pi8(5) pi10(5) pi11(5) pix(2) pi13(5) piy(2) = 25 draws
#pragma once
#include "zk_prove.cuh"   // reduce_once, shared item layout

namespace zk {

enum : int {
  V_SAMPLES = 20,              // default number of sampled repetitions: the literal secparam of zkpAttestList.ts:177
  V_ENT_PER_SAMPLE = 34,       // variable tomEdwards256 points of one sampled 0-bit repetition
  V_SEG = 20,                  // sampled repetitions per MSM segment (one window thread walks <= V_ENT_SEG entries)
  V_ENT_SEG = V_SEG * V_ENT_PER_SAMPLE + 2,       // 682 (+ keyXcom, keyYcom ride with segment 0)
  V_IDX_PAD = 96,
  V_PART_WORDS = 7 * 8,        // per-sample partial sums: gW hW pkX pkY | sR shN sCom
  MSM_C = 6,                   // SIGNED 6-bit bucket windows (tom): digits in [-32, 31], 32 buckets,
  MSM_NWIN = 43,               // 43 windows cover the 258 bits of k + offset (msm_digit6)
  MSM_C_N = 4,                 // P-256 MSM: SIGNED 4-bit windows, digits in [-8, 7], 8 buckets,
  MSM_NWIN_N = 65,             // 65 windows cover the 257 bits of k + offset (msm_digit4)
  // control words of the chunk-wide aggregate check (zk_verify_agg.cuh)
  AGG_SKIP = 0,                // != 0: some proof of the chunk is not eligible, the aggregate kernels return at once
  AGG_TOM_PASS = 1,            // != 0: sum_b (GK_b + W_b) is the identity -> the per-proof tomEdwards256 MSMs are skipped
  AGG_NIST_PASS = 2,           // != 0: sum_b N_b is the identity -> the per-proof P-256 MSMs are skipped
  AGG_CTL_WORDS = 4,
};

// K = number of sampled repetitions (verifyExp's secparam, exp.ts:233-262): 20 in verifySignatureList
ZK_LAYOUT_FN size_t verify_tape_len(int n, int /*reps*/, int K = V_SAMPLES) {
  return (size_t)32 * (2 * n + 1) + V_IDX_PAD + (size_t)32 * 25 * K;
}

struct VerifyCtx {
  int B, S, N, n;
  int K;                       // sampled repetitions (<= S)
  int mode;                    // 0: verifySignatureList; 1: verifyExp alone (exp.ts:233, no GK block, Q given or absent);
                               // 2: verifyMembership alone (gk.ts:197)
  const uint8_t* q_ext;        // mode 1: [B][65] Q points (all-zero = identity) or null (no Q: T1 = g*z)
  int tom_w, tom_nwin;
  const uint8_t* msg_hash;     // [B][32]
  const uint8_t* proofs;       // [B][proof_stride]
  size_t proof_stride;
  const uint32_t* proof_len;   // [B]
  const uint8_t* tape;         // [B][tape_stride]
  size_t tape_stride;
  const uint32_t* ring_m;      // [2^n][8] Montgomery mod q
  const uint32_t* g_tab8;
  const uint32_t* h_tab8;
  int h_w;
  const uint32_t* tg_tab;
  const uint32_t* th_tab;
  const uint8_t* tg_bytes;
  // per proof
  uint32_t* rep_off;    // [B][S]
  uint32_t* gk_off;     // [B]
  uint32_t* tagbits;    // [B][3] tag of each repetition (bit i)
  uint32_t* chal;       // [B][3]
  uint8_t* gk_ok_len;   // [B] 1 if the GK length check passes (gk.ts:208-218)
  uint8_t* gk_tape_bad; // [B] or null: a GK draw was out of range — recorded here and folded into status[] by VReduceTask when
                        //     the GK chain runs beside the exp chain (same precedence as running it after), else written at once
  uint32_t* r_aff;      // [B][16]
  uint32_t* q_aff;      // [B][16]
  uint8_t* q_inf;       // [B]
  uint32_t* rpows;      // [B][RT_NWIN][24]
  uint32_t* rrows;      // [B][RT_NWIN][RT_ROW][24]
  uint32_t* rtab;       // [B][RT_NWIN][RT_ROW][16]
  uint32_t* samp_idx;   // [B][20] sampled repetition index
  uint32_t* samp_draw;  // [B][20] first exp draw of the sample (32-byte units from the exp area)
  // per sample (B*20)
  uint32_t* sp_T;       // [B*20][24] projective T or T1
  uint32_t* sp_T_aff;   // [B*20][16]
  uint8_t* sp_T_inf;    // [B*20]
  // per sample: 2 fixed-base jobs (T1x, T1y) and 5 derived points
  uint32_t *ta_jv, *ta_jr, *ta_proj, *ta_aff;   // [B*20*2]
  uint32_t *td_proj, *td_aff;                   // [B*20*5]
  uint8_t* td_bytes;
  uint32_t* item_chal;  // [B*20][6][3]
  // MSM entries
  uint32_t* ent_scalar; // [B][V_ENT_TOM][8]   canonical mod q
  uint32_t* ent_off;    // [B][V_ENT_TOM]      byte offset of the point inside the proof
  uint32_t* ent_pre;    // [B][V_ENT_TOM][32]  parsed TomPre
  uint32_t* ent_cnt;    // [B][20]             entries used by sample j (2 or 34)
  uint32_t* part;       // [B*20][56]          partial sums (Montgomery mod q / mod n)
  uint32_t* nent_scalar;// [B][21][8]          canonical mod n
  uint32_t* nent_aff;   // [B][21][16]
  uint8_t* nent_skip;   // [B][21]
  // GK
  uint32_t* gk_scalar;  // [B][4n+1][8]
  uint32_t* gk_part;    // [B][2^(n-k)][8] block sums of the ring polynomial (only when n > GK_BLOCK_BITS)
  uint32_t* gk_pre;     // [B][4n+1][32]
  // fixed-base parts: tom jobs [B][2] (0: GK, 1: W) and their points; P-256 fixed part
  uint32_t *fx_jv, *fx_jr, *fx_proj;   // [B*2]
  uint32_t* nfix;       // [B][24] projective sR*R + shN*h
  // MSM window sums and verdicts
  uint32_t* win_w;      // [B][MSM_NWIN][36]
  uint32_t* win_g;      // [B][MSM_NWIN][36]
  uint32_t* win_n;      // [B][MSM_NWIN_N][24]
  uint8_t* id_flags;    // [B][3]  gk, W, N identity
  const uint32_t* agg_ctl;  // aggregate verdicts of the chunk (AGG_*), or null
  // outputs
  uint8_t* ok;          // [B]
  int32_t* status;      // [B]

  ZK_HD int ent_tom() const { return K * V_ENT_PER_SAMPLE + 2; }   // variable tomEdwards256 points per proof (+ keyXcom, keyYcom)
  ZK_HD int ent_nist() const { return K + 1; }                     // A_j + comS1
  ZK_HD int segs() const { return (K + V_SEG - 1) / V_SEG; }
  ZK_HD const uint8_t* proof_of(int b) const { return proofs + (size_t)b * proof_stride; }
  ZK_HD const uint8_t* tape_of(int b) const { return tape + (size_t)b * tape_stride; }
  ZK_HD size_t gk_tape_bytes() const { return mode == 1 ? 0 : (size_t)32 * (2 * n + 1); }
  ZK_HD const uint8_t* exp_tape(int b) const { return tape_of(b) + gk_tape_bytes() + V_IDX_PAD; }
  ZK_HD size_t ta_pt(size_t sample, int j) const { return sample * 2 + j; }   // 0 T1x, 1 T1y
  ZK_HD size_t td_pt(size_t sample, int j) const { return sample * DERS_PER_ITEM + j; }
};

// ---- small helpers ------------------------------------------------------------------------
#if defined(ZKA_PG_WAR256)
// parse a war256 point encoding (SEC1 uncompressed, weier.ts:74-89: 0x04 tag, coordinates < p, on curve) -> affine
// Montgomery; returns validity.  (The identity has no 65-byte encoding in a proof slot: tag 0x00 is malformed here.)
ZK_HD bool tom_parse(uint32_t* xm, uint32_t* ym, const uint8_t* b) {
  uint32_t x[8], y[8];
  limbs_from_be<8>(x, b + 1, 32);
  limbs_from_be<8>(y, b + 33, 32);
  const bool ok = (b[0] == 0x04) && lt_p<FpWar>(x) && lt_p<FpWar>(y);
  reduce_once<FpWar>(x);
  reduce_once<FpWar>(y);
  Warp::to_mont(xm, x);
  Warp::to_mont(ym, y);
  return ok && war_on_curve(xm, ym);
}
#else
// parse a tomEdwards256 point encoding -> image-curve affine Montgomery; returns validity
// (edwards.ts:70-86: 0x04 tag, coordinates < p, on curve)
ZK_HD bool tom_parse(uint32_t* xm, uint32_t* ym, const uint8_t* b) {
  uint32_t x[9], y[9], sa[9];
  limbs_from_be<9>(x, b + 1, 33);
  limbs_from_be<9>(y, b + 34, 33);
  bool ok = (b[0] == 0x04) && lt_p<FpTom>(x) && lt_p<FpTom>(y);
  Tomp::to_mont(xm, x);
  Tomp::to_mont(ym, y);
  tom_const(sa, TOM_SQRTA);
  Tomp::mul(xm, xm, sa);
  return ok && tom_on_curve(xm, ym);
}
#endif
// P-256 point encoding -> affine Montgomery. identity (65 zero bytes) -> inf.
ZK_HD bool p256_parse(P256Aff& a, bool& inf, const uint8_t* b) {
  uint32_t x[8], y[8];
  limbs_from_be<8>(x, b + 1, 32);
  limbs_from_be<8>(y, b + 33, 32);
  inf = (b[0] == 0) && is_zero_n<8>(x) && is_zero_n<8>(y);
  if (inf) { p256_set_generator(a); return true; }
  reduce_once<FpP256>(x);
  reduce_once<FpP256>(y);
  P256p::to_mont(a.x, x);
  P256p::to_mont(a.y, y);
  return b[0] == 0x04 && p256_on_curve(a.x, a.y);
}
// scalar encodings (group.ts:62-66 deserializeScalar: value < order)
ZK_HD bool nscalar_parse(uint32_t* r, const uint8_t* b) {   // 32 bytes, mod p256.n
  limbs_from_be<8>(r, b, 32);
  return lt_p<FnP256>(r);
}
ZK_HD bool wscalar_parse(uint32_t* r, const uint8_t* b) {   // WS bytes (33 tomEdwards256 / 32 war256), mod the group order
  limbs_from_be<8>(r, b + (WS - 32), 32);
  return (WS == 32 || b[0] == 0) && lt_p<FpP256>(r);
}
ZK_HD bool vdraw(uint32_t* r, const uint8_t* p, bool nist) {
  limbs_from_be<8>(r, p, 32);
  return nist ? lt_p<FnP256>(r) : lt_p<FpP256>(r);
}

// ---------------------------------------------------------------------------------------------
// V1 — layout, statement (zkpAttestList.ts:153-164) and R/Q.  One thread per proof.
// ---------------------------------------------------------------------------------------------
struct VLayoutTask {
  VerifyCtx c;
  ZK_HD void operator()(int b) const {
    using Fn = P256n;
    using Fp = P256p;
    c.status[b] = ZKA_OK;
    c.ok[b] = 0;
    const uint8_t* pr = c.proof_of(b);
    const uint32_t len = c.proof_len[b];
    bool bad = len < HEAD_LEN || len > c.proof_stride;
    uint32_t off = HEAD_LEN, tg[3] = {0, 0, 0};
    for (int i = 0; i < c.S && !bad; i++) {
      if (off + 1 > len) { bad = true; break; }
      const uint8_t tag = pr[off];
      if (tag > 1) { bad = true; break; }
      c.rep_off[(size_t)b * c.S + i] = off;
      if (tag) tg[i >> 5] |= 1u << (i & 31);
      off += tag ? REP1_LEN : REP0_LEN;
      if (off > len) bad = true;
    }
    int ngk = 0;
    if (!bad && c.mode == 1) {           // verifyExp alone: the row ends with the last repetition
      if (off != len) bad = true;
      uint32_t z[8];
      zero_n<8>(z);                      // no GK instance: its fixed-base job is 0*g + 0*h
      st<8>(c.fx_jv + (size_t)b * 2 * 8, z);
      st<8>(c.fx_jr + (size_t)b * 2 * 8, z);
    } else if (!bad) {
      if (off + 1 > len) bad = true;
      else {
        ngk = pr[off];
        if (off + (uint32_t)gk_len(ngk) != len) bad = true;
      }
    }
    st<3>(c.tagbits + (size_t)b * 3, tg);
    c.gk_off[b] = off;
    c.gk_ok_len[b] = (!bad && (c.mode == 1 || ngk == c.n)) ? 1 : 0;
    if (bad) {
      ZK_SET_STATUS(c.status + b, ZKA_ERR_MALFORMED);
      // park the offsets on the header so later stages read in-bounds garbage
      for (int i = 0; i < c.S; i++) c.rep_off[(size_t)b * c.S + i] = 0;
      c.gk_off[b] = 0;
    }
    // R, rinv, z1, Q
    P256Aff R;
    bool rinf = false;
    bool okR = bad ? true : p256_parse(R, rinf, pr);
    if (bad) p256_set_generator(R);
    if (!okR) { ZK_SET_STATUS(c.status + b, ZKA_ERR_MALFORMED); p256_set_generator(R); }
    if (rinf) ZK_SET_STATUS(c.status + b, ZKA_ERR_R_INFINITY);   // zkpAttestList.ts:158-160
    p256_st_aff(c.r_aff + (size_t)b * 16, R);
    if (c.mode == 1) {   // Q is an input (or absent): no statement to derive it from
      P256Aff Qa;
      bool qinf = true, okq = true;
      if (c.q_ext) okq = p256_parse(Qa, qinf, c.q_ext + (size_t)b * NP);
      if (!okq) ZK_SET_STATUS(c.status + b, ZKA_ERR_MALFORMED);
      if (qinf || !okq) p256_set_generator(Qa);
      p256_st_aff(c.q_aff + (size_t)b * 16, Qa);
      c.q_inf[b] = (qinf || !okq) ? 1 : 0;
      return;
    }
    uint32_t z[8], rx[8], zm[8], rm[8], rinv[8], t[8], z1[8];
    limbs_from_be<8>(z, c.msg_hash + (size_t)b * 32, 32);
    reduce_once<FnP256>(z);
    Fp::from_mont(rx, R.x);          // coordR.x as an integer (:161), reduced mod n
    reduce_once<FnP256>(rx);
    Fn::to_mont(zm, z);
    Fn::to_mont(rm, rx);
    Fn::inv(rinv, rm);
    Fn::mul(t, rinv, zm);
    Fn::from_mont(z1, t);
    P256Pt Q;
    p256_set_identity(Q);
    p256_accum_fixed(Q, c.g_tab8, z1, 8);
    P256Aff Qa;
    const bool qinf = p256_is_identity(Q);
    if (qinf) {
      p256_set_generator(Qa);
    } else {
      uint32_t zi[8];
      Fp::inv(zi, Q.z);
      Fp::mul(Qa.x, Q.x, zi);
      Fp::mul(Qa.y, Q.y, zi);
    }
    p256_st_aff(c.q_aff + (size_t)b * 16, Qa);
    c.q_inf[b] = qinf ? 1 : 0;
  }
};

// ---------------------------------------------------------------------------------------------
// V2 — full deserialisation checks (what readJson/deserializePoint/deserializeScalar would
// reject: weier.ts:74-89, edwards.ts:70-86, group.ts:62-66).  One thread per (proof, slot):
// slot < S validates repetition `slot`, slot == S the header + GK block.
// ---------------------------------------------------------------------------------------------
struct VValidateTask {
  VerifyCtx c;
  ZK_HD bool wpts(const uint8_t* p, int k) const {
    bool ok = true;
    uint32_t x[PGL], y[PGL];
    for (int i = 0; i < k; i++) ok = tom_parse(x, y, p + (size_t)i * WP) && ok;
    return ok;
  }
  ZK_HD bool wscs(const uint8_t* p, int k) const {
    bool ok = true;
    uint32_t r[8];
    for (int i = 0; i < k; i++) ok = wscalar_parse(r, p + (size_t)i * WS) && ok;
    return ok;
  }
  // (one thread per (proof, slot, part-of-repetition) was measured: 2x SLOWER — more threads re-reading
  //  the same headers; kept at one thread per (proof, slot))
  ZK_HD void operator()(int t) const {
    const int S1 = c.S + 1;
    const int b = t / S1, slot = t % S1;
    if (c.status[b] == ZKA_ERR_MALFORMED) return;
    const uint8_t* pr = c.proof_of(b);
    bool ok = true;
    uint32_t r[8];
    P256Aff a;
    bool inf;
    if (slot == c.S) {
      ok = p256_parse(a, inf, pr + NP) && ok;             // comS1 (R is checked in VLayoutTask)
      ok = wpts(pr + 2 * NP, 2) && ok;                    // keyXcom keyYcom
      if (c.mode != 1) {
        const uint8_t* g = pr + c.gk_off[b];
        const int n = g[0];
        ok = wpts(g + 1, 4 * n) && ok;
        ok = wscs(g + 1 + (size_t)4 * n * WP, 3 * n + 1) && ok;
      }
    } else {
      const uint8_t* rep = pr + c.rep_off[(size_t)b * c.S + slot];
      ok = p256_parse(a, inf, rep + 1) && ok;
      ok = wpts(rep + 1 + NP, 2) && ok;
      const uint8_t* body = rep + REP_HEAD;
      ok = nscalar_parse(r, body) && ok;
      ok = nscalar_parse(r, body + NS) && ok;
      if (rep[0]) {
        ok = wscs(body + 2 * NS, 2) && ok;
      } else {
        const uint8_t* pa = body + 2 * NS;
        ok = wpts(pa, 4) && ok;
        for (int m = 0; m < 4; m++) {
          const uint8_t* mp = pa + 4 * WP + m * MULT_LEN;
          ok = wpts(mp, 6) && ok;
          ok = wscs(mp + 6 * WP, 7) && ok;
        }
        for (int e = 0; e < 2; e++) {
          const uint8_t* ep = pa + 4 * WP + 4 * MULT_LEN + e * EQ_LEN;
          ok = wpts(ep, 2) && ok;
          ok = wscs(ep + 2 * WP, 3) && ok;
        }
        ok = wscs(pa + PA_LEN, 2) && ok;
      }
    }
    if (!ok) ZK_SET_STATUS(c.status + b, ZKA_ERR_MALFORMED);
  }
};

// ---------------------------------------------------------------------------------------------
// V3 — exp challenge (exp.ts:253-259), generateIndices (exp.ts:95-109) and the packed draw
// offsets of the sampled repetitions.  One thread per proof.
// ---------------------------------------------------------------------------------------------
struct VChallengeTask {
  VerifyCtx c;
  ZK_HD void operator()(int b) const {
    const uint8_t* pr = c.proof_of(b);
    Sha256 h;
    h.init();
    h.update(pr + 2 * NP, 2 * WP);
    for (int i = 0; i < c.S; i++) h.update(pr + c.rep_off[(size_t)b * c.S + i] + 1, NP + 2 * WP);
    uint32_t c3[3];
    h.final80(c3);
    st<3>(c.chal + (size_t)b * 3, c3);
    // Knuth shuffle with the pre-filtered index bytes: j = rnd(limit - i) + i
    uint8_t perm[MAX_REPS];
    for (int i = 0; i < c.S; i++) perm[i] = (uint8_t)i;
    const uint8_t* ib = c.tape_of(b) + c.gk_tape_bytes();
    for (int i = 0; i < c.S - 2; i++) {
      uint32_t r = ib[i];
      if (r >= (uint32_t)(c.S - i)) { ZK_SET_STATUS(c.status + b, ZKA_ERR_TAPE_RANGE); r = 0; }
      const int j = (int)r + i;
      const uint8_t k = perm[i];
      perm[i] = perm[j];
      perm[j] = k;
    }
    uint32_t tg[3];
    ld<3>(tg, c.tagbits + (size_t)b * 3);
    uint32_t draw = 0;
    for (int j = 0; j < c.K; j++) {
      const int i = perm[j];
      const uint32_t bit = (c3[i >> 5] >> (i & 31)) & 1u;
      const uint32_t tag = (tg[i >> 5] >> (i & 31)) & 1u;
      if (bit != tag) ZK_SET_STATUS(c.status + b, ZKA_ERR_PARAMS_NOT_FOUND);   // exp.ts:269-271,301-303
      c.samp_idx[(size_t)b * c.K + j] = (uint32_t)i;
      c.samp_draw[(size_t)b * c.K + j] = draw;
      draw += bit ? 3 : 25;
    }
  }
};

// V4 — T = R*alpha (bit 1) or T1 = R*z + Q (bit 0) (exp.ts:272,305,317-319). Per (proof, j).
struct VSampleP256Task {
  VerifyCtx c;
  ZK_HD void operator()(int t) const {
    const int b = t / c.K;
    const int i = c.samp_idx[t];
    const uint8_t* rep = c.proof_of(b) + c.rep_off[(size_t)b * c.S + i];
    uint32_t s[8];
    limbs_from_be<8>(s, rep + REP_HEAD, 32);   // alpha or z: first scalar of the body
    reduce_once<FnP256>(s);
    P256Pt T;
    p256_set_identity(T);
    p256_accum_rtab(T, c.rtab + (size_t)b * RT_ENTRIES * P256_AFF_WORDS, s);
    if (!rep[0] && !c.q_inf[b]) {
      P256Aff Q;
      p256_ld_aff(Q, c.q_aff + (size_t)b * 16);
      p256_madd(T, T, Q);
    }
    p256_st_proj(c.sp_T + (size_t)t * P256_PROJ_WORDS, T);
  }
};

// V5 — recomputed commitments T1x = g*sx + h*r1, T1y = g*sy + h*r2 (exp.ts:326-329) as
// fixed-base jobs.  Per (proof, j); 1-bit samples get the zero job (unused).
struct VSampleJobsTask {
  VerifyCtx c;
  ZK_HD void operator()(int t) const {
    const int b = t / c.K;
    const int i = c.samp_idx[t];
    const uint8_t* rep = c.proof_of(b) + c.rep_off[(size_t)b * c.S + i];
    uint32_t v[8], r[8];
    for (int xy = 0; xy < 2; xy++) {
      zero_n<8>(v);
      zero_n<8>(r);
      if (!rep[0]) {
        P256p::from_mont(v, c.sp_T_aff + (size_t)t * 16 + 8 * xy);
        wscalar_parse(r, rep + REP_HEAD + 2 * NS + PA_LEN + xy * WS);
      }
      st<8>(c.ta_jv + c.ta_pt(t, xy) * 8, v);
      st<8>(c.ta_jr + c.ta_pt(t, xy) * 8, r);
    }
    if (c.sp_T_inf[t]) ZK_SET_STATUS(c.status + b, rep[0] ? ZKA_ERR_T_INFINITY : ZKA_ERR_T1_INFINITY);  // exp.ts:283,323
  }
};

// V6 — C7, C9, C12, Cint, Cint2 (pointAdd.ts:213-215,236,248) by point addition. Per (proof, j).
struct VDerivedTask {
  VerifyCtx c;
  ZK_HD void frombytes(TomPt& p, const uint8_t* b) const {
    uint32_t x[PGL], y[PGL];
    tom_parse(x, y, b);
    tom_from_affine(p, x, y);
  }
  ZK_HD void stp(size_t idx, const TomPt& p) const {
    uint32_t* o = c.td_proj + idx * TOM_PROJ_WORDS;
    tom_st_xyz(o, p.x, p.y, p.z);
  }
  ZK_HD void operator()(int t) const {
    const int b = t / c.K;
    const int i = c.samp_idx[t];
    const uint8_t* pr = c.proof_of(b);
    const uint8_t* rep = pr + c.rep_off[(size_t)b * c.S + i];
    TomPt pkX, pkY, Tx, Ty, T1x, T1y, n, r;
    frombytes(pkX, pr + 2 * NP);
    frombytes(pkY, pr + 2 * NP + WP);
    frombytes(Tx, rep + 1 + NP);
    frombytes(Ty, rep + 1 + NP + WP);
    uint32_t x[PGL], y[PGL];
    ld<PGL>(x, c.ta_aff + c.ta_pt(t, 0) * TOM_AFF_WORDS); ld<PGL>(y, c.ta_aff + c.ta_pt(t, 0) * TOM_AFF_WORDS + PGL);
    tom_from_affine(T1x, x, y);
    ld<PGL>(x, c.ta_aff + c.ta_pt(t, 1) * TOM_AFF_WORDS); ld<PGL>(y, c.ta_aff + c.ta_pt(t, 1) * TOM_AFF_WORDS + PGL);
    tom_from_affine(T1y, x, y);
    tom_neg(n, T1x); tom_add(r, pkX, n); stp(c.td_pt(t, DER_C7), r);
    tom_neg(n, T1y); tom_add(r, pkY, n); stp(c.td_pt(t, DER_C9), r);
    tom_neg(n, Tx);  tom_add(r, T1x, n); stp(c.td_pt(t, DER_C12), r);
    tom_add(r, Tx, T1x); tom_add(r, r, pkX); stp(c.td_pt(t, DER_CINTX), r);
    tom_add(r, T1y, Ty); stp(c.td_pt(t, DER_CINTY), r);
  }
};

// V7 — the six challenges of a sampled 0-bit repetition (mult.ts:156, equality.ts:101).
// One thread per (proof, j, h).
struct VItemHashTask {
  VerifyCtx c;
  ZK_HD void operator()(int t) const {
    const int sample = t / HASHES_PER_ITEM, h = t % HASHES_PER_ITEM;
    const int b = sample / c.K;
    const int i = c.samp_idx[sample];
    const uint8_t* rep = c.proof_of(b) + c.rep_off[(size_t)b * c.S + i];
    uint32_t c3[3] = {0, 0, 0};
    if (!rep[0]) {
      const uint8_t* pa = rep + REP_HEAD + 2 * NS;
      const uint8_t* C8 = pa, *C10 = pa + WP, *C11 = pa + 2 * WP, *C13 = pa + 3 * WP;
      const uint8_t* der = c.td_bytes + c.td_pt(sample, 0) * BSTRIDE;
      Sha256 s;
      s.init();
      if (h < 4) {
        const uint8_t *cx, *cy, *cz;
        if (h == 0)      { cx = der + DER_C7 * BSTRIDE; cy = C8; cz = c.tg_bytes; }
        else if (h == 1) { cx = C8; cy = der + DER_C9 * BSTRIDE; cz = C10; }
        else if (h == 2) { cx = C10; cy = C10; cz = C11; }
        else             { cx = C10; cy = der + DER_C12 * BSTRIDE; cz = C13; }
        s.update(cx, WP); s.update(cy, WP); s.update(cz, WP);
        s.update(pa + 4 * WP + h * MULT_LEN, 6 * WP);     // C4 Ax Ay Az A4_1 A4_2 are contiguous
      } else {
        const int e = h - 4;
        s.update(e == 0 ? C11 : C13, WP);
        s.update(der + (e == 0 ? DER_CINTX : DER_CINTY) * BSTRIDE, WP);
        s.update(pa + 4 * WP + 4 * MULT_LEN + e * EQ_LEN, 2 * WP);
      }
      s.final80(c3);
    }
    st<3>(c.item_chal + (size_t)t * 3, c3);
  }
};

// ---------------------------------------------------------------------------------------------
// V8 — relations of one sampled repetition folded into (variable-point scalars, fixed-base
// partial sums).  All arithmetic mod q = tom.order in Montgomery form.  Per (proof, j).
// ---------------------------------------------------------------------------------------------
struct VRelationsTask {
  VerifyCtx c;
  // entry e of sample j of proof b
  ZK_HD void ent(int b, int j, int e, const uint32_t* s_mont, uint32_t off) const {
    uint32_t v[8];
    Tomq::from_mont(v, s_mont);
    const size_t idx = (size_t)b * c.ent_tom() + (size_t)j * V_ENT_PER_SAMPLE + e;
    st<8>(c.ent_scalar + idx * 8, v);
    c.ent_off[idx] = off;
  }
  ZK_HD void operator()(int t) const {
    using F = Tomq;
    using Fn = P256n;
    const int b = t / c.K, j = t % c.K;
    const int i = c.samp_idx[t];
    const uint8_t* pr = c.proof_of(b);
    const uint32_t roff = c.rep_off[(size_t)b * c.S + i];
    const uint8_t* rep = pr + roff;
    const uint8_t* body = rep + REP_HEAD;
    const uint8_t* dr = c.exp_tape(b) + (size_t)32 * c.samp_draw[t];
    uint32_t* part = c.part + (size_t)t * V_PART_WORDS;
    uint32_t gW[8], hW[8], pX[8], pY[8], sR[8], sH[8], sC[8];
    zero_n<8>(gW); zero_n<8>(hW); zero_n<8>(pX); zero_n<8>(pY); zero_n<8>(sR); zero_n<8>(sH); zero_n<8>(sC);
    bool tape_ok = true;
    // --- multiN: relA (exp.ts:273-279 / 306-316)
    uint32_t rho[8], rm[8], s[8], sm[8], t0[8], t1[8];
    tape_ok = vdraw(rho, dr, true) && tape_ok;
    Fn::to_mont(rm, rho);
    nscalar_parse(s, body);            // alpha | z
    Fn::to_mont(sm, s);
    Fn::mul(sR, rm, sm);               // rho * alpha  (coefficient of R, T = alpha R)
    nscalar_parse(s, body + NS);       // beta1 | z2
    Fn::to_mont(sm, s);
    Fn::mul(sH, rm, sm);
    if (!rep[0]) copy_n<8>(sC, rm);    // + rho * comS1
    {
      uint32_t neg[8], z[8];
      zero_n<8>(z);
      Fn::sub(neg, z, rho);            // -rho mod n, canonical
      st<8>(c.nent_scalar + ((size_t)b * c.ent_nist() + j) * 8, neg);
    }
    // coordinates of T / T1 as proof-group scalars
    uint32_t sx[8], sy[8];
    ld<8>(sx, c.sp_T_aff + (size_t)t * 16);
    ld<8>(sy, c.sp_T_aff + (size_t)t * 16 + 8);
    const uint32_t offTx = roff + 1 + NP, offTy = offTx + WP;
    if (rep[0]) {
      // relTx, relTy (exp.ts:284-298): sx g + beta2 h - Tx ; sy g + beta3 h - Ty
      uint32_t b2[8], b3[8], neg[8], z[8];
      zero_n<8>(z);
      wscalar_parse(b2, body + 2 * NS);
      wscalar_parse(b3, body + 2 * NS + WS);
      tape_ok = vdraw(rho, dr + 32, false) && tape_ok;
      F::to_mont(rm, rho);
      F::mul(t0, rm, sx); F::add(gW, gW, t0);
      F::to_mont(t1, b2); F::mul(t0, rm, t1); F::add(hW, hW, t0);
      F::sub(neg, z, rm); ent(b, j, 0, neg, offTx);
      tape_ok = vdraw(rho, dr + 64, false) && tape_ok;
      F::to_mont(rm, rho);
      F::mul(t0, rm, sy); F::add(gW, gW, t0);
      F::to_mont(t1, b3); F::mul(t0, rm, t1); F::add(hW, hW, t0);
      F::sub(neg, z, rm); ent(b, j, 1, neg, offTy);
      c.ent_cnt[t] = 2;
    } else {
      const uint8_t* pa = body + 2 * NS;
      const uint32_t offPa = roff + REP_HEAD + 2 * NS;
      uint32_t r1[8], r2[8], r1m[8], r2m[8];
      wscalar_parse(r1, pa + PA_LEN);
      wscalar_parse(r2, pa + PA_LEN + WS);
      F::to_mont(r1m, r1);
      F::to_mont(r2m, r2);
      // accumulated coefficients of the primitive points
      uint32_t cTx[8], cTy[8], cC8[8], cC10[8], cC11[8], cC13[8];
      zero_n<8>(cTx); zero_n<8>(cTy); zero_n<8>(cC8); zero_n<8>(cC10); zero_n<8>(cC11); zero_n<8>(cC13);
      // helper lambdas are avoided (host/device portability): explicit code per target kind
      // kind: 0 C7, 1 C8, 2 g(C14), 3 C9, 4 C10, 5 C11, 6 C12, 7 C13, 8 CintX, 9 CintY
#define ZK_ADD_TERM(kind, coef)                                                              \
  do {                                                                                       \
    const int _k = (kind);                                                                   \
    if (_k == 1) F::add(cC8, cC8, coef);                                                     \
    else if (_k == 4) F::add(cC10, cC10, coef);                                              \
    else if (_k == 5) F::add(cC11, cC11, coef);                                              \
    else if (_k == 7) F::add(cC13, cC13, coef);                                              \
    else if (_k == 2) F::add(gW, gW, coef);                                                  \
    else if (_k == 0) { /* C7 = pkX - T1x */                                                 \
      F::add(pX, pX, coef); F::mul(t0, coef, sx); F::sub(gW, gW, t0); F::mul(t0, coef, r1m); F::sub(hW, hW, t0); \
    } else if (_k == 3) { /* C9 = pkY - T1y */                                               \
      F::add(pY, pY, coef); F::mul(t0, coef, sy); F::sub(gW, gW, t0); F::mul(t0, coef, r2m); F::sub(hW, hW, t0); \
    } else if (_k == 6) { /* C12 = T1x - Tx */                                               \
      F::mul(t0, coef, sx); F::add(gW, gW, t0); F::mul(t0, coef, r1m); F::add(hW, hW, t0); F::sub(cTx, cTx, coef); \
    } else if (_k == 8) { /* Cint = Tx + T1x + pkX */                                        \
      F::add(cTx, cTx, coef); F::add(pX, pX, coef); F::mul(t0, coef, sx); F::add(gW, gW, t0); F::mul(t0, coef, r1m); F::add(hW, hW, t0); \
    } else { /* Cint2 = T1y + Ty */                                                          \
      F::add(cTy, cTy, coef); F::mul(t0, coef, sy); F::add(gW, gW, t0); F::mul(t0, coef, r2m); F::add(hW, hW, t0); \
    }                                                                                        \
  } while (0)
      int d = 1;   // draw index inside the sample (0 was relA)
      // order of aggregation: pi8, pi10, pi11, pix, pi13, piy (pointAdd.ts:221-253)
      for (int step = 0; step < 6; step++) {
        const bool is_eq = (step == 3 || step == 5);
        uint32_t cc[8], cm[8], c3[3];
        if (!is_eq) {
          const int m = step < 3 ? step : 3;
          const int kx = m == 0 ? 0 : (m == 1 ? 1 : 4);      // Cx: C7, C8, C10, C10
          const int ky = m == 0 ? 1 : (m == 1 ? 3 : (m == 2 ? 4 : 6));   // Cy: C8, C9, C10, C12
          const int kz = m == 0 ? 2 : (m == 1 ? 4 : (m == 2 ? 5 : 7));   // Cz: g, C10, C11, C13
          ld<3>(c3, c.item_chal + ((size_t)t * HASHES_PER_ITEM + m) * 3);
          challenge_to_limbs(cc, c3);
          F::to_mont(cm, cc);
          const uint8_t* mp = pa + 4 * WP + m * MULT_LEN;
          const uint32_t offM = offPa + 4 * WP + m * MULT_LEN;
          uint32_t ts[7][8];
          for (int q = 0; q < 7; q++) { wscalar_parse(ts[q], mp + 6 * WP + q * WS); F::to_mont(ts[q], ts[q]); }
          // ts: t_x t_y t_z t_rx t_ry t_rz t_r4
          uint32_t rr[5][8];
          for (int q = 0; q < 5; q++) { tape_ok = vdraw(rho, dr + 32 * (d + q), false) && tape_ok; F::to_mont(rr[q], rho); }
          d += 5;
          uint32_t coef[8], neg[8], z[8];
          zero_n<8>(z);
          // rho1: t_x g + t_rx h + c Cx - A_x
          F::mul(t0, rr[0], ts[0]); F::add(gW, gW, t0);
          F::mul(t0, rr[0], ts[3]); F::add(hW, hW, t0);
          F::mul(coef, rr[0], cm); ZK_ADD_TERM(kx, coef);
          F::sub(neg, z, rr[0]); ent(b, j, 6 + 6 * m + 1, neg, offM + WP);
          // rho2: t_y g + t_ry h + c Cy - A_y
          F::mul(t0, rr[1], ts[1]); F::add(gW, gW, t0);
          F::mul(t0, rr[1], ts[4]); F::add(hW, hW, t0);
          F::mul(coef, rr[1], cm);
          // rho5: t_x Cy + c C_4 - A_4_2   (Cy coefficient joins rho2's)
          F::mul(t0, rr[4], ts[0]); F::add(coef, coef, t0);
          ZK_ADD_TERM(ky, coef);
          F::sub(neg, z, rr[1]); ent(b, j, 6 + 6 * m + 2, neg, offM + 2 * WP);
          // rho3: t_z g + t_rz h + c Cz - A_z
          F::mul(t0, rr[2], ts[2]); F::add(gW, gW, t0);
          F::mul(t0, rr[2], ts[5]); F::add(hW, hW, t0);
          F::mul(coef, rr[2], cm); ZK_ADD_TERM(kz, coef);
          F::sub(neg, z, rr[2]); ent(b, j, 6 + 6 * m + 3, neg, offM + 3 * WP);
          // rho4: t_z g + t_r4 h + c C_4 - A_4_1
          F::mul(t0, rr[3], ts[2]); F::add(gW, gW, t0);
          F::mul(t0, rr[3], ts[6]); F::add(hW, hW, t0);
          F::add(coef, rr[3], rr[4]); F::mul(coef, coef, cm);     // C_4: (rho4 + rho5) c
          ent(b, j, 6 + 6 * m + 0, coef, offM);
          F::sub(neg, z, rr[3]); ent(b, j, 6 + 6 * m + 4, neg, offM + 4 * WP);
          F::sub(neg, z, rr[4]); ent(b, j, 6 + 6 * m + 5, neg, offM + 5 * WP);
        } else {
          const int e = step == 3 ? 0 : 1;
          ld<3>(c3, c.item_chal + ((size_t)t * HASHES_PER_ITEM + 4 + e) * 3);
          challenge_to_limbs(cc, c3);
          F::to_mont(cm, cc);
          const uint8_t* ep = pa + 4 * WP + 4 * MULT_LEN + e * EQ_LEN;
          const uint32_t offE = offPa + 4 * WP + 4 * MULT_LEN + e * EQ_LEN;
          uint32_t tx[8], tr1[8], tr2[8], ra[8], rb[8], coef[8], neg[8], z[8];
          zero_n<8>(z);
          wscalar_parse(tx, ep + 2 * WP); F::to_mont(tx, tx);
          wscalar_parse(tr1, ep + 2 * WP + WS); F::to_mont(tr1, tr1);
          wscalar_parse(tr2, ep + 2 * WP + 2 * WS); F::to_mont(tr2, tr2);
          tape_ok = vdraw(rho, dr + 32 * d, false) && tape_ok; F::to_mont(ra, rho);
          tape_ok = vdraw(rho, dr + 32 * (d + 1), false) && tape_ok; F::to_mont(rb, rho);
          d += 2;
          F::add(t1, ra, rb); F::mul(t0, t1, tx); F::add(gW, gW, t0);
          F::mul(t0, ra, tr1); F::add(hW, hW, t0);
          F::mul(t0, rb, tr2); F::add(hW, hW, t0);
          F::mul(coef, ra, cm); ZK_ADD_TERM(e == 0 ? 5 : 7, coef);     // C1 = C11 | C13
          F::mul(coef, rb, cm); ZK_ADD_TERM(e == 0 ? 8 : 9, coef);     // C2 = Cint | Cint2
          F::sub(neg, z, ra); ent(b, j, 30 + 2 * e, neg, offE);
          F::sub(neg, z, rb); ent(b, j, 30 + 2 * e + 1, neg, offE + WP);
        }
      }
#undef ZK_ADD_TERM
      ent(b, j, 0, cTx, offTx);
      ent(b, j, 1, cTy, offTy);
      ent(b, j, 2, cC8, offPa);
      ent(b, j, 3, cC10, offPa + WP);
      ent(b, j, 4, cC11, offPa + 2 * WP);
      ent(b, j, 5, cC13, offPa + 3 * WP);
      c.ent_cnt[t] = V_ENT_PER_SAMPLE;
    }
    if (!tape_ok) ZK_SET_STATUS(c.status + b, ZKA_ERR_TAPE_RANGE);
    st<8>(part, gW); st<8>(part + 8, hW); st<8>(part + 16, pX); st<8>(part + 24, pY);
    st<8>(part + 32, sR); st<8>(part + 40, sH); st<8>(part + 48, sC);
    // multiN variable point A_i
    P256Aff A;
    bool inf;
    p256_parse(A, inf, rep + 1);
    p256_st_aff(c.nent_aff + ((size_t)b * c.ent_nist() + j) * 16, A);
    c.nent_skip[(size_t)b * c.ent_nist() + j] = inf ? 1 : 0;
  }
};

// V9 — per proof: fold the 20 partial sums, emit the fixed-base jobs and the keyXcom/keyYcom/
// comS1 entries; evaluate sR*R + shN*h_nist.  One thread per proof.
struct VReduceTask {
  VerifyCtx c;
  ZK_HD void operator()(int b) const {
    using F = Tomq;
    using Fn = P256n;
    if (c.gk_tape_bad && c.gk_tape_bad[b]) ZK_SET_STATUS(c.status + b, ZKA_ERR_TAPE_RANGE);   // after every exp-side status
    uint32_t gW[8], hW[8], pX[8], pY[8], sR[8], sH[8], sC[8];
    zero_n<8>(gW); zero_n<8>(hW); zero_n<8>(pX); zero_n<8>(pY); zero_n<8>(sR); zero_n<8>(sH); zero_n<8>(sC);
    for (int j = 0; j < c.K; j++) {
      const uint32_t* p = c.part + ((size_t)b * c.K + j) * V_PART_WORDS;
      uint32_t t[8];
      ld<8>(t, p); F::add(gW, gW, t);
      ld<8>(t, p + 8); F::add(hW, hW, t);
      ld<8>(t, p + 16); F::add(pX, pX, t);
      ld<8>(t, p + 24); F::add(pY, pY, t);
      ld<8>(t, p + 32); Fn::add(sR, sR, t);
      ld<8>(t, p + 40); Fn::add(sH, sH, t);
      ld<8>(t, p + 48); Fn::add(sC, sC, t);
    }
    uint32_t v[8];
    // fixed-base job 1 (W): gW g + hW h
    F::from_mont(v, gW); st<8>(c.fx_jv + ((size_t)b * 2 + 1) * 8, v);
    F::from_mont(v, hW); st<8>(c.fx_jr + ((size_t)b * 2 + 1) * 8, v);
    // keyXcom / keyYcom entries
    size_t idx = (size_t)b * c.ent_tom() + (size_t)c.K * V_ENT_PER_SAMPLE;
    F::from_mont(v, pX); st<8>(c.ent_scalar + idx * 8, v); c.ent_off[idx] = 2 * NP;
    F::from_mont(v, pY); st<8>(c.ent_scalar + (idx + 1) * 8, v); c.ent_off[idx + 1] = 2 * NP + WP;
    // multiN: comS1 entry and the fixed part
    Fn::from_mont(v, sC);
    st<8>(c.nent_scalar + ((size_t)b * c.ent_nist() + c.K) * 8, v);
    P256Aff cs;
    bool inf;
    p256_parse(cs, inf, c.proof_of(b) + NP);
    p256_st_aff(c.nent_aff + ((size_t)b * c.ent_nist() + c.K) * 16, cs);
    c.nent_skip[(size_t)b * c.ent_nist() + c.K] = inf ? 1 : 0;
    uint32_t kR[8], kH[8];
    Fn::from_mont(kR, sR);
    Fn::from_mont(kH, sH);
    P256Pt acc;
    p256_set_identity(acc);
    p256_accum_rtab(acc, c.rtab + (size_t)b * RT_ENTRIES * P256_AFF_WORDS, kR);
    p256_accum_fixed(acc, c.h_tab8, kH, c.h_w);
    p256_st_proj(c.nfix + (size_t)b * P256_PROJ_WORDS, acc);
  }
};

// V10 — parse the variable tomEdwards256 points of the MSMs into table-entry form.
struct VParseEntriesTask {
  const uint8_t* proofs;
  size_t proof_stride;
  const uint32_t* off;   // [count] byte offset inside the proof
  uint32_t* pre;         // [count][32]
  int per_proof;
  ZK_HD void operator()(int t) const {
    const int b = t / per_proof;
#if defined(ZKA_PG_WAR256)
    uint32_t x[8], y[8];
    tom_parse(x, y, proofs + (size_t)b * proof_stride + off[t]);
    uint32_t* o = pre + (size_t)t * TOM_PRE_WORDS;
    st<8>(o, x); st<8>(o + 8, y);
#else
    using F = Tomp;
    uint32_t x[9], y[9], k[9], d1[9];
    tom_parse(x, y, proofs + (size_t)b * proof_stride + off[t]);
    tom_const(d1, TOM_D1);
    F::mul(k, x, y);
    F::mul(k, k, d1);
    F::reduce(x); F::reduce(y); F::reduce(k);
    uint32_t* o = pre + (size_t)t * TOM_PRE_WORDS;
    st<9>(o, x); st<9>(o + 9, y); st<9>(o + 18, k);
#endif
  }
};

// V11a — large rings only (n > GK_BLOCK_BITS): the N*n multiplications of the ring polynomial, one
// thread per (proof, block of 2^GK_BLOCK_BITS ring entries).  Every thread re-derives the challenge
// x = H(cl, ca, cb, cd) and the f_j from the proof (cheap next to its 3 * 1024 multiplications).
struct VGkSumTask {
  VerifyCtx c;
  ZK_HD void operator()(int t) const {
    using F = Tomq;
    const int n = c.n, k = gk_block_bits(n);
    const int nblk = 1 << (n - k);
    const int b = t / nblk, blk = t % nblk;
    uint32_t acc[8];
    zero_n<8>(acc);
    if (c.gk_ok_len[b]) {
      const uint8_t* g = c.proof_of(b) + c.gk_off[b];
      const uint8_t* pts = g + 1;
      const uint8_t* fs = pts + (size_t)4 * n * WP;
      Sha256 h;
      h.init();
      h.update(pts, 4 * n * WP);
      uint32_t c3[3], xc[8], xm[8];
      h.final80(c3);
      challenge_to_limbs(xc, c3);
      F::to_mont(xm, xc);
      uint32_t fm[20][8], omf[20][8];
      for (int i = 0; i < n; i++) {
        uint32_t f[8];
        wscalar_parse(f, fs + (size_t)i * WS);
        F::to_mont(fm[i], f);
        F::sub(omf[i], xm, fm[i]);
      }
      gk_block_sum(acc, c.ring_m, omf, fm, n, k, (uint32_t)blk, nullptr);
    }
    st<8>(c.gk_part + (size_t)t * 8, acc);
  }
};

// ---------------------------------------------------------------------------------------------
// V11 — Groth-Kohlweiss relations (gk.ts:220-259).  One thread per proof.
// ---------------------------------------------------------------------------------------------
struct VGkTask {
  VerifyCtx c;
  ZK_HD void operator()(int b) const {
    using F = Tomq;
    const int n = c.n;
    if (c.gk_tape_bad) c.gk_tape_bad[b] = 0;
    if (!c.gk_ok_len[b]) {   // length check fails -> verifyMembership returns false before any draw
      for (int k = 0; k < 4 * n + 1; k++) { uint32_t z[8]; zero_n<8>(z); st<8>(c.gk_scalar + ((size_t)b * (4 * n + 1) + k) * 8, z); }
      uint32_t z[8]; zero_n<8>(z);
      st<8>(c.fx_jv + (size_t)b * 2 * 8, z); st<8>(c.fx_jr + (size_t)b * 2 * 8, z);
      return;
    }
    const uint8_t* g = c.proof_of(b) + c.gk_off[b];
    const uint8_t* pts = g + 1;
    const uint8_t* fs = pts + (size_t)4 * n * WP;
    const uint8_t* zas = fs + (size_t)n * WS;
    const uint8_t* zbs = zas + (size_t)n * WS;
    const uint8_t* zds = zbs + (size_t)n * WS;
    Sha256 h;
    h.init();
    h.update(pts, 4 * n * WP);
    uint32_t c3[3], xc[8], xm[8];
    h.final80(c3);
    challenge_to_limbs(xc, c3);
    F::to_mont(xm, xc);
    const uint8_t* dr = c.tape_of(b);
    uint32_t gS[8], hS[8], t0[8], t1[8], rho[8], r0[8], r1[8], z[8];
    zero_n<8>(gS); zero_n<8>(hS); zero_n<8>(z);
    uint32_t fm[20][8], omf[20][8];     // f_j and x - f_j (Montgomery)
    bool tape_ok = true;
    uint32_t* sc = c.gk_scalar + (size_t)b * (4 * n + 1) * 8;
    for (int i = 0; i < n; i++) {
      uint32_t f[8], za[8], zb[8];
      wscalar_parse(f, fs + (size_t)i * WS);  F::to_mont(fm[i], f);
      wscalar_parse(za, zas + (size_t)i * WS); F::to_mont(za, za);
      wscalar_parse(zb, zbs + (size_t)i * WS); F::to_mont(zb, zb);
      F::sub(omf[i], xm, fm[i]);
      tape_ok = vdraw(rho, dr + 32 * (2 * i), false) && tape_ok;     F::to_mont(r0, rho);
      tape_ok = vdraw(rho, dr + 32 * (2 * i + 1), false) && tape_ok; F::to_mont(r1, rho);
      // rel0: x cl + ca - f g - za h ; rel1: (x - f) cl + cb - zb h
      F::mul(t0, r0, xm); F::mul(t1, r1, omf[i]); F::add(t0, t0, t1);
      F::from_mont(t1, t0); st<8>(sc + (size_t)i * 8, t1);                 // cl_i
      F::from_mont(t1, r0); st<8>(sc + (size_t)(n + i) * 8, t1);           // ca_i
      F::from_mont(t1, r1); st<8>(sc + (size_t)(2 * n + i) * 8, t1);       // cb_i
      F::mul(t0, r0, fm[i]); F::sub(gS, gS, t0);
      F::mul(t0, r0, za); F::sub(hS, hS, t0);
      F::mul(t0, r1, zb); F::sub(hS, hS, t0);
    }
    // total = sum_i v_i prod_j (bit_j(i) ? f_j : x - f_j)   (gk.ts:239-250)
    uint32_t total[8];
    {
      const int k = gk_block_bits(n), nblk = 1 << (n - k);
      if (nblk == 1) {
        gk_block_sum(total, c.ring_m, omf, fm, n, k, 0u, nullptr);
      } else {               // block sums from VGkSumTask
        uint32_t v[8];
        zero_n<8>(total);
        for (int i = 0; i < nblk; i++) {
          ld<8>(v, c.gk_part + ((size_t)b * nblk + i) * 8);
          F::add(total, total, v);
        }
      }
    }
    // relFinal: sum_k -x^k cd_k + x^n com - total g - zd h
    uint32_t rf[8], xp[8], zd[8];
    tape_ok = vdraw(rho, dr + 32 * (2 * n), false) && tape_ok;
    F::to_mont(rf, rho);
    F::set_one(xp);
    for (int k = 0; k < n; k++) {
      F::mul(t0, rf, xp);
      F::sub(t0, z, t0);
      F::from_mont(t1, t0); st<8>(sc + (size_t)(3 * n + k) * 8, t1);       // cd_k
      F::mul(xp, xp, xm);
    }
    F::mul(t0, rf, xp);
    F::from_mont(t1, t0); st<8>(sc + (size_t)(4 * n) * 8, t1);             // com = keyXcom
    F::mul(t0, rf, total); F::sub(gS, gS, t0);
    wscalar_parse(zd, zds); F::to_mont(zd, zd);
    F::mul(t0, rf, zd); F::sub(hS, hS, t0);
    F::from_mont(t1, gS); st<8>(c.fx_jv + (size_t)b * 2 * 8, t1);
    F::from_mont(t1, hS); st<8>(c.fx_jr + (size_t)b * 2 * 8, t1);
    if (!tape_ok) {
      if (c.gk_tape_bad) c.gk_tape_bad[b] = 1;
      else ZK_SET_STATUS(c.status + b, ZKA_ERR_TAPE_RANGE);
    }
  }
};
struct VGkOffsetsTask {   // byte offsets of cl, ca, cb, cd, com for VParseEntriesTask
  VerifyCtx c;
  uint32_t* off;          // [B][4n+1]
  ZK_HD void operator()(int t) const {
    const int per = 4 * c.n + 1;
    const int b = t / per, k = t % per;
    off[t] = c.gk_ok_len[b] ? (k < 4 * c.n ? c.gk_off[b] + 1 + (uint32_t)k * WP : (uint32_t)(2 * NP)) : (uint32_t)(2 * NP);
  }
};

// ---------------------------------------------------------------------------------------------
// Bucket MSM over tomEdwards256 (replaces MultiMult.evaluate, multimult.ts:61-89).
// One thread per (instance, window): 2^c - 1 buckets in local memory, mixed additions into
// buckets, then the running-sum reduction.  Instances: GK (4n+1 entries) and multiW.
// ---------------------------------------------------------------------------------------------
struct alignas(16) U4 { uint32_t x, y, z, w; };
#if defined(ZKA_PG_WAR256)
enum : int { PG_EXT_WORDS = 24, PG_EXT_U4 = 6 };   // a parked projective point: X, Y, Z
ZK_HD void bk_load(TomPt& p, const U4* b) {
  uint32_t w[24];
#pragma unroll
  for (int i = 0; i < 6; i++) { const U4 u = b[i]; w[4 * i] = u.x; w[4 * i + 1] = u.y; w[4 * i + 2] = u.z; w[4 * i + 3] = u.w; }
#pragma unroll
  for (int i = 0; i < 8; i++) { p.x[i] = w[i]; p.y[i] = w[8 + i]; p.z[i] = w[16 + i]; }
}
ZK_HD void bk_store(U4* b, const TomPt& p) {
  uint32_t w[24];
#pragma unroll
  for (int i = 0; i < 8; i++) { w[i] = p.x[i]; w[8 + i] = p.y[i]; w[16 + i] = p.z[i]; }
#pragma unroll
  for (int i = 0; i < 6; i++) { U4 u; u.x = w[4 * i]; u.y = w[4 * i + 1]; u.z = w[4 * i + 2]; u.w = w[4 * i + 3]; b[i] = u; }
}
#else
enum : int { PG_EXT_WORDS = 36, PG_EXT_U4 = 9 };   // a parked extended point: X, Y, T, Z
ZK_HD void bk_load(TomPt& p, const U4* b) {
  uint32_t w[36];
#pragma unroll
  for (int i = 0; i < 9; i++) { const U4 u = b[i]; w[4 * i] = u.x; w[4 * i + 1] = u.y; w[4 * i + 2] = u.z; w[4 * i + 3] = u.w; }
#pragma unroll
  for (int i = 0; i < 9; i++) { p.x[i] = w[i]; p.y[i] = w[9 + i]; p.t[i] = w[18 + i]; p.z[i] = w[27 + i]; }
}
ZK_HD void bk_store(U4* b, const TomPt& p) {
  uint32_t w[36];
#pragma unroll
  for (int i = 0; i < 9; i++) { w[i] = p.x[i]; w[9 + i] = p.y[i]; w[18 + i] = p.t[i]; w[27 + i] = p.z[i]; }
#pragma unroll
  for (int i = 0; i < 9; i++) { U4 u; u.x = w[4 * i]; u.y = w[4 * i + 1]; u.z = w[4 * i + 2]; u.w = w[4 * i + 3]; b[i] = u; }
}
#endif
// Signed window digits without a carry chain: with OFFS = sum_j 32 * 64^j the unsigned 6-bit windows of
// k' = k + OFFS, minus 32, are digits d_j in [-32, 31] with sum_j d_j 64^j = k.  Returns |d_j| (the
// bucket, 0..32) and its sign.  Half the buckets of an unsigned 6-bit window, one window fewer per
// 6 bits than the 5-bit version: 43 x (n + 64) instead of 52 x (n + 62) point operations.
ZK_HD uint32_t msm_digit6(const uint32_t* k, int w, bool& neg) {
  constexpr uint32_t OFFS[9] = {0x20820820u, 0x08208208u, 0x82082082u, 0x20820820u, 0x08208208u,
                                0x82082082u, 0x20820820u, 0x08208208u, 0x00000002u};
  uint32_t kp[9];
  uint64_t c = 0;
#pragma unroll
  for (int i = 0; i < 9; i++) {
    c += (uint64_t)(i < 8 ? k[i] : 0u) + OFFS[i];
    kp[i] = (uint32_t)c;
    c >>= 32;
  }
  const int pos = w * MSM_C, wi = pos >> 5, sh = pos & 31;
  uint64_t v = 0;
#pragma unroll
  for (int i = 0; i < 9; i++) {     // static indexing: select the two limbs the window touches
    if (i == wi) v |= kp[i];
    if (i == wi + 1) v |= (uint64_t)kp[i] << 32;
  }
  const int d = (int)((uint32_t)(v >> sh) & 63u) - 32;
  neg = d < 0;
  return (uint32_t)(d < 0 ? -d : d);
}
struct MsmTomWindowTask {
  // Sorted-bucket Pippenger: a thread first counting-sorts the indices of its window's entries by
  // |digit| (2 bytes of local memory per entry), then walks the sorted entries in one flat loop,
  // summing each bucket in REGISTERS, and finally folds the bucket sums into the running sums
  //   run += S_d ; tot += run      =>   tot = sum_d d * S_d .
  // (The first version kept 31 extended points per thread in local memory and read-modify-wrote
  // one per addition: 4.5 KB/thread thrashed L1/L2 — ncu: 26 GB of DRAM reads per launch, 5 % L1
  // hit rate, 25 % multiplier-pipe utilisation.  profiles/ncu_msm_r1d_*.md)
  const uint32_t* scalar;   // [inst][stride][8]
  const uint32_t* pre;      // [inst][stride][32]
  const uint32_t* cnt;      // [inst][groups] entries used per group (or null: all `group_len` used)
  int stride, groups, group_len, tail;   // entries of an instance = groups*group_len, then `tail` always-used ones
  int seg_groups, segs;     // the groups are walked in `segs` segments of <= seg_groups groups (one thread per segment
                            // and window: <= V_ENT_SEG entries each); the tail rides with segment 0
  uint32_t* win;            // [inst][segs][MSM_NWIN][PG_EXT_WORDS]
  ZK_HD void operator()(int t) const {
    const int is = t / MSM_NWIN, w = t % MSM_NWIN;
    const int inst = is / segs, seg = is % segs;
    constexpr int NB = 33;          // bucket 0 (skipped) .. 32
    const uint32_t* sc = scalar + (size_t)inst * stride * 8;
    const uint32_t* pp = pre + (size_t)inst * stride * TOM_PRE_WORDS;
    const int g0 = seg * seg_groups;
    int ng = groups - g0;
    if (ng > seg_groups) ng = seg_groups;
    if (ng < 0) ng = 0;
    const int tl = seg == 0 ? tail : 0;
    const int gbase = g0 * group_len;           // first entry of this segment's groups
    const int tbase = groups * group_len;       // first tail entry
    const int nloc = ng * group_len;            // local indices [0, nloc) are group entries, [nloc, nloc + tl) the tail
    uint16_t order[V_ENT_SEG];     // sign << 15 | (bucket - 1) << 10 | LOCAL entry index, sorted by bucket
    uint16_t start[NB + 1];
    for (int d = 0; d <= NB; d++) start[d] = 0;
    // pass 1: histogram of buckets
    for (int gidx = 0; gidx <= ng; gidx++) {
      const int m = gidx < ng ? (cnt ? (int)cnt[(size_t)inst * groups + g0 + gidx] : group_len) : tl;
      const int base = gidx < ng ? gbase + gidx * group_len : tbase;
      for (int e = 0; e < m; e++) {
        bool neg;
        const uint32_t bk = msm_digit6(sc + (size_t)(base + e) * 8, w, neg);
        start[bk + 1]++;
      }
    }
    for (int d = 1; d <= NB; d++) start[d] = (uint16_t)(start[d] + start[d - 1]);
    uint16_t fillp[NB];
    for (int d = 0; d < NB; d++) fillp[d] = start[d];
    // pass 2: scatter
    for (int gidx = 0; gidx <= ng; gidx++) {
      const int m = gidx < ng ? (cnt ? (int)cnt[(size_t)inst * groups + g0 + gidx] : group_len) : tl;
      const int base = gidx < ng ? gbase + gidx * group_len : tbase;
      const int lbase = gidx < ng ? gidx * group_len : nloc;
      for (int e = 0; e < m; e++) {
        bool neg;
        const uint32_t bk = msm_digit6(sc + (size_t)(base + e) * 8, w, neg);
        const uint32_t enc = bk ? (((neg ? 1u : 0u) << 15) | ((bk - 1) << 10) | (uint32_t)(lbase + e)) : (uint32_t)(lbase + e);
        order[fillp[bk]++] = (uint16_t)enc;
      }
    }
    // pass 3: ONE flat loop over the entries with a non-zero digit (the trip count is the same for
    // every window of an instance up to a few entries, so the warp does not diverge); a finished
    // bucket sum is parked in local memory exactly once
    U4 S[NB][PG_EXT_U4];
    uint64_t present = 0;
    TomPt acc;
    tom_set_identity(acc);
    int curd = NB - 1;
    const int total = start[NB], first = start[1];
    for (int q = total - 1; q >= first; q--) {
      const uint32_t oe = order[q];
      const int d = (int)((oe >> 10) & 31u) + 1;
      if (d != curd) {
        bk_store(S[curd], acc);
        present |= 1ull << curd;
        tom_set_identity(acc);
        curd = d;
      }
      TomPre pt;
      const int loc = (int)(oe & 1023u);
      tom_ld_pre(pt, pp + (size_t)(loc < nloc ? gbase + loc : tbase + (loc - nloc)) * TOM_PRE_WORDS);
      if (oe & 0x8000u) pg_pre_neg(pt);   // negative digit
      tom_madd<true, TompMsm>(acc, acc, pt);
    }
    bk_store(S[curd], acc);
    present |= 1ull << curd;
    // pass 4: running sums  tot = sum_d d * S_d
    TomPt run, tot;
    tom_set_identity(run);
    tom_set_identity(tot);
    for (int d = NB - 1; d >= 1; d--) {
      if ((present >> d) & 1ull) {
        bk_load(acc, S[d]);
        tom_add(run, run, acc);
      }
      tom_add(tot, tot, run);
    }
    bk_store(reinterpret_cast<U4*>(win + (size_t)t * PG_EXT_WORDS), tot);
  }
};
// Horner over the windows + the fixed-base part; verdict = identity?  One thread per instance.
struct MsmTomCombineTask {
  const uint32_t* win;      // [inst][segs][MSM_NWIN][36]
  const uint32_t* fixed;    // fixed-base commitment of instance i at fixed[(i*fix_stride + fix_off)*27]
  uint8_t* flag;            // verdict of instance i at flag[i*3 + flag_off]
  int fix_stride, fix_off, flag_off;
  int segs = 1;
  ZK_HD void operator()(int inst) const {
    TomPt acc, wsum;
    tom_set_identity(acc);
    for (int w = MSM_NWIN - 1; w >= 0; w--) {
      for (int k = 0; k < MSM_C; k++) tom_dbl(acc, acc);
      for (int sg = 0; sg < segs; sg++) {
        bk_load(wsum, reinterpret_cast<const U4*>(win + (((size_t)inst * segs + sg) * MSM_NWIN + w) * PG_EXT_WORDS));
        tom_add(acc, acc, wsum);
      }
    }
    TomPt f;
    const uint32_t* fp = fixed + ((size_t)inst * fix_stride + fix_off) * TOM_PROJ_WORDS;
    tom_ld_xyz(f.x, f.y, f.z, fp);
#if !defined(ZKA_PG_WAR256)
    // The commitment kernel works on the a = -1 image curve E2 and stores (W : V : Z) with
    // x' = W / (Z sqrt(-d1)), y = Z / V.  Same point in E1 extended coordinates with Z' = Z V:
    //   X = c W V,  Y = Z^2,  T = X Y / Z' = c W Z,   c = 1/sqrt(-d1).
    pg_fixed_to_msm(f);
#endif
    tom_add(acc, acc, f);
    const bool id = pg_is_identity(acc);
    flag[(size_t)inst * 3 + flag_off] = id ? 1 : 0;
  }
};

// Bucket MSM over P-256 (multiN): 21 points, 4-bit windows.  One thread per (proof, window).
// signed 4-bit digits without a carry chain (see msm_digit6): windows of k + sum_{j<64} 8 * 16^j, minus 8
ZK_HD uint32_t msm_digit4(const uint32_t* k, int w, bool& neg) {
  uint32_t kp[9];
  uint64_t c = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    c += (uint64_t)k[i] + 0x88888888u;
    kp[i] = (uint32_t)c;
    c >>= 32;
  }
  kp[8] = (uint32_t)c;
  uint32_t word = 0;
#pragma unroll
  for (int i = 0; i < 9; i++)
    if (i == (w >> 3)) word = kp[i];
  const int d = (int)((word >> (4 * (w & 7))) & 15u) - (w < 64 ? 8 : 0);
  neg = d < 0;
  return (uint32_t)(d < 0 ? -d : d);
}
struct MsmP256WindowTask {
  const uint32_t* scalar;   // [B][nent][8]
  const uint32_t* aff;      // [B][nent][16]
  const uint8_t* skip;      // [B][nent]
  uint32_t* win;            // [B][MSM_NWIN_N][24]
  int nent = V_SAMPLES + 1; // entries per proof: K sampled A_j + comS1
  const uint32_t* ctl = nullptr;
  ZK_HD void operator()(int t) const {
    if (ctl && ctl[AGG_NIST_PASS]) return;
    const int b = t / MSM_NWIN_N, w = t % MSM_NWIN_N;
    P256Pt bucket[8];
    for (int d = 0; d < 8; d++) p256_set_identity(bucket[d]);
    for (int e = 0; e < nent; e++) {
      if (skip[(size_t)b * nent + e]) continue;
      uint32_t k[8];
      ld<8>(k, scalar + ((size_t)b * nent + e) * 8);
      bool neg;
      const uint32_t dgt = msm_digit4(k, w, neg);
      if (dgt) {
        P256Aff q;
        p256_ld_aff(q, aff + ((size_t)b * nent + e) * 16);
        if (neg) P256p::neg(q.y, q.y);
        p256_madd(bucket[dgt - 1], bucket[dgt - 1], q);
      }
    }
    P256Pt run, tot;
    p256_set_identity(run);
    p256_set_identity(tot);
    for (int d = 7; d >= 0; d--) {
      p256_add(run, run, bucket[d]);
      p256_add(tot, tot, run);
    }
    p256_st_proj(win + (size_t)t * P256_PROJ_WORDS, tot);
  }
};
struct MsmP256CombineTask {
  const uint32_t* win;
  const uint32_t* fixed;    // [B][24]
  uint8_t* flag;            // [B][3], writes [b][2]
  ZK_HD void operator()(int b) const {
    P256Pt acc, wsum;
    p256_set_identity(acc);
    for (int w = MSM_NWIN_N - 1; w >= 0; w--) {
      for (int k = 0; k < MSM_C_N; k++) p256_dbl(acc, acc);
      p256_ld_proj(wsum, win + ((size_t)b * MSM_NWIN_N + w) * P256_PROJ_WORDS);
      p256_add(acc, acc, wsum);
    }
    P256Pt f;
    p256_ld_proj(f, fixed + (size_t)b * P256_PROJ_WORDS);
    p256_add(acc, acc, f);
    flag[(size_t)b * 3 + 2] = p256_is_identity(acc) ? 1 : 0;
  }
};

// The two tomEdwards256 MSM instances of a proof (multiW: up to 682 entries, GK: 4n+1) as one grid: the
// short GK threads fill the SMs the long multiW threads leave idle (one 1024-proof batch is 0.8 of a wave).
struct MsmTomWindowBothTask {
  MsmTomWindowTask w, gk;
  int nW, nWp;   // multiW threads (proofs x segments x windows), rounded up to a warp multiple
  int nG;        // GK threads (proofs x windows)
  const uint32_t* ctl = nullptr;
  ZK_HD void operator()(int t) const {
    if (ctl && ctl[AGG_TOM_PASS]) return;
    if (t < nWp) {
      if (t < nW) w(t);
    } else if (t - nWp < nG) {
      gk(t - nWp);
    }
  }
};

#if !defined(ZKA_HOSTSIM) && defined(ZKA_MSM_MINBLOCKS)
template <> struct TaskMinBlocks<MsmTomWindowBothTask> { static constexpr int value = ZKA_MSM_MINBLOCKS; };
#endif

// The three Horner passes of a proof (GK, multiW, multiN) are 250-doubling latency chains run by one
// thread each; launched as ONE grid they overlap instead of queueing (3 B threads are still few).
struct MsmCombineAllTask {
  MsmTomCombineTask gk, w;
  MsmP256CombineTask n;
  int B, Bp;   // Bp = B rounded up to a warp multiple: the P-256 chain never shares a warp with a Tom chain
  const uint32_t* ctl = nullptr;
  ZK_HD void operator()(int t) const {
    const int kind = t / Bp, i = t % Bp;
    if (i >= B) return;
    if (ctl && ctl[kind == 2 ? AGG_NIST_PASS : AGG_TOM_PASS]) return;
    if (kind == 0) gk(i);
    else if (kind == 1) w(i);
    else n(i);
  }
};


// ---- stand-alone sub-proof verifiers (the reference's unit-test / bench surface) -------------------------------
// Rows for the shared tasks are assembled on the device: a 264-byte header followed by the caller's proof bytes.
//   verifyExp        (exp.ts:233):  header = paramsNIST.g | Clambda | Px | Py, body = the repetitions
//   verifyMembership (gk.ts:197):   header = 0 | 0 | com | 0,              body = the GK block
struct VAssembleTask {
  const uint8_t *h0, *h1, *h2, *h3;   // [B][65] [B][65] [B][67] [B][67]; null = zero bytes
  const uint8_t* body;                // [B][body_stride]
  size_t body_stride;
  const uint32_t* body_len;           // [B]
  uint8_t* rows;                      // [B][row_stride]
  size_t row_stride;
  uint32_t* row_len;                  // [B]  = HEAD_LEN + body_len (clamped to the stride)
  int pieces;                         // 64-byte pieces per row
  ZK_HD void operator()(int t) const {
    const int b = t / pieces, j = t % pieces;
    uint32_t bl = body_len[b];
    if ((size_t)bl > body_stride) bl = (uint32_t)body_stride + 1;   // stays "too long" -> MALFORMED
    if (j == 0) row_len[b] = HEAD_LEN + bl;
    uint8_t* row = rows + (size_t)b * row_stride;
    const size_t lo = (size_t)64 * j, hi = lo + 64;
    for (size_t o = lo; o < hi && o < row_stride; o++) {
      uint8_t v = 0;
      if (o < NP) v = h0 ? h0[(size_t)b * NP + o] : 0;
      else if (o < 2 * NP) v = h1 ? h1[(size_t)b * NP + (o - NP)] : 0;
      else if (o < 2 * NP + WP) v = h2 ? h2[(size_t)b * WP + (o - 2 * NP)] : 0;
      else if (o < HEAD_LEN) v = h3 ? h3[(size_t)b * WP + (o - 2 * NP - WP)] : 0;
      else if (o - HEAD_LEN < bl && o - HEAD_LEN < body_stride) v = body[(size_t)b * body_stride + (o - HEAD_LEN)];
      row[o] = v;
    }
  }
};
// verifyMembership alone: layout + deserialisation checks of com and the GK block.  One thread per proof.
struct VGkOnlyLayoutTask {
  VerifyCtx c;
  ZK_HD void operator()(int b) const {
    c.status[b] = ZKA_OK;
    c.ok[b] = 0;
    const uint8_t* pr = c.proof_of(b);
    const uint32_t len = c.proof_len[b];
    bool bad = len < HEAD_LEN + 1 || len > c.proof_stride;
    int ngk = 0;
    if (!bad) {
      ngk = pr[HEAD_LEN];
      if (HEAD_LEN + (uint32_t)gk_len(ngk) != len) bad = true;
    }
    c.gk_off[b] = bad ? 0 : HEAD_LEN;
    bool okp = true;
    if (!bad) {
      uint32_t x[PGL], y[PGL], r[8];
      okp = tom_parse(x, y, pr + 2 * NP);
      const uint8_t* g = pr + HEAD_LEN;
      for (int i = 0; i < 4 * ngk; i++) okp = tom_parse(x, y, g + 1 + (size_t)i * WP) && okp;
      for (int i = 0; i < 3 * ngk + 1; i++) okp = wscalar_parse(r, g + 1 + (size_t)4 * ngk * WP + (size_t)i * WS) && okp;
    }
    if (bad || !okp) ZK_SET_STATUS(c.status + b, ZKA_ERR_MALFORMED);
    c.gk_ok_len[b] = (!bad && okp && ngk == c.n) ? 1 : 0;
  }
};
struct VGkOnlyFinalTask {
  VerifyCtx c;
  ZK_HD void operator()(int b) const {
    const int st = c.status[b];
    c.ok[b] = (st == ZKA_OK && c.gk_ok_len[b] && c.id_flags[(size_t)b * 3]) ? 1 : 0;
  }
};

// ---- verifyEquality / verifyMult / verifyPointAdd alone (equality.ts:80-116, mult.ts:133-175, pointAdd.ts:181-259)
// Row = the statement's commitments (2 / 3 / 6 x 67 bytes, in the reference's argument order) followed by the proof
// (233 / 633 / 3266 bytes).  One thread per statement folds all relations under the tape's randomizers into
//   * scalars of the variable points (inputs and proof points)  -> entries of ONE Pippenger instance,
//   * the coefficients of g and h                                -> one fixed-base commitment,
// exactly like the batched verifier does for a sampled repetition; derived commitments (C7, C9, C12, Cint) are
// expanded onto the inputs they are sums of.
enum : int { SUB_EQ = 0, SUB_MULT = 1, SUB_PADD = 2, SUB_ENT_MAX = 38 };
ZK_LAYOUT_FN int sub_points(int kind) { return kind == SUB_EQ ? 2 : kind == SUB_MULT ? 3 : 6; }
ZK_LAYOUT_FN int sub_proof_len(int kind) { return kind == SUB_EQ ? EQ_LEN : kind == SUB_MULT ? MULT_LEN : PA_LEN; }
ZK_LAYOUT_FN int sub_draws(int kind) { return kind == SUB_EQ ? 2 : kind == SUB_MULT ? 5 : 24; }
ZK_LAYOUT_FN int sub_entries(int kind) { return kind == SUB_EQ ? 4 : kind == SUB_MULT ? 9 : SUB_ENT_MAX; }

// encoding (67 bytes in a 68-byte slot) of an E1 affine point given as Montgomery (x', y)
#if defined(ZKA_PG_WAR256)
ZK_HD void tom_encode_affine(uint8_t* out, const uint32_t* xm, const uint32_t* ym) {
  uint32_t cx[8], cy[8];
  Warp::from_mont(cx, xm);
  Warp::from_mont(cy, ym);
  store_point_words<8, 32>(out, 0x04u, cx, cy);
}
#else
ZK_HD void tom_encode_affine(uint8_t* out, const uint32_t* x1m, const uint32_t* ym) {
  uint32_t isa[9], cx[9], cy[9];
  tom_const(isa, TOM_INVSQRTA);
  Tomp::mul(cx, x1m, isa);
  Tomp::from_mont(cx, cx);
  Tomp::from_mont(cy, ym);
  store_point_words<9, 33>(out, 0x04u, cx, cy);
}
#endif
ZK_HD void tom_encode_proj(uint8_t* out, const TomPt& p) {   // one inversion: low-volume paths only
  uint32_t zi[PGL], x[PGL], y[PGL];
  PGp::inv(zi, p.z);
  PGp::mul(x, p.x, zi);
  PGp::mul(y, p.y, zi);
  tom_encode_affine(out, x, y);
}
struct VSubProofTask {
  int kind;
  const uint8_t* rows;     // [B][stride]
  size_t stride;
  const uint8_t* tape;     // [B][tape_stride] drains (mod tom.order) in Relation.drain call order
  size_t tape_stride;
  const uint8_t* tg_bytes; // encoding of ProofGroup.g (C_14 of pi_8, pointAdd.ts:220)
  uint32_t* ent_scalar;    // [B][SUB_ENT_MAX][8] canonical
  uint32_t* ent_off;       // [B][SUB_ENT_MAX] byte offset of the point in the row
  uint32_t *fx_jv, *fx_jr; // [B][2][8]: job 1 = (coefficient of g, coefficient of h); job 0 = 0
  int32_t* status;         // [B]
  uint8_t* ok;             // [B]

  struct Fold {            // running state of one statement
    uint32_t gW[8], hW[8];
    bool tape_ok;
  };
  ZK_HD static void acc_add(uint32_t* a, const uint32_t* v) { Tomq::add(a, a, v); }
  // aggregateMult (mult.ts:148-175).  bx/by/bz: 67-byte encodings; mp: MultProof bytes; dr: 5 drains.
  // cx/cy/cz: Montgomery coefficient accumulators of Cx, Cy, Cz; es: canonical scalars of C4 Ax Ay Az A41 A42.
  ZK_HD static void mult(Fold& f, const uint8_t* bx, const uint8_t* by, const uint8_t* bz, const uint8_t* mp, const uint8_t* dr,
                         uint32_t* cx, uint32_t* cy, uint32_t* cz, uint32_t (*es)[8]) {
    using F = Tomq;
    Sha256 h;
    h.init();
    h.update(bx, WP); h.update(by, WP); h.update(bz, WP); h.update(mp, 6 * WP);
    uint32_t c3[3], cc[8], cm[8];
    h.final80(c3);
    challenge_to_limbs(cc, c3);
    F::to_mont(cm, cc);
    uint32_t ts[7][8], rr[5][8], rho[8], t0[8], coef[8], neg[8], z[8];
    zero_n<8>(z);
    for (int q = 0; q < 7; q++) { wscalar_parse(ts[q], mp + 6 * WP + q * WS); F::to_mont(ts[q], ts[q]); }
    for (int q = 0; q < 5; q++) { f.tape_ok = vdraw(rho, dr + 32 * q, false) && f.tape_ok; F::to_mont(rr[q], rho); }
    // rho1: t_x g + t_rx h + c Cx - A_x
    F::mul(t0, rr[0], ts[0]); acc_add(f.gW, t0);
    F::mul(t0, rr[0], ts[3]); acc_add(f.hW, t0);
    F::mul(coef, rr[0], cm); acc_add(cx, coef);
    F::sub(neg, z, rr[0]); F::from_mont(es[1], neg);
    // rho2: t_y g + t_ry h + c Cy - A_y ; rho5: t_x Cy + c C_4 - A_4_2
    F::mul(t0, rr[1], ts[1]); acc_add(f.gW, t0);
    F::mul(t0, rr[1], ts[4]); acc_add(f.hW, t0);
    F::mul(coef, rr[1], cm);
    F::mul(t0, rr[4], ts[0]); F::add(coef, coef, t0);
    acc_add(cy, coef);
    F::sub(neg, z, rr[1]); F::from_mont(es[2], neg);
    // rho3: t_z g + t_rz h + c Cz - A_z
    F::mul(t0, rr[2], ts[2]); acc_add(f.gW, t0);
    F::mul(t0, rr[2], ts[5]); acc_add(f.hW, t0);
    F::mul(coef, rr[2], cm); acc_add(cz, coef);
    F::sub(neg, z, rr[2]); F::from_mont(es[3], neg);
    // rho4: t_z g + t_r4 h + c C_4 - A_4_1
    F::mul(t0, rr[3], ts[2]); acc_add(f.gW, t0);
    F::mul(t0, rr[3], ts[6]); acc_add(f.hW, t0);
    F::add(coef, rr[3], rr[4]); F::mul(coef, coef, cm); F::from_mont(es[0], coef);   // C_4: (rho4 + rho5) c
    F::sub(neg, z, rr[3]); F::from_mont(es[4], neg);
    F::sub(neg, z, rr[4]); F::from_mont(es[5], neg);
  }
  // aggregateEquality (equality.ts:94-116): es = canonical scalars of A1, A2
  ZK_HD static void equality(Fold& f, const uint8_t* b1, const uint8_t* b2, const uint8_t* ep, const uint8_t* dr, uint32_t* c1,
                             uint32_t* c2, uint32_t (*es)[8]) {
    using F = Tomq;
    Sha256 h;
    h.init();
    h.update(b1, WP); h.update(b2, WP); h.update(ep, 2 * WP);
    uint32_t c3[3], cc[8], cm[8];
    h.final80(c3);
    challenge_to_limbs(cc, c3);
    F::to_mont(cm, cc);
    uint32_t tx[8], tr1[8], tr2[8], ra[8], rb[8], rho[8], t0[8], t1[8], coef[8], neg[8], z[8];
    zero_n<8>(z);
    wscalar_parse(tx, ep + 2 * WP); F::to_mont(tx, tx);
    wscalar_parse(tr1, ep + 2 * WP + WS); F::to_mont(tr1, tr1);
    wscalar_parse(tr2, ep + 2 * WP + 2 * WS); F::to_mont(tr2, tr2);
    f.tape_ok = vdraw(rho, dr, false) && f.tape_ok; F::to_mont(ra, rho);
    f.tape_ok = vdraw(rho, dr + 32, false) && f.tape_ok; F::to_mont(rb, rho);
    F::add(t1, ra, rb); F::mul(t0, t1, tx); acc_add(f.gW, t0);
    F::mul(t0, ra, tr1); acc_add(f.hW, t0);
    F::mul(t0, rb, tr2); acc_add(f.hW, t0);
    F::mul(coef, ra, cm); acc_add(c1, coef);
    F::mul(coef, rb, cm); acc_add(c2, coef);
    F::sub(neg, z, ra); F::from_mont(es[0], neg);
    F::sub(neg, z, rb); F::from_mont(es[1], neg);
  }
  ZK_HD void emit(int b, int e, const uint32_t* canon, uint32_t off) const {
    st<8>(ent_scalar + ((size_t)b * SUB_ENT_MAX + e) * 8, canon);
    ent_off[(size_t)b * SUB_ENT_MAX + e] = off;
  }
  ZK_HD void emit_m(int b, int e, const uint32_t* mont, uint32_t off) const {
    uint32_t v[8];
    Tomq::from_mont(v, mont);
    emit(b, e, v, off);
  }
  ZK_HD void operator()(int b) const {
    using F = Tomq;
    const uint8_t* row = rows + (size_t)b * stride;
    const uint8_t* dr = tape + (size_t)b * tape_stride;
    const int np = sub_points(kind);
    const uint8_t* proof = row + (size_t)np * WP;
    const uint32_t poff = (uint32_t)(np * WP);
    status[b] = ZKA_OK;
    ok[b] = 0;
    // deserialisation checks of every point and scalar (deserializePoint / deserializeScalar would throw)
    bool good = true;
    {
      uint32_t x[PGL], y[PGL], r[8];
      for (int i = 0; i < np; i++) good = tom_parse(x, y, row + (size_t)i * WP) && good;
      if (kind == SUB_EQ) {
        for (int i = 0; i < 2; i++) good = tom_parse(x, y, proof + (size_t)i * WP) && good;
        for (int i = 0; i < 3; i++) good = wscalar_parse(r, proof + 2 * WP + (size_t)i * WS) && good;
      } else {
        const int nm = kind == SUB_MULT ? 1 : 4;
        const uint8_t* m0 = kind == SUB_MULT ? proof : proof + 4 * WP;
        if (kind == SUB_PADD) for (int i = 0; i < 4; i++) good = tom_parse(x, y, proof + (size_t)i * WP) && good;
        for (int m = 0; m < nm; m++) {
          for (int i = 0; i < 6; i++) good = tom_parse(x, y, m0 + (size_t)m * MULT_LEN + (size_t)i * WP) && good;
          for (int i = 0; i < 7; i++) good = wscalar_parse(r, m0 + (size_t)m * MULT_LEN + 6 * WP + (size_t)i * WS) && good;
        }
        if (kind == SUB_PADD)
          for (int e = 0; e < 2; e++) {
            const uint8_t* ep = proof + 4 * WP + 4 * MULT_LEN + (size_t)e * EQ_LEN;
            for (int i = 0; i < 2; i++) good = tom_parse(x, y, ep + (size_t)i * WP) && good;
            for (int i = 0; i < 3; i++) good = wscalar_parse(r, ep + 2 * WP + (size_t)i * WS) && good;
          }
      }
    }
    Fold f;
    zero_n<8>(f.gW); zero_n<8>(f.hW);
    f.tape_ok = true;
    uint32_t zero[8];
    zero_n<8>(zero);
    for (int e = 0; e < SUB_ENT_MAX; e++) emit(b, e, zero, 0);   // unused entries: scalar 0 on the first input point
    if (!good) {
      ZK_SET_STATUS(status + b, ZKA_ERR_MALFORMED);
    } else if (kind == SUB_EQ) {
      uint32_t c1[8], c2[8], es[2][8];
      zero_n<8>(c1); zero_n<8>(c2);
      equality(f, row, row + WP, proof, dr, c1, c2, es);
      emit_m(b, 0, c1, 0); emit_m(b, 1, c2, WP);
      emit(b, 2, es[0], poff); emit(b, 3, es[1], poff + WP);
    } else if (kind == SUB_MULT) {
      uint32_t cx[8], cy[8], cz[8], es[6][8];
      zero_n<8>(cx); zero_n<8>(cy); zero_n<8>(cz);
      mult(f, row, row + WP, row + 2 * WP, proof, dr, cx, cy, cz, es);
      emit_m(b, 0, cx, 0); emit_m(b, 1, cy, WP); emit_m(b, 2, cz, 2 * WP);
      for (int i = 0; i < 6; i++) emit(b, 3 + i, es[i], poff + (uint32_t)i * WP);
    } else {
      // aggregatePointAdd (pointAdd.ts:199-259): C1..C6 = PX QX RX PY QY RY
      const uint8_t *PX = row, *PY = row + WP, *QX = row + 2 * WP, *QY = row + 3 * WP, *RX = row + 4 * WP, *RY = row + 5 * WP;
      TomPt p1, p2, p3, p4, p5, p6, n, r;
      uint32_t x[PGL], y[PGL];
      tom_parse(x, y, PX); tom_from_affine(p1, x, y);
      tom_parse(x, y, QX); tom_from_affine(p2, x, y);
      tom_parse(x, y, RX); tom_from_affine(p3, x, y);
      tom_parse(x, y, PY); tom_from_affine(p4, x, y);
      tom_parse(x, y, QY); tom_from_affine(p5, x, y);
      tom_parse(x, y, RY); tom_from_affine(p6, x, y);
      uint8_t d7[BSTRIDE], d9[BSTRIDE], d12[BSTRIDE], dix[BSTRIDE], diy[BSTRIDE];
      tom_neg(n, p1); tom_add(r, p2, n); tom_encode_proj(d7, r);        // C7 = C2 - C1
      tom_neg(n, p4); tom_add(r, p5, n); tom_encode_proj(d9, r);        // C9 = C5 - C4
      tom_neg(n, p3); tom_add(r, p1, n); tom_encode_proj(d12, r);       // C12 = C1 - C3
      tom_add(r, p3, p1); tom_add(r, r, p2); tom_encode_proj(dix, r);   // Cint = C3 + C1 + C2
      tom_add(r, p4, p6); tom_encode_proj(diy, r);                      // Cint = C4 + C6
      const uint8_t *C8 = proof, *C10 = proof + WP, *C11 = proof + 2 * WP, *C13 = proof + 3 * WP;
      const uint8_t* mp = proof + 4 * WP;
      const uint8_t* ep = mp + 4 * MULT_LEN;
      uint32_t a7[8], a8[8], a9[8], a10[8], a11[8], a12[8], a13[8], aix[8], aiy[8], ag[8];
      zero_n<8>(a7); zero_n<8>(a8); zero_n<8>(a9); zero_n<8>(a10); zero_n<8>(a11); zero_n<8>(a12); zero_n<8>(a13);
      zero_n<8>(aix); zero_n<8>(aiy); zero_n<8>(ag);
      uint32_t es[6][8], ee[2][8];
      const uint32_t om = poff + 4 * WP, oe = om + 4 * MULT_LEN;
      mult(f, d7, C8, tg_bytes, mp, dr, a7, a8, ag, es);                                  // pi_8
      for (int i = 0; i < 6; i++) emit(b, 10 + i, es[i], om + (uint32_t)i * WP);
      mult(f, C8, d9, C10, mp + MULT_LEN, dr + 32 * 5, a8, a9, a10, es);                  // pi_10
      for (int i = 0; i < 6; i++) emit(b, 16 + i, es[i], om + MULT_LEN + (uint32_t)i * WP);
      mult(f, C10, C10, C11, mp + 2 * MULT_LEN, dr + 32 * 10, a10, a10, a11, es);         // pi_11
      for (int i = 0; i < 6; i++) emit(b, 22 + i, es[i], om + 2 * MULT_LEN + (uint32_t)i * WP);
      equality(f, C11, dix, ep, dr + 32 * 15, a11, aix, ee);                              // pi_x
      emit(b, 34, ee[0], oe); emit(b, 35, ee[1], oe + WP);
      mult(f, C10, d12, C13, mp + 3 * MULT_LEN, dr + 32 * 17, a10, a12, a13, es);         // pi_13
      for (int i = 0; i < 6; i++) emit(b, 28 + i, es[i], om + 3 * MULT_LEN + (uint32_t)i * WP);
      equality(f, C13, diy, ep + EQ_LEN, dr + 32 * 22, a13, aiy, ee);                     // pi_y
      emit(b, 36, ee[0], oe + EQ_LEN); emit(b, 37, ee[1], oe + EQ_LEN + WP);
      acc_add(f.gW, ag);                                                                  // C_14 = g
      // derived commitments expanded onto the inputs
      uint32_t cPX[8], cPY[8], cQX[8], cQY[8], cRX[8], cRY[8];
      F::sub(cPX, a12, a7); F::add(cPX, cPX, aix);       // -C7 +C12 +Cint
      F::add(cQX, a7, aix);
      F::sub(cRX, aix, a12);
      F::sub(cPY, aiy, a9);
      copy_n<8>(cQY, a9);
      copy_n<8>(cRY, aiy);
      emit_m(b, 0, cPX, 0); emit_m(b, 1, cPY, WP); emit_m(b, 2, cQX, 2 * WP); emit_m(b, 3, cQY, 3 * WP);
      emit_m(b, 4, cRX, 4 * WP); emit_m(b, 5, cRY, 5 * WP);
      emit_m(b, 6, a8, poff); emit_m(b, 7, a10, poff + WP); emit_m(b, 8, a11, poff + 2 * WP); emit_m(b, 9, a13, poff + 3 * WP);
    }
    if (good && !f.tape_ok) ZK_SET_STATUS(status + b, ZKA_ERR_TAPE_RANGE);
    uint32_t v[8];
    zero_n<8>(v);
    st<8>(fx_jv + (size_t)b * 16, v); st<8>(fx_jr + (size_t)b * 16, v);
    F::from_mont(v, f.gW); st<8>(fx_jv + (size_t)b * 16 + 8, v);
    F::from_mont(v, f.hW); st<8>(fx_jr + (size_t)b * 16 + 8, v);
  }
};
struct VSubFinalTask {
  int32_t* status;
  const uint8_t* id_flags;   // [B][3], verdict at [b][1]
  uint8_t* ok;
  ZK_HD void operator()(int b) const { ok[b] = (status[b] == ZKA_OK && id_flags[(size_t)b * 3 + 1]) ? 1 : 0; }
};
struct VConcatTask {   // rows[b] = a[b] (la bytes) || c[b] (lc bytes)
  const uint8_t *a, *c;
  int la, lc;
  uint8_t* rows;
  size_t stride;
  ZK_HD void operator()(int t) const {
    const int per = la + lc;
    const int b = t / per, o = t % per;
    rows[(size_t)b * stride + o] = o < la ? a[(size_t)b * la + o] : c[(size_t)b * lc + (o - la)];
  }
};

// final verdict (zkpAttestList.ts:165-183): GK first, then exp
struct VFinalTask {
  VerifyCtx c;
  ZK_HD void operator()(int b) const {
    const uint8_t* fl = c.id_flags + (size_t)b * 3;
    // the aggregate check of the chunk stands for every per-proof identity test of its group (zk_verify_agg.cuh)
    const bool tom_pass = c.agg_ctl && c.agg_ctl[AGG_TOM_PASS], nist_pass = c.agg_ctl && c.agg_ctl[AGG_NIST_PASS];
    const bool f[3] = {tom_pass || fl[0], tom_pass || fl[1], nist_pass || fl[2]};
    const int st = c.status[b];
    const bool gk = c.gk_ok_len[b] && f[0];
    if (st == ZKA_ERR_MALFORMED || st == ZKA_ERR_R_INFINITY) { c.ok[b] = 0; return; }
    if (!gk) {
      // verifyMembership returned false before verifyExp could throw: not an error
      if (st != ZKA_ERR_TAPE_RANGE) c.status[b] = ZKA_OK;
      c.ok[b] = 0;
      return;
    }
    c.ok[b] = (st == ZKA_OK && f[1] && f[2]) ? 1 : 0;
  }
};

}  // namespace zk
