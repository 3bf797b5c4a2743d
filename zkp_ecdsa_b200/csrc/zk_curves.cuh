// zk_curves.cuh — group law for P-256 and tomEdwards256 on the GPU.
//
// P-256: homogeneous projective (X:Y:Z), complete a=-3 formulas of Renes-Costello-Batina
//   2015 (Alg. 4 add, Alg. 5 mixed add, Alg. 6 double) — the same formulas the reference
//   uses (/root/reference/src/curves/weier.ts:133-230), on Montgomery residues.
// tomEdwards256: the reference works on  a x^2 + y^2 = 1 + d x^2 y^2  with Hisil et al.
//   extended coordinates (/root/reference/src/curves/edwards.ts:141-183).  Here every point
//   is moved once to the isomorphic curve  x'^2 + y^2 = 1 + d' x'^2 y'^2  (x' = sqrt(a) x,
//   d' = d/a; a is a square, d a non-square, so the unified addition stays complete) which
//   saves the multiplication by `a` in every addition; x is mapped back when a point is
//   normalised.  Only affine bytes (toBytes) are observable, so results are bit-identical.
#pragma once
#include "zk_field.cuh"

namespace zk {

// ------------------------------------------------------------------------------------ P-256
struct P256Pt {  // projective, Montgomery residues mod p256.p
  uint32_t x[8], y[8], z[8];
};
struct P256Aff {  // affine Montgomery residues; inf != 0 marks the identity
  uint32_t x[8], y[8];
};

ZK_HD void p256_const_b(uint32_t* r) {
  constexpr uint32_t t[8] = ZK_P256_B_MONT;
#pragma unroll
  for (int i = 0; i < 8; i++) r[i] = t[i];
}
ZK_HD void p256_set_identity(P256Pt& p) {
  zero_n<8>(p.x);
  P256p::set_one(p.y);
  zero_n<8>(p.z);
}
ZK_HD void p256_set_generator(P256Aff& g) {
  constexpr uint32_t gx[8] = ZK_P256_GX_MONT;
  constexpr uint32_t gy[8] = ZK_P256_GY_MONT;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    g.x[i] = gx[i];
    g.y[i] = gy[i];
  }
}
ZK_HD void p256_from_affine(P256Pt& p, const P256Aff& a) {
  copy_n<8>(p.x, a.x);
  copy_n<8>(p.y, a.y);
  P256p::set_one(p.z);
}
ZK_HD bool p256_is_identity(const P256Pt& p) { return is_zero_n<8>(p.z); }

// y^2 == x^3 - 3x + b  (weier.ts:56-70 with z = 1), Montgomery residues
ZK_HD bool p256_on_curve(const uint32_t* x, const uint32_t* y) {
  using F = P256p;
  uint32_t l[8], r[8], t[8], b[8];
  F::sqr(l, y);
  F::sqr(t, x);
  F::mul(r, t, x);
  F::add(t, x, x);
  F::add(t, t, x);
  F::sub(r, r, t);
  p256_const_b(b);
  F::add(r, r, b);
  return eq_n<8>(l, r);
}

// RCB15 Algorithm 6 (a = -3): 8M + 3S + 2m_b  (weier.ts:133-175)
ZK_HD void p256_dbl(P256Pt& r, const P256Pt& p) {
  using F = P256p;
  uint32_t t0[8], t1[8], t2[8], t3[8], x3[8], y3[8], z3[8], b[8];
  p256_const_b(b);
  F::sqr(t0, p.x);
  F::sqr(t1, p.y);
  F::sqr(t2, p.z);
  F::mul(t3, p.x, p.y);
  F::add(t3, t3, t3);
  F::mul(z3, p.x, p.z);
  F::add(z3, z3, z3);
  F::mul(y3, b, t2);
  F::sub(y3, y3, z3);
  F::add(x3, y3, y3);
  F::add(y3, x3, y3);
  F::sub(x3, t1, y3);
  F::add(y3, t1, y3);
  F::mul(y3, x3, y3);
  F::mul(x3, x3, t3);
  F::add(t3, t2, t2);
  F::add(t2, t2, t3);
  F::mul(z3, b, z3);
  F::sub(z3, z3, t2);
  F::sub(z3, z3, t0);
  F::add(t3, z3, z3);
  F::add(z3, z3, t3);
  F::add(t3, t0, t0);
  F::add(t0, t3, t0);
  F::sub(t0, t0, t2);
  F::mul(t0, t0, z3);
  F::add(y3, y3, t0);
  F::mul(t0, p.y, p.z);
  F::add(t0, t0, t0);
  F::mul(z3, t0, z3);
  F::sub(x3, x3, z3);
  F::mul(z3, t0, t1);
  F::add(z3, z3, z3);
  F::add(z3, z3, z3);
  copy_n<8>(r.x, x3);
  copy_n<8>(r.y, y3);
  copy_n<8>(r.z, z3);
}

// Jacobian doubling for a = -3 ("dbl-2001-b", 3M + 5S = 8 multiplications instead of 13): used only
// in the long DOUBLING CHAINS (16^j R for the per-proof table, the one-shot ladder u2*pk), which
// are latency-bound with one thread per proof.  Doubling has no exceptional case on a prime-order
// curve (no points of order 2) and maps the identity (Z = 0) to itself.
struct P256Jac {
  uint32_t x[8], y[8], z[8];   // x = X/Z^2, y = Y/Z^3
};
ZK_HD void p256_jac_dbl(P256Jac& r, const P256Jac& p) {
  using F = P256p;
  uint32_t delta[8], gamma[8], beta[8], alpha[8], t0[8], t1[8];
  F::sqr(delta, p.z);
  F::sqr(gamma, p.y);
  F::mul(beta, p.x, gamma);
  F::sub(t0, p.x, delta);
  F::add(t1, p.x, delta);
  F::mul(alpha, t0, t1);
  F::add(t0, alpha, alpha);
  F::add(alpha, t0, alpha);            // 3 (X - delta)(X + delta)
  F::add(t0, p.y, p.z);
  F::sqr(t0, t0);
  F::sub(t0, t0, gamma);
  F::sub(r.z, t0, delta);              // Z3 = (Y + Z)^2 - gamma - delta
  F::add(t0, beta, beta);
  F::add(t0, t0, t0);                  // 4 beta
  F::add(t1, t0, t0);                  // 8 beta
  F::sqr(r.x, alpha);
  F::sub(r.x, r.x, t1);                // X3 = alpha^2 - 8 beta
  F::sub(t0, t0, r.x);
  F::mul(t0, alpha, t0);
  F::sqr(t1, gamma);
  F::add(t1, t1, t1);
  F::add(t1, t1, t1);
  F::add(t1, t1, t1);                  // 8 gamma^2
  F::sub(r.y, t0, t1);                 // Y3 = alpha (4 beta - X3) - 8 gamma^2
}
// Jacobian (X:Y:Z) -> homogeneous (X Z : Y : Z^3);  homogeneous (X:Y:Z) -> Jacobian (X Z : Y Z^2 : Z)
ZK_HD void p256_jac_to_hom(P256Pt& r, const P256Jac& p) {
  using F = P256p;
  uint32_t z2[8];
  F::sqr(z2, p.z);
  F::mul(r.x, p.x, p.z);
  copy_n<8>(r.y, p.y);
  F::mul(r.z, z2, p.z);
}
ZK_HD void p256_hom_to_jac(P256Jac& r, const P256Pt& p) {
  using F = P256p;
  uint32_t z2[8];
  F::sqr(z2, p.z);
  F::mul(r.x, p.x, p.z);
  F::mul(r.y, p.y, z2);
  copy_n<8>(r.z, p.z);
  if (is_zero_n<8>(p.z)) {   // identity (0:Y:0) -> (1:1:0); (0:0:0) would not survive the way back
    F::set_one(r.x);
    F::set_one(r.y);
  }
}

// shared tail of RCB15 Alg. 4 / Alg. 5 after t0,t1,t2,t3,t4,y3 are formed:
//   t0 = X1X2, t1 = Y1Y2, t2 = Z1Z2, t3 = X1Y2+X2Y1, t4 = Y1Z2+Y2Z1, y3 = X1Z2+X2Z1
ZK_HD void p256_add_tail(P256Pt& r, uint32_t* t0, uint32_t* t1, uint32_t* t2, uint32_t* t3, uint32_t* t4,
                         uint32_t* y3) {
  using F = P256p;
  uint32_t x3[8], z3[8], b[8];
  p256_const_b(b);
  F::mul(z3, b, t2);
  F::sub(x3, y3, z3);
  F::add(z3, x3, x3);
  F::add(x3, x3, z3);
  F::sub(z3, t1, x3);
  F::add(x3, t1, x3);
  F::mul(y3, b, y3);
  F::add(t1, t2, t2);
  F::add(t2, t1, t2);
  F::sub(y3, y3, t2);
  F::sub(y3, y3, t0);
  F::add(t1, y3, y3);
  F::add(y3, t1, y3);
  F::add(t1, t0, t0);
  F::add(t0, t1, t0);
  F::sub(t0, t0, t2);
  F::mul(t1, t4, y3);
  F::mul(t2, t0, y3);
  F::mul(y3, x3, z3);
  F::add(y3, y3, t2);
  F::mul(x3, t3, x3);
  F::sub(x3, x3, t1);
  F::mul(z3, t4, z3);
  F::mul(t1, t3, t0);
  F::add(z3, z3, t1);
  copy_n<8>(r.x, x3);
  copy_n<8>(r.y, y3);
  copy_n<8>(r.z, z3);
}

// RCB15 Algorithm 4 (a = -3): 12M + 2m_b  (weier.ts:176-230)
ZK_HD void p256_add(P256Pt& r, const P256Pt& p, const P256Pt& q) {
  using F = P256p;
  uint32_t t0[8], t1[8], t2[8], t3[8], t4[8], x3[8], y3[8];
  F::mul(t0, p.x, q.x);
  F::mul(t1, p.y, q.y);
  F::mul(t2, p.z, q.z);
  F::add(t3, p.x, p.y);
  F::add(t4, q.x, q.y);
  F::mul(t3, t3, t4);
  F::add(t4, t0, t1);
  F::sub(t3, t3, t4);
  F::add(t4, p.y, p.z);
  F::add(x3, q.y, q.z);
  F::mul(t4, t4, x3);
  F::add(x3, t1, t2);
  F::sub(t4, t4, x3);
  F::add(x3, p.x, p.z);
  F::add(y3, q.x, q.z);
  F::mul(x3, x3, y3);
  F::add(y3, t0, t2);
  F::sub(y3, x3, y3);
  p256_add_tail(r, t0, t1, t2, t3, t4, y3);
}

// RCB15 Algorithm 5 (mixed, Z2 = 1, a = -3): 11M + 2m_b.  q must not be the identity.
ZK_HD void p256_madd(P256Pt& r, const P256Pt& p, const P256Aff& q) {
  using F = P256p;
  uint32_t t0[8], t1[8], t2[8], t3[8], t4[8], y3[8];
  F::mul(t0, p.x, q.x);
  F::mul(t1, p.y, q.y);
  F::add(t3, p.x, p.y);
  F::add(t4, q.x, q.y);
  F::mul(t3, t3, t4);
  F::add(t4, t0, t1);
  F::sub(t3, t3, t4);
  F::mul(t4, q.y, p.z);
  F::add(t4, t4, p.y);
  F::mul(y3, q.x, p.z);
  F::add(y3, y3, p.x);
  copy_n<8>(t2, p.z);
  p256_add_tail(r, t0, t1, t2, t3, t4, y3);
}

// ----------------------------------------------------------------------------- tomEdwards256
// Extended coordinates on the a'=1 image curve, lazy Montgomery residues mod tom.p.
struct TomPt {
  uint32_t x[9], y[9], t[9], z[9];
};
// Precomputed affine table entry: (x', y, k = d' x' y), canonical residues (< p).
struct TomPre {
  uint32_t x[9], y[9], k[9];
};

ZK_HD void tom_const(uint32_t* r, int which) {
  constexpr uint32_t sa[9] = ZK_TOM_SQRTA_MONT;
  constexpr uint32_t isa[9] = ZK_TOM_INVSQRTA_MONT;
  constexpr uint32_t d1[9] = ZK_TOM_D1_MONT;
  constexpr uint32_t gx[9] = ZK_TOM_GX1_MONT;
  constexpr uint32_t gy[9] = ZK_TOM_GY_MONT;
  constexpr uint32_t s2[9] = ZK_TOM_SQRTND1_MONT;
  constexpr uint32_t is2[9] = ZK_TOM_INVSQRTND1_MONT;
  constexpr uint32_t dd2[9] = ZK_TOM_2D2_MONT;
  constexpr uint32_t isd[9] = ZK_TOM_INVSQRTND_MONT;
#pragma unroll
  for (int i = 0; i < 9; i++)
    r[i] = which == 0 ? sa[i] : which == 1 ? isa[i] : which == 2 ? d1[i] : which == 3 ? gx[i] : which == 4 ? gy[i]
           : which == 5 ? s2[i] : which == 6 ? is2[i] : which == 7 ? dd2[i] : isd[i];
}
enum { TOM_SQRTA = 0, TOM_INVSQRTA = 1, TOM_D1 = 2, TOM_GX1 = 3, TOM_GY = 4, TOM_SQRTND1 = 5, TOM_INVSQRTND1 = 6,
       TOM_2D2 = 7, TOM_INVSQRTND = 8 };

// ---- second image curve E2: -w^2 + v^2 = 1 + d2 w^2 v^2, (w, v) = (sqrt(-d1) x', 1/y) -------------
// Used ONLY by the prover's fixed-base commitment kernel (all its points lie in the prime-order
// subgroup generated by g, where the a = -1 formulas have no exceptional cases; gen_consts.py).
// Table entry: (v - w, v + w, 2 d2 w v).  Mixed addition "madd-2008-hwcd-3": 7M.
// Field with the 258-bit multiplier INLINED at every use (no call, no argument marshalling): for loop bodies
// that contain one mixed addition (7-8 products, ~40 KB of code) and are not unrolled.
struct TompInl : Tomp {
  ZK_HD static void mul(uint32_t* r, const uint32_t* a, const uint32_t* b) {
#if defined(__CUDA_ARCH__) && !defined(ZKA_NO_PTX_MUL)
    ptx::tom_mul_body(r, a, b);
#else
    Tomp::mul(r, a, b);
#endif
  }
  ZK_HD static void sqr(uint32_t* r, const uint32_t* a) { mul(r, a, a); }
};
#if defined(ZKA_COMMIT_INLINE)
using TompCommit = TompInl;
#else
using TompCommit = Tomp;
#endif
#if defined(ZKA_MSM_INLINE)
using TompMsm = TompInl;
#else
using TompMsm = Tomp;
#endif

template <bool kNeedT, class F = Tomp>
ZK_HD void tom2_madd(TomPt& r, const TomPt& p, const TomPre& q) {   // q.x = v-w, q.y = v+w, q.k = 2 d2 w v
  uint32_t A[9], B[9], C[9], D[9], E[9], Fv[9], G[9], H[9];
  F::sub(A, p.y, p.x);
  F::mul(A, A, q.x);
  F::add(B, p.y, p.x);
  F::mul(B, B, q.y);
  F::mul(C, p.t, q.k);
  F::add(D, p.z, p.z);
  F::sub(E, B, A);
  F::sub(Fv, D, C);
  F::add(G, D, C);
  F::add(H, B, A);
  F::mul(r.x, E, Fv);
  F::mul(r.y, G, H);
  if (kNeedT) F::mul(r.t, E, H);
  F::mul(r.z, Fv, G);
}

ZK_HD void tom_set_identity(TomPt& p) {
  zero_n<9>(p.x);
  Tomp::set_one(p.y);
  zero_n<9>(p.t);
  Tomp::set_one(p.z);
}
// from affine image-curve coordinates (x', y) (Montgomery)
ZK_HD void tom_from_affine(TomPt& p, const uint32_t* x1, const uint32_t* y) {
  copy_n<9>(p.x, x1);
  copy_n<9>(p.y, y);
  Tomp::mul(p.t, x1, y);
  Tomp::set_one(p.z);
}
ZK_HD void tom_set_generator(TomPt& p) {
  uint32_t gx[9], gy[9];
  tom_const(gx, TOM_GX1);
  tom_const(gy, TOM_GY);
  tom_from_affine(p, gx, gy);
}
// x'^2 + y^2 == 1 + d' x'^2 y^2   (edwards.ts:52-65 on the image curve, z = 1)
ZK_HD bool tom_on_curve(const uint32_t* x1, const uint32_t* y) {
  using F = Tomp;
  uint32_t xx[9], yy[9], l[9], r[9], d1[9], one[9];
  F::sqr(xx, x1);
  F::sqr(yy, y);
  F::add(l, xx, yy);
  F::mul(r, xx, yy);
  tom_const(d1, TOM_D1);
  F::mul(r, r, d1);
  F::set_one(one);
  F::add(r, r, one);
  // l == r  <=>  (l - r) * 1 / R is 0 or p: one Montgomery product (output < 2p) and two compares instead of
  // two 14-step canonical-reduction ladders (this test runs once per point of every proof in VValidateTask)
  uint32_t d[9], o[9];
  F::sub(d, l, r);               // l + 8p - r, r < 3p
  zero_n<9>(o);
  o[0] = 1;
  F::mul(d, d, o);
  uint32_t nz = 0, np = 0;
#pragma unroll
  for (int i = 0; i < 9; i++) { nz |= d[i]; np |= d[i] ^ FpTom::p(i); }
  return nz == 0 || np == 0;
}

// Hisil et al. 2008 section 3.1 unified addition with a = 1 (edwards.ts:161-183): 9M
ZK_HD void tom_add(TomPt& r, const TomPt& p, const TomPt& q) {
  using F = Tomp;
  uint32_t A[9], B[9], C[9], D[9], E[9], Fv[9], G[9], H[9], d1[9];
  tom_const(d1, TOM_D1);
  F::mul(A, p.x, q.x);
  F::mul(B, p.y, q.y);
  F::mul(C, p.t, q.t);
  F::mul(C, C, d1);
  F::mul(D, p.z, q.z);
  F::add(E, p.x, p.y);
  F::add(H, q.x, q.y);
  F::mul(E, E, H);
  F::sub(E, E, A);
  F::sub(E, E, B);
  F::sub(Fv, D, C);
  F::add(G, D, C);
  F::sub(H, B, A);
  F::mul(r.x, E, Fv);
  F::mul(r.y, G, H);
  F::mul(r.t, E, H);
  F::mul(r.z, Fv, G);
}
// mixed addition with a precomputed entry (Z2 = 1, k = d' x2 y2): 7M (+1M for T3)
template <bool kNeedT, class F = Tomp>
ZK_HD void tom_madd(TomPt& r, const TomPt& p, const TomPre& q) {
  uint32_t A[9], B[9], C[9], E[9], Fv[9], G[9], H[9];
  F::mul(A, p.x, q.x);
  F::mul(B, p.y, q.y);
  F::mul(C, p.t, q.k);
  F::add(E, p.x, p.y);
  F::add(H, q.x, q.y);
  F::mul(E, E, H);
  F::sub(E, E, A);
  F::sub(E, E, B);
  F::sub(Fv, p.z, C);
  F::add(G, p.z, C);
  F::sub(H, B, A);
  F::mul(r.x, E, Fv);
  F::mul(r.y, G, H);
  if (kNeedT) F::mul(r.t, E, H);
  F::mul(r.z, Fv, G);
}
// Hisil et al. 2008 section 3.3 doubling with a = 1 (edwards.ts:141-160): 4M + 4S
ZK_HD void tom_dbl(TomPt& r, const TomPt& p) {
  using F = Tomp;
  uint32_t A[9], B[9], C[9], E[9], G[9], Fv[9], H[9];
  F::sqr(A, p.x);
  F::sqr(B, p.y);
  F::sqr(C, p.z);
  F::add(C, C, C);
  F::add(E, p.x, p.y);
  F::sqr(E, E);
  F::sub(E, E, A);
  F::sub(E, E, B);
  F::add(G, A, B);   // D + B with D = a'A = A
  F::sub(Fv, G, C);
  F::sub(H, A, B);   // D - B
  F::mul(r.x, E, Fv);
  F::mul(r.y, G, H);
  F::mul(r.t, E, H);
  F::mul(r.z, Fv, G);
}
ZK_HD void tom_neg(TomPt& r, const TomPt& p) {
  Tomp::neg(r.x, p.x);
  copy_n<9>(r.y, p.y);
  Tomp::neg(r.t, p.t);
  copy_n<9>(r.z, p.z);
}

}  // namespace zk
