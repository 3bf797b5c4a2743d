// zkattest_cpu.cpp — C++ CPU restatement of the reference's hot path (TEST / BASELINE INFRASTRUCTURE).
//
// This file is part of oracle/: it is the "reference algorithm on host cores" arm of bench.py
// (`cpu_baseline.kind = "port-c++"`, `--impl reference`) and a fast checker for the parity tests.  It is
// NEVER linked into or loaded by the product (zkp_ecdsa_b200/libzkattest.so has no CPU path).
//
// It follows the reference's own algorithms, not the GPU's:
//   * Point.mul / Point.dblmul: 4-bit fixed windows over the hex digits of the scalar, 16-entry tables
//     rebuilt per call                                   /root/reference/src/curves/group.ts:97-152
//   * P-256: Renes-Costello-Batina complete a = -3 formulas, homogeneous projective
//                                                        /root/reference/src/curves/weier.ts:133-230
//   * tomEdwards256: Hisil et al. extended coordinates with general a, d
//                                                        /root/reference/src/curves/edwards.ts:141-183
//   * toBytes = toAffine (one modular inversion) + big-endian   weier.ts:231-255, edwards.ts:184-203
//   * hashPoints = SHA-256, first 10 bytes               group.ts:221-233
//   * Pedersen / Equality / Mult / PointAdd / Exp / GK / interpolate / Relation / MultiMult (Bos-Coster)
//                                                        src/commit/*.ts, src/exp/*.ts, src/proofGK/*.ts,
//                                                        src/curves/multimult.ts
//   * proveSignatureList / verifySignatureList / generateParamsList   src/zkpAttestList.ts:88-184
// Differences from the TypeScript, all unobservable in the bytes: BigInt `(a*b) % p` becomes Montgomery
// multiplication on 64-bit limbs (unsigned __int128), invEuclid becomes a Fermat ladder (same value for a
// prime modulus, 0 -> 0), crypto.getRandomValues becomes the caller's tape (include/zkattest.h contract:
// a draw >= its modulus is flagged ZKA_ERR_TAPE_RANGE instead of redrawn — the host pre-filters).
//
// It exports the same C ABI as the product (include/zkattest.h) so that the ctypes binding and every
// parity helper of tests/common.py can drive it unchanged; `device` is ignored.
#include "zkattest.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <string>
#include <thread>
#include <vector>

namespace {

typedef uint64_t u64;
typedef unsigned __int128 u128;
constexpr int ML = 5;   // limbs of the widest modulus (tom.p, 258 bits)

struct Fe {
  u64 v[ML];
};
inline Fe fe_zero() { Fe r; memset(&r, 0, sizeof r); return r; }
inline bool fe_is_zero(const Fe& a) { return (a.v[0] | a.v[1] | a.v[2] | a.v[3] | a.v[4]) == 0; }
inline bool fe_eq(const Fe& a, const Fe& b) { return memcmp(&a, &b, sizeof(Fe)) == 0; }
inline int fe_cmp(const Fe& a, const Fe& b) {
  for (int i = ML - 1; i >= 0; i--) {
    if (a.v[i] < b.v[i]) return -1;
    if (a.v[i] > b.v[i]) return 1;
  }
  return 0;
}
inline u64 fe_add_raw(Fe& r, const Fe& a, const Fe& b) {
  u128 c = 0;
  for (int i = 0; i < ML; i++) { c += (u128)a.v[i] + b.v[i]; r.v[i] = (u64)c; c >>= 64; }
  return (u64)c;
}
inline u64 fe_sub_raw(Fe& r, const Fe& a, const Fe& b) {   // returns borrow
  u64 br = 0;
  for (int i = 0; i < ML; i++) {
    u128 d = (u128)a.v[i] - b.v[i] - br;
    r.v[i] = (u64)d;
    br = (u64)(d >> 64) & 1;
  }
  return br;
}
inline int fe_bitlen(const Fe& a) {
  for (int i = ML - 1; i >= 0; i--)
    if (a.v[i]) return 64 * i + 64 - __builtin_clzll(a.v[i]);
  return 0;
}
Fe fe_from_hex(const char* h) {
  Fe r = fe_zero();
  int len = (int)strlen(h);
  for (int i = 0; i < len; i++) {
    char ch = h[len - 1 - i];
    u64 d = ch <= '9' ? ch - '0' : (ch | 32) - 'a' + 10;
    r.v[i / 16] |= d << (4 * (i % 16));
  }
  return r;
}
Fe fe_from_be(const uint8_t* b, int n) {
  Fe r = fe_zero();
  for (int k = 0; k < n; k++) {
    int pos = n - 1 - k;
    if (pos / 8 < ML) r.v[pos / 8] |= (u64)b[k] << (8 * (pos % 8));
  }
  return r;
}
void fe_to_be(uint8_t* b, const Fe& a, int n) {
  for (int k = 0; k < n; k++) {
    int pos = n - 1 - k;
    b[k] = pos / 8 < ML ? (uint8_t)(a.v[pos / 8] >> (8 * (pos % 8))) : 0;
  }
}

// ------------------------------------------------------------------------------ modular arithmetic
// big.ts:36-119 (posMod, expMod, invMod) on Montgomery residues; all values canonical in [0, p).
struct Mod {
  int n;        // 64-bit limbs
  Fe p, rr, one, pm2;
  u64 n0;       // -1/p mod 2^64
  int nbytes;   // sizeFieldBytes(): ceil(bits / 8)   (group.ts:49-52)
};
template <int N>
inline void mmul_n(u64* r, const u64* a, const u64* b, const Mod& m) {
  u64 t[N + 2];
  for (int i = 0; i < N + 2; i++) t[i] = 0;
  for (int i = 0; i < N; i++) {
    u128 c = 0;
    for (int j = 0; j < N; j++) { c += (u128)a[j] * b[i] + t[j]; t[j] = (u64)c; c >>= 64; }
    c += t[N]; t[N] = (u64)c; t[N + 1] = (u64)(c >> 64);
    const u64 q = t[0] * m.n0;
    c = ((u128)q * m.p.v[0] + t[0]) >> 64;
    for (int j = 1; j < N; j++) { c += (u128)q * m.p.v[j] + t[j]; t[j - 1] = (u64)c; c >>= 64; }
    c += t[N]; t[N - 1] = (u64)c; t[N] = t[N + 1] + (u64)(c >> 64);
  }
  // t < 2p: one conditional subtraction
  u64 u[N], br = 0;
  for (int i = 0; i < N; i++) { u128 d = (u128)t[i] - m.p.v[i] - br; u[i] = (u64)d; br = (u64)(d >> 64) & 1; }
  const bool ge = t[N] != 0 || br == 0;
  for (int i = 0; i < N; i++) r[i] = ge ? u[i] : t[i];
}
inline Fe mmul(const Fe& a, const Fe& b, const Mod& m) {
  Fe r = fe_zero();
  if (m.n == 4) mmul_n<4>(r.v, a.v, b.v, m); else mmul_n<5>(r.v, a.v, b.v, m);
  return r;
}
inline Fe madd(const Fe& a, const Fe& b, const Mod& m) {
  Fe s, t;
  u64 c = fe_add_raw(s, a, b);
  u64 br = fe_sub_raw(t, s, m.p);
  return (c || !br) ? t : s;
}
inline Fe msub(const Fe& a, const Fe& b, const Mod& m) {
  Fe s, t;
  u64 br = fe_sub_raw(s, a, b);
  fe_add_raw(t, s, m.p);
  return br ? t : s;
}
inline Fe mneg(const Fe& a, const Mod& m) { return msub(fe_zero(), a, m); }
inline Fe to_mont(const Fe& a, const Mod& m) { return mmul(a, m.rr, m); }
inline Fe from_mont(const Fe& a, const Mod& m) { Fe o = fe_zero(); o.v[0] = 1; return mmul(a, o, m); }
Fe mpow(const Fe& a, const Fe& e, const Mod& m) {   // expMod, big.ts:44-59 (a Montgomery, e plain)
  Fe r = m.one, q = a;
  const int bl = fe_bitlen(e);
  for (int i = 0; i < bl; i++) {
    if ((e.v[i / 64] >> (i % 64)) & 1) r = mmul(r, q, m);
    q = mmul(q, q, m);
  }
  return r;
}
inline Fe minv(const Fe& a, const Mod& m) { return mpow(a, m.pm2, m); }   // invMod: 0 -> 0 like invEuclid
Mod make_mod(const char* hex) {
  Mod m;
  m.p = fe_from_hex(hex);
  const int bits = fe_bitlen(m.p);
  m.n = (bits + 63) / 64;
  m.nbytes = (bits + 7) / 8;
  u64 inv = 1;
  for (int i = 0; i < 7; i++) inv *= 2 - m.p.v[0] * inv;
  m.n0 = (u64)0 - inv;
  Fe r = fe_zero();
  r.v[0] = 1;
  Fe two_r = r;
  for (int pass = 0; pass < 2; pass++) {
    for (int i = 0; i < 64 * m.n; i++) {   // r <- 2r mod p
      Fe s, t;
      u64 c = fe_add_raw(s, r, r);
      u64 br = fe_sub_raw(t, s, m.p);
      r = (c || !br) ? t : s;
    }
    if (pass == 0) two_r = r;
  }
  m.one = two_r;
  m.rr = r;
  Fe two = fe_zero();
  two.v[0] = 2;
  fe_sub_raw(m.pm2, m.p, two);
  return m;
}

// instances.ts:22-54
const Mod P256P = make_mod("ffffffff00000001000000000000000000000000ffffffffffffffffffffffff");
const Mod P256N = make_mod("ffffffff00000000ffffffffffffffffbce6faada7179e84f3b9cac2fc632551");
const Mod TOMP = make_mod("3fffffffc000000040000000000000002ae382c7957cc4ff9713c3d82bc47d3af");

// scalars mod a group order: plain canonical integers (Group.Scalar, group.ts:159-218)
inline Fe s_add(const Fe& a, const Fe& b, const Mod& q) { return madd(a, b, q); }
inline Fe s_sub(const Fe& a, const Fe& b, const Mod& q) { return msub(a, b, q); }
inline Fe s_neg(const Fe& a, const Mod& q) { return mneg(a, q); }
inline Fe s_mul(const Fe& a, const Fe& b, const Mod& q) { return mmul(mmul(a, b, q), q.rr, q); }
inline Fe s_inv(const Fe& a, const Mod& q) { return from_mont(minv(to_mont(a, q), q), q); }
inline Fe s_small(u64 v) { Fe r = fe_zero(); r.v[0] = v; return r; }
inline Fe s_reduce(const Fe& a, const Mod& q) {   // newScalar: value mod order, for inputs < 2^256 < 2q... (loop: any)
  Fe r = a;
  while (fe_cmp(r, q.p) >= 0) { Fe t; fe_sub_raw(t, r, q.p); r = t; }
  return r;
}
Fe s_pow_small(const Fe& x, int e, const Mod& q) {   // expMod(x, e, order) for small e
  Fe r = s_small(1);
  for (int i = 0; i < e; i++) r = s_mul(r, x, q);
  return r;
}

// ------------------------------------------------------------------------------------ SHA-256
struct Sha256 {
  uint32_t h[8];
  uint8_t buf[64];
  uint64_t len;
  int fill;
  static uint32_t rotr(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }
  void init() {
    static const uint32_t iv[8] = {0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a, 0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19};
    memcpy(h, iv, sizeof h);
    len = 0;
    fill = 0;
  }
  void block(const uint8_t* p) {
    static const uint32_t K[64] = {
        0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5, 0xd807aa98, 0x12835b01,
        0x243185be, 0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174, 0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc,
        0x2de92c6f, 0x4a7484aa, 0x5cb0a9dc, 0x76f988da, 0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147,
        0x06ca6351, 0x14292967, 0x27b70a85, 0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85,
        0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3, 0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070, 0x19a4c116, 0x1e376c08,
        0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f, 0x682e6ff3, 0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208,
        0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2};
    uint32_t w[64];
    for (int i = 0; i < 16; i++) w[i] = ((uint32_t)p[4 * i] << 24) | ((uint32_t)p[4 * i + 1] << 16) | ((uint32_t)p[4 * i + 2] << 8) | p[4 * i + 3];
    for (int i = 16; i < 64; i++) {
      uint32_t s0 = rotr(w[i - 15], 7) ^ rotr(w[i - 15], 18) ^ (w[i - 15] >> 3);
      uint32_t s1 = rotr(w[i - 2], 17) ^ rotr(w[i - 2], 19) ^ (w[i - 2] >> 10);
      w[i] = w[i - 16] + s0 + w[i - 7] + s1;
    }
    uint32_t a = h[0], b = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7];
    for (int i = 0; i < 64; i++) {
      uint32_t S1 = rotr(e, 6) ^ rotr(e, 11) ^ rotr(e, 25), ch = (e & f) ^ (~e & g);
      uint32_t t1 = hh + S1 + ch + K[i] + w[i];
      uint32_t S0 = rotr(a, 2) ^ rotr(a, 13) ^ rotr(a, 22), mj = (a & b) ^ (a & c) ^ (b & c);
      uint32_t t2 = S0 + mj;
      hh = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
    }
    h[0] += a; h[1] += b; h[2] += c; h[3] += d; h[4] += e; h[5] += f; h[6] += g; h[7] += hh;
  }
  void update(const uint8_t* p, size_t n) {
    len += n;
    while (n) {
      size_t k = std::min<size_t>(n, 64 - fill);
      memcpy(buf + fill, p, k);
      fill += (int)k; p += k; n -= k;
      if (fill == 64) { block(buf); fill = 0; }
    }
  }
  void final(uint8_t out[32]) {
    uint64_t bits = len * 8;
    uint8_t pad = 0x80;
    update(&pad, 1);
    uint8_t z = 0;
    while (fill != 56) update(&z, 1);
    uint8_t lb[8];
    for (int i = 0; i < 8; i++) lb[i] = (uint8_t)(bits >> (56 - 8 * i));
    update(lb, 8);
    for (int i = 0; i < 8; i++) { out[4 * i] = h[i] >> 24; out[4 * i + 1] = h[i] >> 16; out[4 * i + 2] = h[i] >> 8; out[4 * i + 3] = h[i]; }
  }
};

// ------------------------------------------------------------------------------------- errors
struct ZkErr { int status; };   // mirrors a reference `throw` (status = ZKA_ERR_*)

// ------------------------------------------------------------------------------------- tape
struct Tape {   // stand-in for crypto.getRandomValues (big.ts:175)
  const uint8_t* p;
  size_t len, pos;
  Fe rnd32(const Mod& q) {   // rnd(order) with the C-ABI contract: out of range -> TAPE_RANGE, no redraw
    if (pos + 32 > len) throw ZkErr{ZKA_ERR_TAPE_RANGE};
    Fe r = fe_from_be(p + pos, 32);
    pos += 32;
    if (fe_cmp(r, q.p) >= 0) throw ZkErr{ZKA_ERR_TAPE_RANGE};
    return r;
  }
};

// ------------------------------------------------------------------------------------- groups
struct WGroup {   // WeierstrassGroup, a = -3 (weier.ts:25-90)
  const Mod* f;
  const Mod* q;
  Fe b, gx, gy;   // Montgomery
};
struct WPt { Fe x, y, z; };
struct EGroup {   // TEdwards (edwards.ts:25-87)
  const Mod* f;
  const Mod* q;
  Fe a, d, gx, gy;
};
struct EPt { Fe x, y, t, z; };

WGroup make_p256() {
  WGroup g;
  g.f = &P256P; g.q = &P256N;
  g.b = to_mont(fe_from_hex("5ac635d8aa3a93e7b3ebbd55769886bc651d06b0cc53b0f63bce3c3e27d2604b"), P256P);
  g.gx = to_mont(fe_from_hex("6b17d1f2e12c4247f8bce6e563a440f277037d812deb33a0f4a13945d898c296"), P256P);
  g.gy = to_mont(fe_from_hex("4fe342e2fe1a7f9b8ee7eb4a7c0f9e162bce33576b315ececbb6406837bf51f5"), P256P);
  return g;
}
EGroup make_tom() {
  EGroup g;
  g.f = &TOMP; g.q = &P256P;   // order = p256.p (instances.ts:48)
  g.a = to_mont(fe_from_hex("1abce3fd8e1d7a21252515332a512e09d4249bd5b1ec35e316c02254fe8cedf5d"), TOMP);
  g.d = to_mont(fe_from_hex("051781d9823abde00ec99295ba542c8b1401874bcbeb9e9c861174c7bca6a02aa"), TOMP);
  g.gx = to_mont(fe_from_hex("7907055d0a7d4abc3eafdc25d431d9659fbe007ee2d8ddc4e906206ea9ba4fdb"), TOMP);
  g.gy = to_mont(fe_from_hex("be231cb9f9bf18319c9f081141559b0a33dddccd2221f0464a9cd57081b01a01"), TOMP);
  return g;
}
const WGroup P256 = make_p256();
const EGroup TOM = make_tom();

// ---- Weierstrass (weier.ts:97-261)
inline WPt identity(const WGroup& g) { return WPt{fe_zero(), g.f->one, fe_zero()}; }
inline WPt generator(const WGroup& g) { return WPt{g.gx, g.gy, g.f->one}; }
inline bool is_identity(const WGroup&, const WPt& p) { return fe_is_zero(p.x) && !fe_is_zero(p.y) && fe_is_zero(p.z); }
inline WPt neg(const WGroup& g, const WPt& p) { return WPt{p.x, mneg(p.y, *g.f), p.z}; }
inline bool pt_eq(const WGroup& g, const WPt& a, const WPt& b) {   // weier.ts:120-128
  const Mod& m = *g.f;
  return fe_eq(mmul(a.x, b.z, m), mmul(b.x, a.z, m)) && fe_eq(mmul(a.y, b.z, m), mmul(b.y, a.z, m));
}
WPt dbl(const WGroup& g, const WPt& p) {   // weier.ts:133-175 (RCB15 Alg. 6)
  const Mod& m = *g.f;
  Fe t0 = mmul(p.x, p.x, m), t1 = mmul(p.y, p.y, m), t2 = mmul(p.z, p.z, m), t3 = mmul(p.x, p.y, m);
  t3 = madd(t3, t3, m);
  Fe z3 = mmul(p.x, p.z, m);
  z3 = madd(z3, z3, m);
  Fe y3 = mmul(g.b, t2, m);
  y3 = msub(y3, z3, m);
  Fe x3 = madd(y3, y3, m);
  y3 = madd(x3, y3, m);
  x3 = msub(t1, y3, m);
  y3 = madd(t1, y3, m);
  y3 = mmul(x3, y3, m);
  x3 = mmul(x3, t3, m);
  t3 = madd(t2, t2, m);
  t2 = madd(t2, t3, m);
  z3 = mmul(g.b, z3, m);
  z3 = msub(z3, t2, m);
  z3 = msub(z3, t0, m);
  t3 = madd(z3, z3, m);
  z3 = madd(z3, t3, m);
  t3 = madd(t0, t0, m);
  t0 = madd(t3, t0, m);
  t0 = msub(t0, t2, m);
  t0 = mmul(t0, z3, m);
  y3 = madd(y3, t0, m);
  t0 = mmul(p.y, p.z, m);
  t0 = madd(t0, t0, m);
  z3 = mmul(t0, z3, m);
  x3 = msub(x3, z3, m);
  z3 = mmul(t0, t1, m);
  z3 = madd(z3, z3, m);
  z3 = madd(z3, z3, m);
  return WPt{x3, y3, z3};
}
WPt add(const WGroup& g, const WPt& p, const WPt& q) {   // weier.ts:176-230 (RCB15 Alg. 4)
  const Mod& m = *g.f;
  Fe t0 = mmul(p.x, q.x, m), t1 = mmul(p.y, q.y, m), t2 = mmul(p.z, q.z, m);
  Fe t3 = madd(p.x, p.y, m), t4 = madd(q.x, q.y, m);
  t3 = mmul(t3, t4, m);
  t4 = madd(t0, t1, m);
  t3 = msub(t3, t4, m);
  t4 = madd(p.y, p.z, m);
  Fe x3 = madd(q.y, q.z, m);
  t4 = mmul(t4, x3, m);
  x3 = madd(t1, t2, m);
  t4 = msub(t4, x3, m);
  x3 = madd(p.x, p.z, m);
  Fe y3 = madd(q.x, q.z, m);
  x3 = mmul(x3, y3, m);
  y3 = madd(t0, t2, m);
  y3 = msub(x3, y3, m);
  Fe z3 = mmul(g.b, t2, m);
  x3 = msub(y3, z3, m);
  z3 = madd(x3, x3, m);
  x3 = madd(x3, z3, m);
  z3 = msub(t1, x3, m);
  x3 = madd(t1, x3, m);
  y3 = mmul(g.b, y3, m);
  t1 = madd(t2, t2, m);
  t2 = madd(t1, t2, m);
  y3 = msub(y3, t2, m);
  y3 = msub(y3, t0, m);
  t1 = madd(y3, y3, m);
  y3 = madd(t1, y3, m);
  t1 = madd(t0, t0, m);
  t0 = madd(t1, t0, m);
  t0 = msub(t0, t2, m);
  t1 = mmul(t4, y3, m);
  t2 = mmul(t0, y3, m);
  y3 = mmul(x3, z3, m);
  y3 = madd(y3, t2, m);
  x3 = mmul(t3, x3, m);
  x3 = msub(x3, t1, m);
  z3 = mmul(t4, z3, m);
  t1 = mmul(t3, t0, m);
  z3 = madd(z3, t1, m);
  return WPt{x3, y3, z3};
}
// toAffine (weier.ts:231-243): false for the identity; x, y plain integers
bool to_affine(const WGroup& g, const WPt& p, Fe& x, Fe& y) {
  if (is_identity(g, p)) return false;
  const Mod& m = *g.f;
  Fe zi = minv(p.z, m);
  x = from_mont(mmul(p.x, zi, m), m);
  y = from_mont(mmul(p.y, zi, m), m);
  return true;
}
// toBytes into a FIXED slot: the identity (one 0x00 byte in weier.ts:247) is 65 zero bytes (include/zkattest.h)
void to_bytes(const WGroup& g, const WPt& p, uint8_t* out) {
  Fe x, y;
  const int cs = g.f->nbytes;
  if (!to_affine(g, p, x, y)) { memset(out, 0, 1 + 2 * cs); return; }
  out[0] = 0x04;
  fe_to_be(out + 1, x, cs);
  fe_to_be(out + 1 + cs, y, cs);
}
bool on_group(const WGroup& g, const Fe& xm, const Fe& ym) {   // weier.ts:56-70 with z = 1
  const Mod& m = *g.f;
  Fe l = mmul(ym, ym, m);
  Fe r = mmul(mmul(xm, xm, m), xm, m);
  Fe t = madd(madd(xm, xm, m), xm, m);
  r = madd(msub(r, t, m), g.b, m);
  return fe_eq(l, r);
}
// deserializePoint (weier.ts:74-89); the fixed slot gives the identity as 65 zero bytes
bool deserialize(const WGroup& g, const uint8_t* b, WPt& out, bool* inf = nullptr) {
  const int cs = g.f->nbytes;
  bool allz = true;
  for (int i = 0; i < 1 + 2 * cs; i++) allz = allz && b[i] == 0;
  if (inf) *inf = allz;
  if (allz) { out = identity(g); return true; }
  if (b[0] != 0x04) return false;
  Fe x = s_reduce(fe_from_be(b + 1, cs), *g.f), y = s_reduce(fe_from_be(b + 1 + cs, cs), *g.f);   // no range check in weier.ts
  Fe xm = to_mont(x, *g.f), ym = to_mont(y, *g.f);
  if (!on_group(g, xm, ym)) return false;
  out = WPt{xm, ym, g.f->one};
  return true;
}

// ---- twisted Edwards (edwards.ts:94-210)
inline EPt identity(const EGroup& g) { return EPt{fe_zero(), g.f->one, fe_zero(), g.f->one}; }
inline EPt generator(const EGroup& g) { return EPt{g.gx, g.gy, mmul(g.gx, g.gy, *g.f), g.f->one}; }
inline bool is_identity(const EGroup&, const EPt& p) {   // edwards.ts:117-125
  return fe_is_zero(p.x) && !fe_is_zero(p.y) && fe_is_zero(p.t) && !fe_is_zero(p.z) && fe_eq(p.y, p.z);
}
inline EPt neg(const EGroup& g, const EPt& p) { return EPt{mneg(p.x, *g.f), p.y, mneg(p.t, *g.f), p.z}; }
inline bool pt_eq(const EGroup& g, const EPt& a, const EPt& b) {
  const Mod& m = *g.f;
  return fe_eq(mmul(a.x, b.z, m), mmul(b.x, a.z, m)) && fe_eq(mmul(a.y, b.z, m), mmul(b.y, a.z, m));
}
EPt dbl(const EGroup& g, const EPt& p) {   // edwards.ts:141-160 (Hisil 3.3)
  const Mod& m = *g.f;
  Fe A = mmul(p.x, p.x, m), B = mmul(p.y, p.y, m), C = mmul(p.z, p.z, m);
  C = madd(C, C, m);
  Fe D = mmul(g.a, A, m);
  Fe EE = madd(p.x, p.y, m);
  Fe E = msub(msub(mmul(EE, EE, m), A, m), B, m);
  Fe G = madd(D, B, m), F = msub(G, C, m), H = msub(D, B, m);
  return EPt{mmul(E, F, m), mmul(G, H, m), mmul(E, H, m), mmul(F, G, m)};
}
EPt add(const EGroup& g, const EPt& p, const EPt& q) {   // edwards.ts:161-183 (Hisil 3.1, unified)
  const Mod& m = *g.f;
  Fe A = mmul(p.x, q.x, m), B = mmul(p.y, q.y, m);
  Fe C = mmul(mmul(g.d, p.t, m), q.t, m);
  Fe D = mmul(p.z, q.z, m);
  Fe E = msub(msub(mmul(madd(p.x, p.y, m), madd(q.x, q.y, m), m), A, m), B, m);
  Fe F = msub(D, C, m), G = madd(D, C, m), H = msub(B, mmul(g.a, A, m), m);
  return EPt{mmul(E, F, m), mmul(G, H, m), mmul(E, H, m), mmul(F, G, m)};
}
bool to_affine(const EGroup& g, const EPt& p, Fe& x, Fe& y) {   // edwards.ts:184-193
  const Mod& m = *g.f;
  Fe zi = minv(p.z, m);
  x = from_mont(mmul(p.x, zi, m), m);
  y = from_mont(mmul(p.y, zi, m), m);
  return true;
}
void to_bytes(const EGroup& g, const EPt& p, uint8_t* out) {   // edwards.ts:195-203
  Fe x, y;
  to_affine(g, p, x, y);
  const int cs = g.f->nbytes;
  out[0] = 0x04;
  fe_to_be(out + 1, x, cs);
  fe_to_be(out + 1 + cs, y, cs);
}
bool deserialize(const EGroup& g, const uint8_t* b, EPt& out, bool* inf = nullptr) {   // edwards.ts:70-86
  if (inf) *inf = false;
  const Mod& m = *g.f;
  const int cs = m.nbytes;
  if (b[0] != 0x04) return false;
  Fe x = fe_from_be(b + 1, cs), y = fe_from_be(b + 1 + cs, cs);
  if (fe_cmp(x, m.p) >= 0 || fe_cmp(y, m.p) >= 0) return false;
  Fe xm = to_mont(x, m), ym = to_mont(y, m);
  Fe x2 = mmul(xm, xm, m), y2 = mmul(ym, ym, m);
  Fe l = madd(mmul(g.a, x2, m), y2, m);
  Fe r = madd(m.one, mmul(g.d, mmul(x2, y2, m), m), m);
  if (!fe_eq(l, r)) return false;
  out = EPt{xm, ym, mmul(xm, ym, m), m.one};
  return true;
}

// ---- Point.mul / Point.dblmul (group.ts:97-152): hex digits MSB first, 16-entry tables per call
inline int hex_digits(const Fe& k) { int b = fe_bitlen(k); return b == 0 ? 1 : (b + 3) / 4; }
inline int hex_digit(const Fe& k, int i) { return (int)((k.v[i / 16] >> (4 * (i % 16))) & 15); }
template <class G, class P>
P pt_mul(const G& g, const P& base, const Fe& k) {
  P tab[16];
  P cur = identity(g);
  for (int d = 0; d < 16; d++) { tab[d] = cur; cur = add(g, cur, base); }
  P q = identity(g);
  for (int i = hex_digits(k) - 1; i >= 0; i--) {
    q = dbl(g, dbl(g, dbl(g, dbl(g, q))));
    q = add(g, q, tab[hex_digit(k, i)]);
  }
  return q;
}
template <class G, class P>
P pt_dblmul(const G& g, const P& p1, const Fe& k1, const P& p2, const Fe& k2) {
  P t1[16], t2[16];
  P c1 = identity(g), c2 = identity(g);
  for (int d = 0; d < 16; d++) { t1[d] = c1; t2[d] = c2; c1 = add(g, c1, p1); c2 = add(g, c2, p2); }
  P q = identity(g);
  for (int i = std::max(hex_digits(k1), hex_digits(k2)) - 1; i >= 0; i--) {
    q = dbl(g, dbl(g, dbl(g, dbl(g, q))));
    q = add(g, q, t1[hex_digit(k1, i)]);
    q = add(g, q, t2[hex_digit(k2, i)]);
  }
  return q;
}
template <class G> struct PtOf;
template <> struct PtOf<WGroup> { typedef WPt T; };
template <> struct PtOf<EGroup> { typedef EPt T; };
template <class G> inline int pt_bytes(const G& g) { return 1 + 2 * g.f->nbytes; }
template <class G> inline int sc_bytes(const G& g) { return g.f->nbytes; }   // Scalar.toBytes: FIELD size (group.ts:196-199)

// hashPoints (group.ts:221-233)
struct PointHasher {
  Sha256 s;
  PointHasher() { s.init(); }
  template <class G, class P> void pt(const G& g, const P& p) {
    uint8_t b[80];
    to_bytes(g, p, b);
    // weier.ts:247: the P-256 identity hashes as ONE zero byte
    if (b[0] == 0) s.update(b, 1); else s.update(b, pt_bytes(g));
  }
  Fe done() {
    uint8_t d[32];
    s.final(d);
    return fe_from_be(d, 10);
  }
};

// --------------------------------------------------------------------------- commitments (src/commit)
template <class G> struct Commit { typename PtOf<G>::T p; Fe r; };
template <class G> struct Pedersen {   // PedersenParams (pedersen.ts:40-59)
  const G* c;
  typename PtOf<G>::T g, h;
  Commit<G> commit(const Fe& v, Tape& t) const {   // pedersen.ts:53-58: h.dblmul(r, g, v)
    Fe r = t.rnd32(*c->q);
    return Commit<G>{pt_dblmul(*c, h, r, g, v), r};
  }
};
typedef Pedersen<EGroup> PedE;
typedef Commit<EGroup> ComE;
inline ComE c_add(const EGroup& g, const ComE& a, const ComE& b) { return ComE{add(g, a.p, b.p), s_add(a.r, b.r, *g.q)}; }
inline ComE c_sub(const EGroup& g, const ComE& a, const ComE& b) { return ComE{add(g, a.p, neg(g, b.p)), s_sub(a.r, b.r, *g.q)}; }

struct EqProof { EPt A1, A2; Fe tx, tr1, tr2; };
struct MultProof { EPt C4, Ax, Ay, Az, A41, A42; Fe tx, ty, tz, trx, try_, trz, tr4; };
struct PointAddProof { EPt C8, C10, C11, C13; MultProof pi8, pi10, pi11, pi13; EqProof pix, piy; };

EqProof prove_equality(const PedE& P, const Fe& x, const ComE& C1, const ComE& C2, Tape& t) {   // equality.ts:60-78
  const EGroup& g = *P.c;
  const Mod& q = *g.q;
  Fe k = t.rnd32(q);
  ComE A1 = P.commit(k, t), A2 = P.commit(k, t);
  PointHasher h;
  h.pt(g, C1.p); h.pt(g, C2.p); h.pt(g, A1.p); h.pt(g, A2.p);
  Fe c = h.done();
  return EqProof{A1.p, A2.p, s_sub(k, s_mul(c, x, q), q), s_sub(A1.r, s_mul(c, C1.r, q), q), s_sub(A2.r, s_mul(c, C2.r, q), q)};
}
MultProof prove_mult(const PedE& P, const Fe& x, const Fe& y, const Fe& z, const ComE& Cx, const ComE& Cy, const ComE& Cz,
                     Tape& t) {   // mult.ts:93-131
  const EGroup& g = *P.c;
  const Mod& q = *g.q;
  EPt C4 = pt_mul(g, Cy.p, x);
  Fe r4 = s_mul(Cy.r, x, q);
  Fe kx = t.rnd32(q), ky = t.rnd32(q), kz = t.rnd32(q);
  ComE Ax = P.commit(kx, t), Ay = P.commit(ky, t), Az = P.commit(kz, t), A41 = P.commit(kz, t);
  EPt A42 = pt_mul(g, Cy.p, kx);
  PointHasher h;
  h.pt(g, Cx.p); h.pt(g, Cy.p); h.pt(g, Cz.p); h.pt(g, C4); h.pt(g, Ax.p); h.pt(g, Ay.p); h.pt(g, Az.p); h.pt(g, A41.p); h.pt(g, A42);
  Fe c = h.done();
  MultProof pi;
  pi.C4 = C4; pi.Ax = Ax.p; pi.Ay = Ay.p; pi.Az = Az.p; pi.A41 = A41.p; pi.A42 = A42;
  pi.tx = s_sub(kx, s_mul(c, x, q), q);
  pi.ty = s_sub(ky, s_mul(c, y, q), q);
  pi.tz = s_sub(kz, s_mul(c, z, q), q);
  pi.trx = s_sub(Ax.r, s_mul(c, Cx.r, q), q);
  pi.try_ = s_sub(Ay.r, s_mul(c, Cy.r, q), q);
  pi.trz = s_sub(Az.r, s_mul(c, Cz.r, q), q);
  pi.tr4 = s_sub(A41.r, s_mul(c, r4, q), q);
  return pi;
}

// ------------------------------------------------------------ Relation / MultiMult (multimult.ts)
template <class G> struct MultiMult {
  typedef typename PtOf<G>::T P;
  struct Pair { P pt; Fe s; };
  const G* g;
  std::vector<Pair> pairs;
  std::vector<std::pair<P, int>> known;
  explicit MultiMult(const G& gg) : g(&gg) {}
  void add_known(const P& pt) {   // :42-48
    for (auto& k : known)
      if (pt_eq(*g, pt, k.first)) return;
    pairs.push_back(Pair{pt, fe_zero()});
    known.push_back({pt, (int)pairs.size() - 1});
  }
  void insert(const P& pt, const Fe& s) {   // :50-59
    for (auto& k : known)
      if (pt_eq(*g, pt, k.first)) { pairs[k.second].s = s_add(pairs[k.second].s, s, *g->q); return; }
    pairs.push_back(Pair{pt, s});
  }
  static bool less(const Pair& a, const Pair& b) { return fe_cmp(a.s, b.s) < 0; }
  void bubbleup(size_t index) {   // 1-based (:111-123)
    while (index > 1) {
      size_t parent = index / 2;
      if (less(pairs[parent - 1], pairs[index - 1])) { std::swap(pairs[parent - 1], pairs[index - 1]); index = parent; }
      else return;
    }
  }
  void pushdown(size_t parent) {   // :125-145
    for (;;) {
      size_t son = 2 * parent, daughter = son + 1;
      if (son > pairs.size()) return;
      size_t child = son;
      if (daughter <= pairs.size() && less(pairs[son - 1], pairs[daughter - 1])) child = daughter;
      if (less(pairs[parent - 1], pairs[child - 1])) { std::swap(pairs[parent - 1], pairs[child - 1]); parent = child; }
      else return;
    }
  }
  P evaluate() {   // Bos-Coster (:61-89)
    if (pairs.empty()) return identity(*g);
    if (pairs.size() == 1) return pt_mul(*g, pairs[0].pt, pairs[0].s);
    for (size_t i = 0; i < pairs.size(); i++) bubbleup(i + 1);
    for (;;) {
      if (pairs.size() == 1) return pt_mul(*g, pairs[0].pt, pairs[0].s);
      std::swap(pairs[0], pairs.back());
      Pair a = pairs.back();
      pairs.pop_back();
      pushdown(1);
      Pair& b = pairs[0];
      if (fe_is_zero(b.s)) return pt_mul(*g, a.pt, a.s);
      Pair c{a.pt, s_sub(a.s, b.s, *g->q)};
      b.pt = add(*g, b.pt, a.pt);
      if (!fe_is_zero(c.s)) { pairs.push_back(c); bubbleup(pairs.size()); }
    }
  }
};
template <class G> struct Relation {   // :147-174
  typedef typename PtOf<G>::T P;
  const G* g;
  std::vector<typename MultiMult<G>::Pair> pairs;
  explicit Relation(const G& gg) : g(&gg) {}
  void insert(const P& pt, const Fe& s) { pairs.push_back({pt, s}); }
  void drain(MultiMult<G>& m, Tape& t) {
    Fe rz = t.rnd32(*g->q);
    for (auto& pr : pairs) m.insert(pr.pt, s_mul(pr.s, rz, *g->q));
  }
};

void aggregate_equality(const PedE& P, const EPt& C1, const EPt& C2, const EqProof& pi, MultiMult<EGroup>& multi, Tape& t) {   // equality.ts:94-116
  const EGroup& g = *P.c;
  PointHasher h;
  h.pt(g, C1); h.pt(g, C2); h.pt(g, pi.A1); h.pt(g, pi.A2);
  Fe c = h.done(), one = s_small(1);
  Relation<EGroup> r1(g), r2(g);
  r1.insert(P.g, pi.tx); r1.insert(P.h, pi.tr1); r1.insert(C1, c); r1.insert(neg(g, pi.A1), one);
  r2.insert(P.g, pi.tx); r2.insert(P.h, pi.tr2); r2.insert(C2, c); r2.insert(neg(g, pi.A2), one);
  r1.drain(multi, t);
  r2.drain(multi, t);
}
void aggregate_mult(const PedE& P, const EPt& Cx, const EPt& Cy, const EPt& Cz, const MultProof& pi, MultiMult<EGroup>& multi,
                    Tape& t) {   // mult.ts:148-175
  const EGroup& g = *P.c;
  PointHasher h;
  h.pt(g, Cx); h.pt(g, Cy); h.pt(g, Cz); h.pt(g, pi.C4); h.pt(g, pi.Ax); h.pt(g, pi.Ay); h.pt(g, pi.Az); h.pt(g, pi.A41); h.pt(g, pi.A42);
  Fe c = h.done(), one = s_small(1);
  Relation<EGroup> rx(g), ry(g), rz(g), r41(g), r42(g);
  rx.insert(P.g, pi.tx); rx.insert(P.h, pi.trx); rx.insert(Cx, c); rx.insert(neg(g, pi.Ax), one);
  ry.insert(P.g, pi.ty); ry.insert(P.h, pi.try_); ry.insert(Cy, c); ry.insert(neg(g, pi.Ay), one);
  rz.insert(P.g, pi.tz); rz.insert(P.h, pi.trz); rz.insert(Cz, c); rz.insert(neg(g, pi.Az), one);
  r41.insert(P.g, pi.tz); r41.insert(P.h, pi.tr4); r41.insert(pi.C4, c); r41.insert(neg(g, pi.A41), one);
  r42.insert(Cy, pi.tx); r42.insert(pi.C4, c); r42.insert(neg(g, pi.A42), one);
  rx.drain(multi, t); ry.drain(multi, t); rz.drain(multi, t); r41.drain(multi, t); r42.drain(multi, t);
}

// ------------------------------------------------------------------------ pointAdd.ts / exp.ts
PointAddProof prove_point_add(const PedE& P, const WPt& Pp, const WPt& Q, const WPt& R, const ComE& PX, const ComE& PY,
                              const ComE& QX, const ComE& QY, const ComE& RX, const ComE& RY, Tape& t) {   // pointAdd.ts:92-163
  const EGroup& g = *P.c;
  const Mod& q = *g.q;
  if (!pt_eq(P256, add(P256, Pp, Q), R)) throw ZkErr{ZKA_ERR_POINTS_DONT_ADD};
  const ComE &C1 = PX, &C2 = QX, &C3 = RX, &C4 = PY, &C5 = QY, &C6 = RY;
  Fe x1, y1, x2, y2, x3, y3;
  if (!to_affine(P256, Pp, x1, y1)) throw ZkErr{ZKA_ERR_T1_INFINITY};
  if (!to_affine(P256, Q, x2, y2)) throw ZkErr{ZKA_ERR_INVALID_PK};
  if (!to_affine(P256, R, x3, y3)) throw ZkErr{ZKA_ERR_T_INFINITY};
  // coordinates are integers < p256.p = tom.order: already canonical scalars of the proof group
  Fe i7 = s_sub(x2, x1, q), i8 = s_inv(i7, q), i9 = s_sub(y2, y1, q), i10 = s_mul(i8, i9, q), i11 = s_mul(i10, i10, q);
  Fe i12 = s_sub(x1, x3, q), i13 = s_mul(i10, i12, q);
  ComE C7 = c_sub(g, C2, C1);
  ComE C8 = P.commit(i8, t);
  ComE C9 = c_sub(g, C5, C4);
  ComE C10 = P.commit(i10, t);
  ComE C11 = P.commit(i11, t);
  ComE C12 = c_sub(g, C1, C3);
  ComE C13 = P.commit(i13, t);
  ComE C14{P.g, fe_zero()};
  PointAddProof pi;
  pi.C8 = C8.p; pi.C10 = C10.p; pi.C11 = C11.p; pi.C13 = C13.p;
  pi.pi8 = prove_mult(P, i7, i8, s_small(1), C7, C8, C14, t);
  pi.pi10 = prove_mult(P, i8, i9, i10, C8, C9, C10, t);
  pi.pi11 = prove_mult(P, i10, i10, i11, C10, C10, C11, t);
  ComE Cint = c_add(g, c_add(g, C3, C1), C2);
  pi.pix = prove_equality(P, i11, C11, Cint, t);
  pi.pi13 = prove_mult(P, i10, i12, i13, C10, C12, C13, t);
  Cint = c_add(g, C6, C4);
  pi.piy = prove_equality(P, i13, C13, Cint, t);
  return pi;
}
void aggregate_point_add(const PedE& P, const EPt& PX, const EPt& PY, const EPt& QX, const EPt& QY, const EPt& RX, const EPt& RY,
                         const PointAddProof& pi, MultiMult<EGroup>& multi, Tape& t) {   // pointAdd.ts:199-259
  const EGroup& g = *P.c;
  const EPt &C1 = PX, &C2 = QX, &C3 = RX, &C4 = PY, &C5 = QY, &C6 = RY;
  EPt C7 = add(g, C2, neg(g, C1)), C9 = add(g, C5, neg(g, C4)), C12 = add(g, C1, neg(g, C3));
  aggregate_mult(P, C7, pi.C8, P.g, pi.pi8, multi, t);
  aggregate_mult(P, pi.C8, C9, pi.C10, pi.pi10, multi, t);
  aggregate_mult(P, pi.C10, pi.C10, pi.C11, pi.pi11, multi, t);
  EPt Cint = add(g, add(g, C3, C1), C2);
  aggregate_equality(P, pi.C11, Cint, pi.pix, multi, t);
  aggregate_mult(P, pi.C10, C12, pi.C13, pi.pi13, multi, t);
  Cint = add(g, C4, C6);
  aggregate_equality(P, pi.C13, Cint, pi.piy, multi, t);
}

struct ExpRep {   // ExpProof (exp.ts:26-84)
  int tag;        // 1: alpha beta1 beta2 beta3; 0: z z2 proof r1 r2
  WPt A;
  EPt Tx, Ty;
  Fe s0, s1, s2, s3;   // tag 1: alpha beta1 beta2 beta3 ; tag 0: z z2 r1 r2
  PointAddProof pa;
};
typedef Pedersen<WGroup> PedN;
typedef Commit<WGroup> ComN;

std::vector<ExpRep> prove_exp(const PedN& N, const PedE& W, const Fe& s, const ComN& Cs, const WPt& Pp, const ComE& Px,
                              const ComE& Py, int secparam, Tape& t, const WPt* Q) {   // exp.ts:126-231
  const WGroup& gn = *N.c;
  const Mod& qn = *gn.q;
  std::vector<Fe> alpha(secparam), r(secparam);
  std::vector<WPt> T(secparam), A(secparam);
  std::vector<ComE> Tx(secparam), Ty(secparam);
  for (int i = 0; i < secparam; i++) {
    alpha[i] = t.rnd32(qn);
    r[i] = t.rnd32(qn);
    T[i] = pt_mul(gn, N.g, alpha[i]);
    A[i] = add(gn, T[i], pt_mul(gn, N.h, r[i]));
    Fe x, y;
    if (!to_affine(gn, T[i], x, y)) throw ZkErr{ZKA_ERR_T_INFINITY};
    Tx[i] = W.commit(x, t);
    Ty[i] = W.commit(y, t);
  }
  PointHasher h;
  h.pt(*W.c, Px.p); h.pt(*W.c, Py.p);
  for (int i = 0; i < secparam; i++) { h.pt(gn, A[i]); h.pt(*W.c, Tx[i].p); h.pt(*W.c, Ty[i].p); }
  Fe challenge = h.done();
  std::vector<ExpRep> out(secparam);
  for (int i = 0; i < secparam; i++) {
    ExpRep& e = out[i];
    e.A = A[i]; e.Tx = Tx[i].p; e.Ty = Ty[i].p;
    const bool bit = (challenge.v[i / 64] >> (i % 64)) & 1;   // isOdd(challenge); challenge >>= 1
    if (bit) {
      e.tag = 1;
      e.s0 = alpha[i]; e.s1 = r[i]; e.s2 = Tx[i].r; e.s3 = Ty[i].r;
    } else {
      e.tag = 0;
      Fe z = s_sub(alpha[i], s, qn);
      WPt T1 = pt_mul(gn, N.g, z);
      if (Q) T1 = add(gn, T1, *Q);
      Fe x, y;
      if (!to_affine(gn, T1, x, y)) throw ZkErr{ZKA_ERR_T1_INFINITY};
      ComE T1x = W.commit(x, t), T1y = W.commit(y, t);
      e.pa = prove_point_add(W, T1, Pp, T[i], T1x, T1y, Px, Py, Tx[i], Ty[i], t);
      e.s0 = z; e.s1 = s_sub(r[i], Cs.r, qn); e.s2 = T1x.r; e.s3 = T1y.r;
    }
  }
  return out;
}

// ------------------------------------------------------------------------------ gk.ts / interpolate.ts
struct GkProof { std::vector<EPt> cl, ca, cb, cd; std::vector<Fe> f, za, zb; Fe zd; };
inline int ceil_log2(uint32_t v) { int n = 0; while ((1ull << n) < v) n++; return n; }
inline EPt gk_commit(const PedE& P, const Fe& val, const Fe& blinder) { return pt_dblmul(*P.c, P.g, val, P.h, blinder); }   // gk.ts:88-92

std::vector<Fe> interpolate(const std::vector<Fe>& x, const std::vector<Fe>& y, const Mod& q) {   // interpolate.ts:27-70
  const int n = (int)x.size();
  std::vector<Fe> s(n + 1, fe_zero()), coeff(n, fe_zero());
  if (n == 0) return coeff;
  s[n] = s_small(1);
  s[n - 1] = s_neg(x[0], q);
  for (int i = 1; i < n; i++) {
    for (int j = n - i - 1; j < n - 1; j++) s[j] = s_sub(s[j], s_mul(x[i], s[j + 1], q), q);
    s[n - 1] = s_sub(s[n - 1], x[i], q);
  }
  for (int i = 0; i < n; i++) {
    Fe phi = fe_zero();
    for (int j = n; j >= 1; j--) phi = s_add(s_mul(s_small((u64)j), s[j], q), s_mul(x[i], phi, q), q);
    Fe ff = s_inv(phi, q), b = s_small(1);
    for (int j = n - 1; j >= 0; j--) {
      coeff[j] = s_add(coeff[j], s_mul(s_mul(b, ff, q), y[i], q), q);
      b = s_add(s[j], s_mul(x[i], b, q), q);
    }
  }
  return coeff;
}

GkProof prove_membership(const PedE& P, const ComE& com, uint32_t index, const std::vector<Fe>& initial, Tape& t) {   // gk.ts:94-195
  const EGroup& g = *P.c;
  const Mod& q = *g.q;
  const int n = ceil_log2((uint32_t)initial.size());
  std::vector<Fe> values = initial;   // pad (gk.ts:75-86)
  values.resize((size_t)1 << n, initial[0]);
  std::vector<int> eli(n);
  for (int i = 0; i < n; i++) eli[i] = (index >> i) & 1;
  std::vector<Fe> ri(n), ai(n), si(n), ti(n), rho(n);
  for (int i = 0; i < n; i++) { ri[i] = t.rnd32(q); ai[i] = t.rnd32(q); si[i] = t.rnd32(q); ti[i] = t.rnd32(q); rho[i] = t.rnd32(q); }
  GkProof pr;
  for (int i = 0; i < n; i++) {
    pr.cl.push_back(gk_commit(P, s_small((u64)eli[i]), ri[i]));
    pr.ca.push_back(gk_commit(P, ai[i], si[i]));
    pr.cb.push_back(gk_commit(P, eli[i] ? ai[i] : fe_zero(), ti[i]));
  }
  std::vector<Fe> omegas(n), dv(n);
  std::vector<Fe> p;
  for (int w = 0; w < n; w++) {
    omegas[w] = s_small((u64)w);
    std::vector<Fe> f0(n), ratio(n);
    Fe prod = s_small(1);
    for (int j = 0; j < n; j++) {
      Fe wj = s_small((u64)w);
      f0[j] = eli[j] ? s_neg(ai[j], q) : s_sub(wj, ai[j], q);
      Fe f1 = eli[j] ? s_add(wj, ai[j], q) : ai[j];
      ratio[j] = s_mul(f1, s_inv(f0[j], q), q);
      prod = s_mul(prod, f0[j], q);
    }
    p.assign(1, prod);
    for (int i = 0; i < n; i++) {
      const size_t old = p.size();
      for (size_t j = 0; j < old; j++) p.push_back(s_mul(ratio[i], p[j], q));
    }
    Fe dval = fe_zero();
    for (size_t i = 0; i < values.size(); i++) dval = s_add(dval, s_mul(s_sub(values[index], values[i], q), p[i], q), q);
    dv[w] = dval;
  }
  std::vector<Fe> di = interpolate(omegas, dv, q);
  for (int i = 0; i < n; i++) pr.cd.push_back(gk_commit(P, di[i], rho[i]));
  PointHasher h;
  for (auto* arr : {&pr.cl, &pr.ca, &pr.cb, &pr.cd})
    for (auto& pt : *arr) h.pt(g, pt);
  Fe x = h.done();
  Fe zd = s_mul(com.r, s_pow_small(x, n, q), q);
  for (int i = 0; i < n; i++) {
    Fe f = s_add(eli[i] ? x : fe_zero(), ai[i], q);
    pr.f.push_back(f);
    pr.za.push_back(s_add(s_mul(ri[i], x, q), si[i], q));
    pr.zb.push_back(s_add(s_mul(ri[i], s_sub(x, f, q), q), ti[i], q));
  }
  for (int i = 0; i < n; i++) zd = s_sub(zd, s_mul(rho[i], s_pow_small(x, i, q), q), q);
  pr.zd = zd;
  return pr;
}

bool verify_membership(const PedE& P, const EPt& com, const std::vector<Fe>& init_vec, const GkProof& pr, Tape& t) {   // gk.ts:197-262
  const EGroup& g = *P.c;
  const Mod& q = *g.q;
  const int n = ceil_log2((uint32_t)init_vec.size());
  std::vector<Fe> vec = init_vec;
  vec.resize((size_t)1 << n, init_vec[0]);
  for (size_t l : {pr.cl.size(), pr.ca.size(), pr.cb.size(), pr.cd.size(), pr.f.size(), pr.za.size(), pr.zb.size()})
    if ((size_t)n != l) return false;
  MultiMult<EGroup> multi(g);
  PointHasher h;
  for (auto* arr : {&pr.cl, &pr.ca, &pr.cb, &pr.cd})
    for (auto& pt : *arr) h.pt(g, pt);
  Fe x = h.done();
  multi.add_known(P.g);
  multi.add_known(P.h);
  const Fe one = s_small(1);
  for (int i = 0; i < n; i++) {
    Relation<EGroup> r0(g), r1(g);
    r0.insert(pr.cl[i], x); r0.insert(pr.ca[i], one); r0.insert(P.g, s_neg(pr.f[i], q)); r0.insert(P.h, s_neg(pr.za[i], q));
    r0.drain(multi, t);
    r1.insert(pr.cl[i], s_sub(x, pr.f[i], q)); r1.insert(pr.cb[i], one); r1.insert(P.h, s_neg(pr.zb[i], q));
    r1.drain(multi, t);
  }
  Fe total = fe_zero();
  for (size_t i = 0; i < vec.size(); i++) {
    Fe pix = s_small(1);
    for (int j = 0; j < n; j++) pix = s_mul(pix, (i >> j) & 1 ? pr.f[j] : s_sub(x, pr.f[j], q), q);
    total = s_add(total, s_mul(vec[i], pix, q), q);
  }
  Relation<EGroup> rf(g);
  for (int i = 0; i < n; i++) rf.insert(pr.cd[i], s_neg(s_pow_small(x, i, q), q));
  rf.insert(com, s_pow_small(x, n, q));
  rf.insert(P.g, s_neg(total, q));
  rf.insert(P.h, s_neg(pr.zd, q));
  rf.drain(multi, t);
  return is_identity(g, multi.evaluate());
}

// generateIndices (exp.ts:95-109) with the C-ABI verifier tape: byte i is rnd(limit - i), pre-filtered
std::vector<int> generate_indices(int limit, const uint8_t* idx_bytes) {
  std::vector<int> ret(limit);
  for (int i = 0; i < limit; i++) ret[i] = i;
  for (int i = 0; i < limit - 2; i++) {
    int r = idx_bytes[i];
    if (r >= limit - i) throw ZkErr{ZKA_ERR_TAPE_RANGE};
    std::swap(ret[i], ret[r + i]);
  }
  return ret;
}

bool verify_exp(const PedN& N, const PedE& W, const WPt& Clambda, const EPt& Px, const EPt& Py, const std::vector<ExpRep>& pi,
                int secparam, const uint8_t* idx_bytes, Tape& t, const WPt* Q) {   // exp.ts:233-349
  if (secparam > (int)pi.size()) throw ZkErr{ZKA_E_ARG};
  const WGroup& gn = *N.c;
  const EGroup& gw = *W.c;
  MultiMult<EGroup> multiW(gw);
  MultiMult<WGroup> multiN(gn);
  multiW.add_known(W.g); multiW.add_known(W.h);
  multiN.add_known(N.g); multiN.add_known(N.h); multiN.add_known(Clambda);
  PointHasher h;
  h.pt(gw, Px); h.pt(gw, Py);
  for (auto& e : pi) { h.pt(gn, e.A); h.pt(gw, e.Tx); h.pt(gw, e.Ty); }
  Fe challenge = h.done();
  std::vector<int> indices = generate_indices((int)pi.size(), idx_bytes);
  const Fe one = s_small(1);
  for (int j = 0; j < secparam; j++) {
    const int i = indices[j];
    const ExpRep& e = pi[i];
    const bool bit = (challenge.v[i / 64] >> (i % 64)) & 1;
    if (bit) {
      if (e.tag != 1) throw ZkErr{ZKA_ERR_PARAMS_NOT_FOUND};
      WPt T = pt_mul(gn, N.g, e.s0);
      Relation<WGroup> relA(gn);
      relA.insert(T, one); relA.insert(N.h, e.s1); relA.insert(neg(gn, e.A), one);
      relA.drain(multiN, t);
      Fe sx, sy;
      if (!to_affine(gn, T, sx, sy)) throw ZkErr{ZKA_ERR_T_INFINITY};
      Relation<EGroup> relTx(gw), relTy(gw);
      relTx.insert(W.g, sx); relTx.insert(W.h, e.s2); relTx.insert(neg(gw, e.Tx), one);
      relTy.insert(W.g, sy); relTy.insert(W.h, e.s3); relTy.insert(neg(gw, e.Ty), one);
      relTx.drain(multiW, t);
      relTy.drain(multiW, t);
    } else {
      if (e.tag != 0) throw ZkErr{ZKA_ERR_PARAMS_NOT_FOUND};
      WPt T1 = pt_mul(gn, N.g, e.s0);
      Relation<WGroup> relA(gn);
      relA.insert(T1, one); relA.insert(Clambda, one); relA.insert(neg(gn, e.A), one); relA.insert(N.h, e.s1);
      relA.drain(multiN, t);
      if (Q) T1 = add(gn, T1, *Q);
      Fe sx, sy;
      if (!to_affine(gn, T1, sx, sy)) throw ZkErr{ZKA_ERR_T1_INFINITY};
      EPt T1x = pt_dblmul(gw, W.g, sx, W.h, e.s2), T1y = pt_dblmul(gw, W.g, sy, W.h, e.s3);
      aggregate_point_add(W, T1x, T1y, Px, Py, e.Tx, e.Ty, e.pa, multiW, t);
    }
  }
  const bool okW = is_identity(gw, multiW.evaluate());
  const bool okN = is_identity(gn, multiN.evaluate());
  return okW && okN;
}

// ------------------------------------------------------------------ flat layout (include/zkattest.h)
enum { NP = 65, WP = 67, NS = 32, WS = 33, EQ_LEN = 2 * WP + 3 * WS, MULT_LEN = 6 * WP + 7 * WS,
       PA_LEN = 4 * WP + 4 * MULT_LEN + 2 * EQ_LEN, REP_HEAD = 1 + NP + 2 * WP, REP1_LEN = REP_HEAD + 2 * NS + 2 * WS,
       REP0_LEN = REP_HEAD + 2 * NS + PA_LEN + 2 * WS, HEAD_LEN = 2 * NP + 2 * WP };
inline size_t gk_len(int n) { return 1 + (size_t)4 * n * WP + (size_t)(3 * n + 1) * WS; }
inline size_t proof_len(int z, int n, int reps) { return HEAD_LEN + (size_t)z * REP0_LEN + (size_t)(reps - z) * REP1_LEN + gk_len(n); }

struct Wr {
  uint8_t* o;
  void npt(const WPt& p) { to_bytes(P256, p, o); o += NP; }
  void wpt(const EPt& p) { to_bytes(TOM, p, o); o += WP; }
  void nsc(const Fe& s) { fe_to_be(o, s, NS); o += NS; }
  void wsc(const Fe& s) { fe_to_be(o, s, WS); o += WS; }
  void eq(const EqProof& p) { wpt(p.A1); wpt(p.A2); wsc(p.tx); wsc(p.tr1); wsc(p.tr2); }
  void mult(const MultProof& p) {
    wpt(p.C4); wpt(p.Ax); wpt(p.Ay); wpt(p.Az); wpt(p.A41); wpt(p.A42);
    wsc(p.tx); wsc(p.ty); wsc(p.tz); wsc(p.trx); wsc(p.try_); wsc(p.trz); wsc(p.tr4);
  }
  void pa(const PointAddProof& p) {
    wpt(p.C8); wpt(p.C10); wpt(p.C11); wpt(p.C13);
    mult(p.pi8); mult(p.pi10); mult(p.pi11); mult(p.pi13); eq(p.pix); eq(p.piy);
  }
};
struct Rd {   // deserializePoint / deserializeScalar semantics; any failure -> MALFORMED
  const uint8_t* b;
  size_t len, o;
  const uint8_t* take(size_t n) {
    if (o + n > len) throw ZkErr{ZKA_ERR_MALFORMED};
    const uint8_t* p = b + o;
    o += n;
    return p;
  }
  WPt npt() { WPt p; if (!deserialize(P256, take(NP), p)) throw ZkErr{ZKA_ERR_MALFORMED}; return p; }
  EPt wpt() { EPt p; if (!deserialize(TOM, take(WP), p)) throw ZkErr{ZKA_ERR_MALFORMED}; return p; }
  Fe nsc() { Fe s = fe_from_be(take(NS), NS); if (fe_cmp(s, P256N.p) >= 0) throw ZkErr{ZKA_ERR_MALFORMED}; return s; }
  Fe wsc() { Fe s = fe_from_be(take(WS), WS); if (fe_cmp(s, P256P.p) >= 0) throw ZkErr{ZKA_ERR_MALFORMED}; return s; }
  EqProof eq() { EqProof p; p.A1 = wpt(); p.A2 = wpt(); p.tx = wsc(); p.tr1 = wsc(); p.tr2 = wsc(); return p; }
  MultProof mult() {
    MultProof p;
    p.C4 = wpt(); p.Ax = wpt(); p.Ay = wpt(); p.Az = wpt(); p.A41 = wpt(); p.A42 = wpt();
    p.tx = wsc(); p.ty = wsc(); p.tz = wsc(); p.trx = wsc(); p.try_ = wsc(); p.trz = wsc(); p.tr4 = wsc();
    return p;
  }
  PointAddProof pa() {
    PointAddProof p;
    p.C8 = wpt(); p.C10 = wpt(); p.C11 = wpt(); p.C13 = wpt();
    p.pi8 = mult(); p.pi10 = mult(); p.pi11 = mult(); p.pi13 = mult(); p.pix = eq(); p.piy = eq();
    return p;
  }
};

}  // namespace

// ============================================================================================ C ABI
struct zka_ctx {
  std::string err;
  int threads = 1;
};
struct zka_params {
  uint32_t sec_level;
  WPt h_nist;
  EPt h_proof;
};

namespace {

// proveSignatureList (zkpAttestList.ts:104-145) -> flat bytes; returns the per-proof status
int prove_one(const zka_params* P, const uint8_t* msg_hash, const uint8_t* sig, const uint8_t* pk, uint32_t which,
              const std::vector<Fe>& keys, const uint8_t* tape, size_t tape_len, uint8_t* out, uint32_t* out_len) {
  *out_len = 0;
  try {
    const Mod& n = P256N;
    WPt pkp;
    bool inf = false;
    if (pk[0] != 0x04 || !deserialize(P256, pk, pkp, &inf) || inf) throw ZkErr{ZKA_ERR_INVALID_PK};
    Fe pkx, pky;
    to_affine(P256, pkp, pkx, pky);
    // truncateToN is the identity for a 32-byte hash (:80-86); everything is reduced mod n by newScalar / posMod
    Fe z = s_reduce(fe_from_be(msg_hash, 32), n), r = s_reduce(fe_from_be(sig, 32), n), s = s_reduce(fe_from_be(sig + 32, 32), n);
    Fe sinv = s_inv(s, n), u1 = s_mul(sinv, z, n), u2 = s_mul(sinv, r, n);
    WPt G = generator(P256);
    WPt R = add(P256, pt_mul(P256, G, u1), pt_mul(P256, pkp, u2));
    Fe rinv = s_inv(r, n), s1 = s_mul(rinv, s, n), z1 = s_mul(rinv, z, n);
    WPt Q = pt_mul(P256, G, z1);
    if (which >= keys.size()) throw ZkErr{ZKA_ERR_BAD_INDEX};
    Tape t{tape, tape_len, 0};
    PedN sigexp{&P256, R, P->h_nist};
    PedE W{&TOM, generator(TOM), P->h_proof};
    ComN comS1 = sigexp.commit(s1, t);
    ComE pkX = W.commit(pkx, t), pkY = W.commit(pky, t);
    const int S = (int)P->sec_level;
    std::vector<ExpRep> reps = prove_exp(sigexp, W, s1, comS1, pkp, pkX, pkY, S, t, &Q);
    GkProof gk = prove_membership(W, pkX, which, keys, t);
    // serialise; a P-256 identity in a point slot cannot be encoded (include/zkattest.h ZKA_ERR_IDENTITY_ENC)
    Wr w{out};
    if (is_identity(P256, comS1.p)) throw ZkErr{ZKA_ERR_IDENTITY_ENC};
    w.npt(R); w.npt(comS1.p); w.wpt(pkX.p); w.wpt(pkY.p);
    for (auto& e : reps) {
      if (is_identity(P256, e.A)) throw ZkErr{ZKA_ERR_IDENTITY_ENC};
      *w.o++ = (uint8_t)e.tag;
      w.npt(e.A); w.wpt(e.Tx); w.wpt(e.Ty);
      if (e.tag) { w.nsc(e.s0); w.nsc(e.s1); w.wsc(e.s2); w.wsc(e.s3); }
      else { w.nsc(e.s0); w.nsc(e.s1); w.pa(e.pa); w.wsc(e.s2); w.wsc(e.s3); }
    }
    const int nn = (int)gk.cl.size();
    *w.o++ = (uint8_t)nn;
    for (auto* arr : {&gk.cl, &gk.ca, &gk.cb, &gk.cd})
      for (auto& p : *arr) w.wpt(p);
    for (auto* arr : {&gk.f, &gk.za, &gk.zb})
      for (auto& sc : *arr) w.wsc(sc);
    w.wsc(gk.zd);
    *out_len = (uint32_t)(w.o - out);
    return ZKA_OK;
  } catch (const ZkErr& e) {
    return e.status;
  }
}

// verifySignatureList (zkpAttestList.ts:147-184); verifier tape layout of include/zkattest.h / zk_verify.cuh
int verify_one(const zka_params* P, const uint8_t* msg_hash, const std::vector<Fe>& keys, const uint8_t* proof, size_t len,
               const uint8_t* tape, size_t tape_len, uint8_t* ok, int samples) {
  *ok = 0;
  const int S = (int)P->sec_level;
  const int n = ceil_log2((uint32_t)keys.size());
  try {
    // --- readJson-equivalent: every point / scalar is validated while parsing
    Rd r{proof, len, 0};
    WPt R;
    bool rinf = false;
    if (!deserialize(P256, r.take(NP), R, &rinf)) throw ZkErr{ZKA_ERR_MALFORMED};
    WPt comS1 = r.npt();
    EPt kx = r.wpt(), ky = r.wpt();
    std::vector<ExpRep> reps(S);
    for (int i = 0; i < S; i++) {
      ExpRep& e = reps[i];
      const uint8_t tag = *r.take(1);
      if (tag > 1) throw ZkErr{ZKA_ERR_MALFORMED};
      e.tag = tag;
      e.A = r.npt(); e.Tx = r.wpt(); e.Ty = r.wpt();
      e.s0 = r.nsc(); e.s1 = r.nsc();
      if (tag) { e.s2 = r.wsc(); e.s3 = r.wsc(); }
      else { e.pa = r.pa(); e.s2 = r.wsc(); e.s3 = r.wsc(); }
    }
    const int ngk = *r.take(1);
    if (r.o + gk_len(ngk) - 1 != len) throw ZkErr{ZKA_ERR_MALFORMED};
    GkProof gk;
    for (auto* arr : {&gk.cl, &gk.ca, &gk.cb, &gk.cd})
      for (int i = 0; i < ngk; i++) arr->push_back(r.wpt());
    for (auto* arr : {&gk.f, &gk.za, &gk.zb})
      for (int i = 0; i < ngk; i++) arr->push_back(r.wsc());
    gk.zd = r.wsc();
    if (rinf) throw ZkErr{ZKA_ERR_R_INFINITY};   // zkpAttestList.ts:158-160
    // --- statement
    const Mod& nn = P256N;
    Fe z = s_reduce(fe_from_be(msg_hash, 32), nn);
    Fe rx, ry;
    to_affine(P256, R, rx, ry);
    Fe rinv = s_inv(s_reduce(rx, nn), nn), z1 = s_mul(rinv, z, nn);
    WPt Q = pt_mul(P256, generator(P256), z1);
    PedN sigexp{&P256, R, P->h_nist};
    PedE W{&TOM, generator(TOM), P->h_proof};
    const size_t gbytes = (size_t)32 * (2 * n + 1);
    if (tape_len < gbytes + 96) throw ZkErr{ZKA_ERR_TAPE_RANGE};
    Tape tg{tape, gbytes, 0};
    if (!verify_membership(W, kx, keys, gk, tg)) return ZKA_OK;   // false before verifyExp can throw
    Tape te{tape + gbytes + 96, tape_len - gbytes - 96, 0};
    const bool okv = verify_exp(sigexp, W, comS1, kx, ky, reps, samples, tape + gbytes, te, &Q);
    *ok = okv ? 1 : 0;
    return ZKA_OK;
  } catch (const ZkErr& e) {
    return e.status;
  }
}

template <class F>
void parallel_for(int threads, uint32_t count, F fn) {
  if (threads <= 1 || count <= 1) {
    for (uint32_t i = 0; i < count; i++) fn(i);
    return;
  }
  std::atomic<uint32_t> next(0);
  std::vector<std::thread> th;
  const int nt = (int)std::min<uint32_t>((uint32_t)threads, count);
  for (int k = 0; k < nt; k++)
    th.emplace_back([&] {
      for (;;) {
        uint32_t i = next.fetch_add(1);
        if (i >= count) return;
        fn(i);
      }
    });
  for (auto& t : th) t.join();
}

std::vector<Fe> ring_scalars(const uint8_t* ring, uint32_t N) {   // pad() wraps each key in newScalar (gk.ts:77)
  std::vector<Fe> k(N);
  for (uint32_t i = 0; i < N; i++) k[i] = s_reduce(fe_from_be(ring + (size_t)i * 32, 32), P256P);
  return k;
}

}  // namespace

extern "C" {

int zka_version(void) { return 1; }
const char* zka_last_error(const zka_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }
uint64_t zka_launch_count(const zka_ctx*) { return 0; }
void* zka_get_stream(zka_ctx*) { return nullptr; }
int zka_set_profiling(zka_ctx*, int) { return 0; }
int zka_profile_reset(zka_ctx*) { return 0; }
size_t zka_profile_json(zka_ctx*, char* buf, size_t cap) {
  if (buf && cap >= 3) strcpy(buf, "{}");
  return 3;
}
int zka_config(const zka_ctx*, int* w, int* nw, int* ch) {
  if (w) *w = 4;
  if (nw) *nw = 64;
  if (ch) *ch = 1;
  return 0;
}
int zka_init(int, zka_ctx** out) {
  if (!out) return ZKA_E_ARG;
  zka_ctx* c = new zka_ctx();
  if (const char* e = getenv("ZKA_CPU_THREADS")) c->threads = std::max(1, atoi(e));
  *out = c;
  return 0;
}
void zka_shutdown(zka_ctx* ctx) { delete ctx; }

int zka_params_generate(zka_ctx* ctx, const uint8_t rnd[64], uint8_t h_nist[65], uint8_t h_proof[67]) {   // zkpAttestList.ts:88-92
  if (!ctx || !rnd || !h_nist || !h_proof) return ZKA_E_ARG;
  Fe a = fe_from_be(rnd, 32), b = fe_from_be(rnd + 32, 32);
  if (fe_cmp(a, P256N.p) >= 0 || fe_cmp(b, P256P.p) >= 0) { ctx->err = "params draw out of range"; return ZKA_E_ARG; }
  to_bytes(P256, pt_mul(P256, generator(P256), a), h_nist);
  to_bytes(TOM, pt_mul(TOM, generator(TOM), b), h_proof);
  return 0;
}
int zka_params_create(zka_ctx* ctx, const uint8_t h_nist[65], const uint8_t h_proof[67], uint32_t sec_level, zka_params** out) {
  if (!ctx || !h_nist || !h_proof || !out) return ZKA_E_ARG;
  if (sec_level < 1 || sec_level > 80) { ctx->err = "sec_level must be in [1,80]"; return ZKA_E_ARG; }
  zka_params* P = new zka_params();
  P->sec_level = sec_level;
  bool inf = false;
  if (!deserialize(P256, h_nist, P->h_nist, &inf) || inf || !deserialize(TOM, h_proof, P->h_proof)) {
    delete P;
    ctx->err = "params: h point not on its group";
    return ZKA_E_ARG;
  }
  *out = P;
  return 0;
}
void zka_params_destroy(zka_params* P) { delete P; }

int zka_key_to_int(zka_ctx* ctx, uint32_t count, const uint8_t* pk, uint8_t* x_out, int32_t* status) {   // zkpAttestList.ts:94-102
  if (!ctx || !pk || !x_out) return ZKA_E_ARG;
  for (uint32_t i = 0; i < count; i++) {
    WPt p;
    bool inf = false;
    const bool ok = deserialize(P256, pk + (size_t)i * 65, p, &inf) && !inf;
    Fe x = fe_zero(), y;
    if (ok) to_affine(P256, p, x, y);
    else x = from_mont(P256.gx, P256P);
    fe_to_be(x_out + (size_t)i * 32, x, 32);
    if (status) status[i] = ok ? ZKA_OK : ZKA_ERR_INVALID_PK;
  }
  return 0;
}

size_t zka_proof_max_len(uint32_t ring_size, uint32_t sec_level) { return proof_len((int)sec_level, ceil_log2(ring_size), (int)sec_level); }
size_t zka_prove_tape_len(uint32_t ring_size, uint32_t sec_level) {
  return (size_t)32 * (3 + 4 * sec_level + 40 * sec_level + 5 * ceil_log2(ring_size));
}
size_t zka_verify_tape_len_ex(uint32_t ring_size, uint32_t, uint32_t samples) {
  return (size_t)32 * (2 * ceil_log2(ring_size) + 1) + 96 + (size_t)32 * 25 * samples;
}
size_t zka_verify_tape_len(uint32_t ring_size, uint32_t sec) { return zka_verify_tape_len_ex(ring_size, sec, 20); }

int zka_prove_batch(zka_ctx* ctx, const zka_params* P, uint32_t B, const uint8_t* msg_hash, const uint8_t* sig, const uint8_t* pk,
                    const uint32_t* which, const uint8_t* ring, uint32_t N, const uint8_t* tape, size_t tape_stride,
                    uint8_t* proofs, size_t proof_stride, uint32_t* proof_len_out, int32_t* status) {
  if (!ctx || !P || !msg_hash || !sig || !pk || !which || !ring || !tape || !proofs || !proof_len_out || !status) return ZKA_E_ARG;
  if (B == 0) return 0;
  if (N < 2 || N > (1u << 20)) { ctx->err = "ring size must be in [2, 2^20]"; return ZKA_E_ARG; }
  if (proof_stride < zka_proof_max_len(N, P->sec_level)) { ctx->err = "proof_stride < zka_proof_max_len"; return ZKA_E_ARG; }
  if (tape_stride < (size_t)32 * (3 + 4 * P->sec_level + 5 * ceil_log2(N))) { ctx->err = "tape_stride too small"; return ZKA_E_ARG; }
  const std::vector<Fe> keys = ring_scalars(ring, N);
  parallel_for(ctx->threads, B, [&](uint32_t b) {
    uint8_t* row = proofs + (size_t)b * proof_stride;
    status[b] = prove_one(P, msg_hash + (size_t)b * 32, sig + (size_t)b * 64, pk + (size_t)b * 65, which[b], keys,
                          tape + (size_t)b * tape_stride, tape_stride, row, proof_len_out + b);
    if (status[b] != ZKA_OK) { memset(row, 0, proof_stride); proof_len_out[b] = 0; }
  });
  return 0;
}

int zka_verify_batch_ex(zka_ctx* ctx, const zka_params* P, uint32_t B, const uint8_t* msg_hash, const uint8_t* ring, uint32_t N,
                        const uint8_t* proofs, size_t proof_stride, const uint32_t* proof_len_in, const uint8_t* tape, size_t tape_stride,
                        uint8_t* ok, int32_t* status, uint32_t samples) {
  if (!ctx || !P || !msg_hash || !ring || !proofs || !proof_len_in || !tape || !ok || !status) return ZKA_E_ARG;
  if (B == 0) return 0;
  if (N < 2 || N > (1u << 20)) { ctx->err = "ring size must be in [2, 2^20]"; return ZKA_E_ARG; }
  if (samples < 1) { ctx->err = "samples must be >= 1"; return ZKA_E_ARG; }
  if (P->sec_level < samples) { ctx->err = "security level not achieved"; return ZKA_E_ARG; }   // exp.ts:243-245
  if (tape_stride < zka_verify_tape_len_ex(N, P->sec_level, samples)) { ctx->err = "tape_stride < zka_verify_tape_len"; return ZKA_E_ARG; }
  const std::vector<Fe> keys = ring_scalars(ring, N);
  parallel_for(ctx->threads, B, [&](uint32_t b) {
    const size_t len = std::min<size_t>(proof_len_in[b], proof_stride);
    status[b] = proof_len_in[b] > proof_stride ? ZKA_ERR_MALFORMED
                                               : verify_one(P, msg_hash + (size_t)b * 32, keys, proofs + (size_t)b * proof_stride, len,
                                                            tape + (size_t)b * tape_stride, tape_stride, ok + b, (int)samples);
    if (status[b] != ZKA_OK) ok[b] = 0;
  });
  return 0;
}

int zka_verify_batch(zka_ctx* ctx, const zka_params* P, uint32_t B, const uint8_t* msg_hash, const uint8_t* ring, uint32_t N,
                     const uint8_t* proofs, size_t proof_stride, const uint32_t* proof_len_in, const uint8_t* tape, size_t tape_stride,
                     uint8_t* ok, int32_t* status) {
  return zka_verify_batch_ex(ctx, P, B, msg_hash, ring, N, proofs, proof_stride, proof_len_in, tape, tape_stride, ok, status, 20);
}

// ---- layer-wise entry points
int zka_tom_commit_batch(zka_ctx* ctx, const zka_params* P, uint32_t count, const uint8_t* v, const uint8_t* r, uint8_t* out) {
  if (!ctx || !P || !v || !r || !out) return ZKA_E_ARG;
  parallel_for(ctx->threads, count, [&](uint32_t i) {
    Fe vv = s_reduce(fe_from_be(v + (size_t)i * 32, 32), P256P), rr = s_reduce(fe_from_be(r + (size_t)i * 32, 32), P256P);
    to_bytes(TOM, pt_dblmul(TOM, P->h_proof, rr, generator(TOM), vv), out + (size_t)i * 67);   // pedersen.ts:56
  });
  return 0;
}
int zka_p256_mul_batch(zka_ctx* ctx, uint32_t count, const uint8_t* base, const uint8_t* k, uint8_t* out) {
  if (!ctx || !k || !out) return ZKA_E_ARG;
  parallel_for(ctx->threads, count, [&](uint32_t i) {
    WPt b = generator(P256);
    if (base && !deserialize(P256, base + (size_t)i * 65, b)) b = generator(P256);
    to_bytes(P256, pt_mul(P256, b, s_reduce(fe_from_be(k + (size_t)i * 32, 32), P256N)), out + (size_t)i * 65);
  });
  return 0;
}
int zka_field_op_batch(zka_ctx* ctx, int field, int op, uint32_t count, const uint8_t* a, const uint8_t* b, uint8_t* out) {
  if (!ctx || !a || !out || field < 0 || field > 2 || op < 0 || op > 4 || (op < 3 && !b)) return ZKA_E_ARG;
  const Mod& m = field == 0 ? P256P : field == 1 ? P256N : TOMP;
  const int nb = field == 2 ? 33 : 32;
  for (uint32_t i = 0; i < count; i++) {
    Fe x = s_reduce(fe_from_be(a + (size_t)i * nb, nb), m), y = b ? s_reduce(fe_from_be(b + (size_t)i * nb, nb), m) : fe_zero();
    Fe r = op == 0 ? s_mul(x, y, m) : op == 1 ? s_add(x, y, m) : op == 2 ? s_sub(x, y, m) : s_inv(x, m);
    fe_to_be(out + (size_t)i * nb, r, nb);
  }
  return 0;
}
int zka_hash80_batch(zka_ctx* ctx, uint32_t count, const uint8_t* msgs, size_t msg_stride, const uint32_t* len, uint8_t* out) {
  if (!ctx || !msgs || !len || !out) return ZKA_E_ARG;
  for (uint32_t i = 0; i < count; i++) {
    Sha256 s;
    s.init();
    s.update(msgs + (size_t)i * msg_stride, len[i]);
    uint8_t d[32];
    s.final(d);
    memcpy(out + (size_t)i * 10, d, 10);
  }
  return 0;
}

}  // extern "C"
