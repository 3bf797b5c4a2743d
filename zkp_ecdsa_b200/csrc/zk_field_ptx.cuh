// zk_field_ptx.cuh — sm_100a multiplier kernels for the two hot fields (device only).
//
// Measured on B200 (tools/imad_peak.cu): IMAD.WIDE.U32 (32x32+64) issues at 31 /clk/SM — half
// the rate of a 32-bit IMAD — while IADD3 runs at >110 /clk/SM on the other pipe and dual-issues
// with it.  So the cost of a modular multiplication is its number of 32x32 products; carries,
// shifts and small-constant multiples are free as long as they stay on the ALU pipe.
//
// Both routines are "finely integrated product scanning" (FIPS) Montgomery multiplications:
// column k accumulates sum a_i b_{k-i} + sum m_i p_{k-i} in a 96-bit register triple.  Each
// product is `mad.lo.cc / madc.hi.cc / addc`, which ptxas fuses into ONE IMAD.WIDE.U32 with
// carry-out plus one IADD3.X (checked with cuobjdump).
//
//  * tom.p  (9 limbs, lazy): p = [p0 p1 p2 p3 | 2 | 0 | 4 | 0xfffffffc | 3]; the quotient
//    digit m_i meets only five generic limbs (p0..p3, p7); 2m, 4m, 3m are shifts/adds.
//    81 + 45 = 126 products instead of 171.
//  * p256.p (8 limbs, strict): p = 2^256 - 2^224 + 2^192 + 2^96 - 1 and -1/p = 1 mod 2^32, so
//    m_i = column word and m_i*p is five signed word additions: column k gets
//    -m_k + m_{k-3} + m_{k-6} - m_{k-7} + m_{k-8}.  64 products instead of 136, no quotient mults.
#pragma once
// included from zk_field.cuh after the field descriptors and multi-word helpers

#if defined(__CUDA_ARCH__)
namespace zk {
namespace ptx {

__device__ __forceinline__ void mac3(uint32_t& a0, uint32_t& a1, uint32_t& a2, uint32_t x, uint32_t y) {
  asm("mad.lo.cc.u32 %0, %3, %4, %0;\n\t"
      "madc.hi.cc.u32 %1, %3, %4, %1;\n\t"
      "addc.u32 %2, %2, 0;"
      : "+r"(a0), "+r"(a1), "+r"(a2)
      : "r"(x), "r"(y));
}
__device__ __forceinline__ void add3(uint32_t& a0, uint32_t& a1, uint32_t& a2, uint32_t lo, uint32_t hi) {
  asm("add.cc.u32 %0, %0, %3;\n\t"
      "addc.cc.u32 %1, %1, %4;\n\t"
      "addc.u32 %2, %2, 0;"
      : "+r"(a0), "+r"(a1), "+r"(a2)
      : "r"(lo), "r"(hi));
}
// signed 96-bit accumulator += / -= an unsigned word
__device__ __forceinline__ void addw(uint32_t& a0, uint32_t& a1, uint32_t& a2, uint32_t w) {
  asm("add.cc.u32 %0, %0, %3;\n\t"
      "addc.cc.u32 %1, %1, 0;\n\t"
      "addc.u32 %2, %2, 0;"
      : "+r"(a0), "+r"(a1), "+r"(a2)
      : "r"(w));
}
__device__ __forceinline__ void subw(uint32_t& a0, uint32_t& a1, uint32_t& a2, uint32_t w) {
  asm("sub.cc.u32 %0, %0, %3;\n\t"
      "subc.cc.u32 %1, %1, 0;\n\t"
      "subc.u32 %2, %2, 0;"
      : "+r"(a0), "+r"(a1), "+r"(a2)
      : "r"(w));
}

// x << n / x >> (32 - n) as funnel shifts: plain C shifts become IMAD.SHL / IMAD.U32 on the multiplier
// pipe (ptxas balances pipes assuming the FMA pipe is idle), which is the one pipe this code saturates.
__device__ __forceinline__ uint32_t shl_alu(uint32_t x, uint32_t n) {
  uint32_t r;
  asm("shf.l.clamp.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(0u), "r"(x), "r"(n));
  return r;
}
__device__ __forceinline__ uint32_t shr_top_alu(uint32_t x, uint32_t n) {   // x >> (32 - n)
  uint32_t r;
  asm("shf.l.clamp.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(x), "r"(0u), "r"(n));
  return r;
}

// r = a*b/2^288 mod tom.p, lazy: inputs < 2^13 p, output < 2p (no final subtraction).
__device__ __forceinline__ void tom_mul_body(uint32_t* r, const uint32_t* a, const uint32_t* b) {
  constexpr int N = 9;
  constexpr uint32_t P0 = FpTom::p(0), P1 = FpTom::p(1), P2 = FpTom::p(2), P3 = FpTom::p(3), P7 = FpTom::p(7);
  static_assert(FpTom::p(4) == 2 && FpTom::p(5) == 0 && FpTom::p(6) == 4 && FpTom::p(8) == 3, "tom.p limb structure");
  uint32_t m[N], t[N];
  uint32_t a0 = 0, a1 = 0, a2 = 0;
  // (two accumulator triples per column for more ILP were measured: commit kernels 6 % slower,
  //  single-thread chains only 9 % faster — kept single)
#pragma unroll
  for (int k = 0; k < 2 * N; k++) {
    // a_i * b_{k-i}
#pragma unroll
    for (int i = 0; i < N; i++) {
      const int j = k - i;
      if (j >= 0 && j < N) mac3(a0, a1, a2, a[i], b[j]);
    }
    // m_i * p_{k-i}, i < min(k, N)   (p_j generic for j in {1,2,3,7}; shifts for j in {4,6,8})
#pragma unroll
    for (int i = 0; i < N; i++) {
      const int j = k - i;
      if (i < k && j >= 1 && j < N) {
        if (j == 1) mac3(a0, a1, a2, m[i], P1);
        else if (j == 2) mac3(a0, a1, a2, m[i], P2);
        else if (j == 3) mac3(a0, a1, a2, m[i], P3);
        else if (j == 7) mac3(a0, a1, a2, m[i], P7);
        else if (j == 4) add3(a0, a1, a2, shl_alu(m[i], 1), shr_top_alu(m[i], 1));
        else if (j == 6) add3(a0, a1, a2, shl_alu(m[i], 2), shr_top_alu(m[i], 2));
        else if (j == 8) { add3(a0, a1, a2, shl_alu(m[i], 1), shr_top_alu(m[i], 1)); add3(a0, a1, a2, m[i], 0u); }
      }
    }
    if (k < N) {
      m[k] = a0 * FpTom::kN0Inv;
      mac3(a0, a1, a2, m[k], P0);   // a0 becomes 0
    } else {
      t[k - N] = a0;
    }
    a0 = a1; a1 = a2; a2 = 0;
  }
#pragma unroll
  for (int i = 0; i < N; i++) r[i] = t[i];
}

// r = a*b/2^256 mod p256.p, strict: inputs < p, output < p.
__device__ __forceinline__ void p256_mul_body(uint32_t* r, const uint32_t* a, const uint32_t* b) {
  constexpr int N = 8;
  uint32_t m[N], t[N + 1];
  uint32_t a0 = 0, a1 = 0, a2 = 0;   // signed 96-bit column accumulator (two's complement)
#pragma unroll
  for (int k = 0; k < 2 * N; k++) {
#pragma unroll
    for (int i = 0; i < N; i++) {
      const int j = k - i;
      if (j >= 0 && j < N) mac3(a0, a1, a2, a[i], b[j]);
    }
    // + m_{k-3} + m_{k-6} - m_{k-7} + m_{k-8}
    if (k - 3 >= 0 && k - 3 < N) addw(a0, a1, a2, m[k - 3 < 0 ? 0 : (k - 3 >= N ? 0 : k - 3)]);
    if (k - 6 >= 0 && k - 6 < N) addw(a0, a1, a2, m[k - 6 < 0 ? 0 : (k - 6 >= N ? 0 : k - 6)]);
    if (k - 7 >= 0 && k - 7 < N) subw(a0, a1, a2, m[k - 7 < 0 ? 0 : (k - 7 >= N ? 0 : k - 7)]);
    if (k - 8 >= 0 && k - 8 < N) addw(a0, a1, a2, m[k - 8 < 0 ? 0 : (k - 8 >= N ? 0 : k - 8)]);
    if (k < N) {
      m[k] = a0;          // -1/p = 1 mod 2^32: the quotient digit is the column word itself
      a0 = 0;             // column - m_k*1 clears the low word exactly (no borrow: a0 - a0)
    } else {
      t[k - N] = a0;
    }
    a0 = a1; a1 = a2; a2 = (uint32_t)((int32_t)a2 >> 31);
  }
  t[N] = a0;   // 0 or 1
  uint32_t u[N];
  uint32_t br = sub_p<FpP256>(u, t);
  const bool ge = (t[N] != 0) || (br == 0);
#pragma unroll
  for (int i = 0; i < N; i++) r[i] = ge ? u[i] : t[i];
}

// The multipliers are real (non-inlined) functions: operands and result travel in registers
// (by-value structs; the device ABI keeps them in R4..), and every warp of the SM executes the
// same ~5 KB body.  Fully inlining them made the commitment kernel ~130 KB of straight-line
// code and the warps stalled on instruction fetch (ncu: stall_no_instruction 2.7 per issue).
struct V9 { uint32_t v[9]; };
struct V8 { uint32_t v[8]; };
static __device__ __noinline__ V9 tom_mul_fn(V9 a, V9 b) {
  V9 r;
  tom_mul_body(r.v, a.v, b.v);
  return r;
}
static __device__ __noinline__ V8 p256_mul_fn(V8 a, V8 b) {
  V8 r;
  p256_mul_body(r.v, a.v, b.v);
  return r;
}
__device__ __forceinline__ void tom_mul(uint32_t* r, const uint32_t* a, const uint32_t* b) {
#if defined(ZKA_INLINE_MUL)
  tom_mul_body(r, a, b);
#else
  V9 x, y;
#pragma unroll
  for (int i = 0; i < 9; i++) { x.v[i] = a[i]; y.v[i] = b[i]; }
  V9 z = tom_mul_fn(x, y);
#pragma unroll
  for (int i = 0; i < 9; i++) r[i] = z.v[i];
#endif
}
__device__ __forceinline__ void p256_mul(uint32_t* r, const uint32_t* a, const uint32_t* b) {
#if defined(ZKA_INLINE_MUL)
  p256_mul_body(r, a, b);
#else
  V8 x, y;
#pragma unroll
  for (int i = 0; i < 8; i++) { x.v[i] = a[i]; y.v[i] = b[i]; }
  V8 z = p256_mul_fn(x, y);
#pragma unroll
  for (int i = 0; i < 8; i++) r[i] = z.v[i];
#endif
}

}  // namespace ptx
}  // namespace zk
#endif
