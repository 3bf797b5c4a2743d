// mac_patterns.cu — which way of arranging the 81 products of a 9x9-limb multiplication keeps the
// IMAD.WIDE pipe busiest?  (a) product scanning: every product goes into one 96-bit column accumulator
// (mad.lo.cc / madc.hi.cc / addc), (b) operand scanning with even/odd accumulator arrays: every row is
// two carry chains of mad.lo.cc / madc.hi.cc pairs over distinct accumulators.  No reduction, registers only.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/mac_patterns tools/mac_patterns.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
constexpr int ITERS = 2048;
constexpr int N = 9;

__device__ __forceinline__ void mac3(uint32_t& a0, uint32_t& a1, uint32_t& a2, uint32_t x, uint32_t y) {
  asm("mad.lo.cc.u32 %0, %3, %4, %0;\n\tmadc.hi.cc.u32 %1, %3, %4, %1;\n\taddc.u32 %2, %2, 0;"
      : "+r"(a0), "+r"(a1), "+r"(a2) : "r"(x), "r"(y));
}
struct V { uint32_t v[N]; };
struct W { uint32_t v[2 * N]; };

__device__ __noinline__ W mul_ps(V a, V b) {   // product scanning
  W r;
  uint32_t a0 = 0, a1 = 0, a2 = 0;
#pragma unroll
  for (int k = 0; k < 2 * N - 1; k++) {
#pragma unroll
    for (int i = 0; i < N; i++) {
      const int j = k - i;
      if (j >= 0 && j < N) mac3(a0, a1, a2, a.v[i], b.v[j]);
    }
    r.v[k] = a0; a0 = a1; a1 = a2; a2 = 0;
  }
  r.v[2 * N - 1] = a0;
  return r;
}

// row: lanes (acc[2t], acc[2t+1]) += x * b[j0 + 2t], t = 0..CNT-1, ONE carry chain in ONE asm statement
// (the carry flag must not live across asm statements); the carry-out is added to acc[2*CNT].
__device__ __forceinline__ void row_chain5(uint32_t* acc, uint32_t x, const uint32_t* b) {
  asm("mad.lo.cc.u32 %0, %11, %12, %0;\n\tmadc.hi.cc.u32 %1, %11, %12, %1;\n\t"
      "madc.lo.cc.u32 %2, %11, %13, %2;\n\tmadc.hi.cc.u32 %3, %11, %13, %3;\n\t"
      "madc.lo.cc.u32 %4, %11, %14, %4;\n\tmadc.hi.cc.u32 %5, %11, %14, %5;\n\t"
      "madc.lo.cc.u32 %6, %11, %15, %6;\n\tmadc.hi.cc.u32 %7, %11, %15, %7;\n\t"
      "madc.lo.cc.u32 %8, %11, %16, %8;\n\tmadc.hi.cc.u32 %9, %11, %16, %9;\n\t"
      "addc.u32 %10, %10, 0;"
      : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]), "+r"(acc[6]), "+r"(acc[7]),
        "+r"(acc[8]), "+r"(acc[9]), "+r"(acc[10])
      : "r"(x), "r"(b[0]), "r"(b[2]), "r"(b[4]), "r"(b[6]), "r"(b[8]));
}
__device__ __forceinline__ void row_chain4(uint32_t* acc, uint32_t x, const uint32_t* b) {
  asm("mad.lo.cc.u32 %0, %9, %10, %0;\n\tmadc.hi.cc.u32 %1, %9, %10, %1;\n\t"
      "madc.lo.cc.u32 %2, %9, %11, %2;\n\tmadc.hi.cc.u32 %3, %9, %11, %3;\n\t"
      "madc.lo.cc.u32 %4, %9, %12, %4;\n\tmadc.hi.cc.u32 %5, %9, %12, %5;\n\t"
      "madc.lo.cc.u32 %6, %9, %13, %6;\n\tmadc.hi.cc.u32 %7, %9, %13, %7;\n\t"
      "addc.u32 %8, %8, 0;"
      : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]), "+r"(acc[6]), "+r"(acc[7]),
        "+r"(acc[8])
      : "r"(x), "r"(b[1]), "r"(b[3]), "r"(b[5]), "r"(b[7]));
}
__device__ __noinline__ W mul_os(V a, V b) {   // operand scanning, even/odd arrays: T = E + (O << 32)
  uint32_t E[2 * N + 2], O[2 * N + 2];
#pragma unroll
  for (int i = 0; i < 2 * N + 2; i++) { E[i] = 0; O[i] = 0; }
#pragma unroll
  for (int i = 0; i < N; i++) {
    // a_i * b_j lands at limbs (i+j, i+j+1): even i+j -> E lanes, odd i+j -> O lanes (O offset by one limb)
    if ((i & 1) == 0) {
      row_chain5(E + i, a.v[i], b.v);       // j = 0,2,4,6,8: positions i+j even
      row_chain4(O + i, a.v[i], b.v);       // j = 1,3,5,7: T limb i+j = O index i+j-1
    } else {
      row_chain5(O + i - 1, a.v[i], b.v);   // i+j odd for even j
      row_chain4(E + i + 1, a.v[i], b.v);   // i+j even for odd j
    }
  }
  W r;
  r.v[0] = E[0];
  // T = E + (O << 32): one ripple add (kept out of asm: it is not what is being measured)
  uint64_t carry = 0;
#pragma unroll
  for (int k = 1; k < 2 * N; k++) {
    const uint64_t t = (uint64_t)E[k] + O[k - 1] + carry;
    r.v[k] = (uint32_t)t;
    carry = t >> 32;
  }
  return r;
}

template <int MODE>
__global__ void __launch_bounds__(128) kern(uint32_t* out, uint32_t seed) {
  V a, b;
  for (int i = 0; i < N; i++) { a.v[i] = seed * (i + 1) + threadIdx.x + blockIdx.x; b.v[i] = seed + 31 * i; }
  for (int it = 0; it < ITERS; it++) {
    W r = MODE == 0 ? mul_ps(a, b) : mul_os(a, b);
#pragma unroll
    for (int i = 0; i < N; i++) a.v[i] = r.v[i] ^ r.v[i + N];
  }
  uint32_t s = 0;
  for (int i = 0; i < N; i++) s ^= a.v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void check(uint32_t* out) {   // both patterns must agree
  V a, b;
  for (int i = 0; i < N; i++) { a.v[i] = 0x9e3779b9u * (i + 1) + threadIdx.x * 77u; b.v[i] = 0xffffffffu - 31 * i * threadIdx.x; }
  W x = mul_ps(a, b), y = mul_os(a, b);
  uint32_t bad = 0;
  for (int i = 0; i < 2 * N; i++) bad |= x.v[i] ^ y.v[i];
  out[threadIdx.x] = bad;
}
template <class K>
double run(K k, int blocks, uint32_t* out) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<<<blocks, 128>>>(out, 12345u);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  k<<<blocks, 128>>>(out, 12345u);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  return ms;
}
int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  const int sms = p.multiProcessorCount;
  uint32_t* out;
  cudaMalloc(&out, (size_t)sms * 16 * 128 * 4);
  check<<<1, 128>>>(out);
  uint32_t h[128];
  cudaMemcpy(h, out, sizeof h, cudaMemcpyDeviceToHost);
  uint32_t bad = 0;
  for (int i = 0; i < 128; i++) bad |= h[i];
  printf("{\"patterns_agree\": %s", bad ? "false" : "true");
  for (int cps : {2, 4, 5, 8}) {
    const int blocks = sms * cps;
    const double thr = (double)blocks * 128 * ITERS * 81;
    printf(", \"ps_cta%d_gmacs\": %.1f, \"os_cta%d_gmacs\": %.1f", cps, thr / (run(kern<0>, blocks, out) * 1e-3) / 1e9, cps,
           thr / (run(kern<1>, blocks, out) * 1e-3) / 1e9);
  }
  printf("}\n");
  return 0;
}
