// imad_peak.cu — integer-pipe peak microbenchmark (roofline denominator for the modmul kernels).
// Measures sustained per-SM throughput of (a) IMAD.WIDE.U32 (32x32+64->64), (b) the carry-chained
// IMAD.WIDE.U32.X form produced by mad.lo.cc/madc.hi.cc pairs, (c) 32-bit IMAD, (d) IADD3.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/imad_peak tools/imad_peak.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int ITERS = 4096;
constexpr int CH = 8;   // independent chains per thread

__global__ void k_wide(uint64_t* out, uint32_t a, uint32_t b) {
  uint64_t acc[CH];
  for (int i = 0; i < CH; i++) acc[i] = threadIdx.x + i;
  uint32_t x = a + threadIdx.x, y = b;
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int i = 0; i < CH; i++) acc[i] = (uint64_t)x * y + acc[i];   // IMAD.WIDE.U32
    x ^= (uint32_t)acc[0];
  }
  uint64_t s = 0;
  for (int i = 0; i < CH; i++) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_chain(uint64_t* out, uint32_t a, uint32_t b) {
  uint32_t lo[CH], hi[CH];
  for (int i = 0; i < CH; i++) { lo[i] = threadIdx.x + i; hi[i] = i; }
  uint32_t x = a + threadIdx.x, y = b, top = 0;
  for (int it = 0; it < ITERS; it++) {
    asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(lo[0]), "+r"(hi[0]) : "r"(x), "r"(y));
#pragma unroll
    for (int i = 1; i < CH; i++)
      asm volatile("madc.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(lo[i]), "+r"(hi[i]) : "r"(x), "r"(y));
    asm volatile("addc.u32 %0, %0, 0;" : "+r"(top));
    x ^= lo[0];
  }
  uint64_t s = top;
  for (int i = 0; i < CH; i++) s += ((uint64_t)hi[i] << 32) | lo[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_imad32(uint64_t* out, uint32_t a, uint32_t b) {
  uint32_t acc[CH];
  for (int i = 0; i < CH; i++) acc[i] = threadIdx.x + i;
  uint32_t x = a + threadIdx.x, y = b;
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int i = 0; i < CH; i++) acc[i] = x * y + acc[i];
    x ^= acc[0];
  }
  uint64_t s = 0;
  for (int i = 0; i < CH; i++) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_iadd3(uint64_t* out, uint32_t a, uint32_t b) {
  uint32_t acc[CH];
  for (int i = 0; i < CH; i++) acc[i] = threadIdx.x + i;
  uint32_t x = a + threadIdx.x, y = b;
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int i = 0; i < CH; i++) asm volatile("add.u32 %0, %0, %1; add.u32 %0, %0, %2;" : "+r"(acc[i]) : "r"(x), "r"(y));
    x ^= acc[0];
  }
  uint64_t s = 0;
  for (int i = 0; i < CH; i++) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// mixed: one IMAD.WIDE + one IADD3 per slot (do the two pipes dual-issue?)
__global__ void k_mixed(uint64_t* out, uint32_t a, uint32_t b) {
  uint64_t acc[CH];
  uint32_t ad[CH];
  for (int i = 0; i < CH; i++) { acc[i] = threadIdx.x + i; ad[i] = i; }
  uint32_t x = a + threadIdx.x, y = b;
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int i = 0; i < CH; i++) {
      acc[i] = (uint64_t)x * y + acc[i];
      asm volatile("add.u32 %0, %0, %1;" : "+r"(ad[i]) : "r"(x));
    }
    x ^= (uint32_t)acc[0];
  }
  uint64_t s = 0;
  for (int i = 0; i < CH; i++) s += acc[i] + ad[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class K>
double run(K kern, const char* name, double ops_per_thread, int blocks, int threads, uint64_t* out, int sm, double mhz) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  kern<<<blocks, threads>>>(out, 3, 5);
  cudaDeviceSynchronize();
  float best = 1e30f;
  for (int r = 0; r < 5; r++) {
    cudaEventRecord(e0);
    kern<<<blocks, threads>>>(out, 3 + r, 5);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  double total = ops_per_thread * (double)blocks * threads;
  double gops = total / (best * 1e-3) / 1e9;
  printf("{\"kernel\": \"%s\", \"ms\": %.4f, \"gops\": %.1f, \"ops_per_clk_per_sm_at_%.0fMHz\": %.2f}\n", name, best, gops, mhz,
         gops * 1e9 / (mhz * 1e6) / sm);
  return gops;
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  int sm = p.multiProcessorCount;
  int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  double mhz = clk / 1000.0;
  printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_mhz_attr\": %.0f}\n", p.name, sm, mhz);
  int threads = 256, blocks = sm * 8;
  uint64_t* out; cudaMalloc(&out, (size_t)blocks * threads * 8);
  double per = (double)ITERS * CH;
  run(k_wide, "imad_wide_u32", per, blocks, threads, out, sm, mhz);
  run(k_chain, "imad_wide_u32_carry_chain", per, blocks, threads, out, sm, mhz);
  run(k_imad32, "imad_u32", per, blocks, threads, out, sm, mhz);
  run(k_iadd3, "iadd_x2", per * 2, blocks, threads, out, sm, mhz);
  run(k_mixed, "imad_wide_plus_iadd(count imad only)", per, blocks, threads, out, sm, mhz);
  return 0;
}
