// mul_peak.cu — how fast do the library's own Montgomery multipliers run when nothing else is going on?
// Every thread keeps CH independent product chains in registers (no memory traffic); the result is
// reported as 32x32+64 MACs per second so it can be set against tools/imad_peak (the pipe peak) and
// against the commit kernels' algorithmic MAC rate (bench.py roofline.achieved).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Izkp_ecdsa_b200/csrc -Iinclude -o tools/mul_peak tools/mul_peak.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "zk_field.cuh"

using namespace zk;
constexpr int ITERS = 2048;

template <int CH>
__global__ void __launch_bounds__(128) k_tom(uint32_t* out, uint32_t seed) {
  uint32_t a[CH][9], b[9];
  for (int c = 0; c < CH; c++)
    for (int i = 0; i < 9; i++) a[c][i] = seed * (i + 1) + threadIdx.x + 977 * c + blockIdx.x;
  for (int i = 0; i < 9; i++) b[i] = seed + 31 * i;
  a[0][8] &= 0xff; b[8] &= 3;
  for (int c = 0; c < CH; c++) a[c][8] &= 0xff;
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int c = 0; c < CH; c++) Tomp::mul(a[c], a[c], b);
  }
  uint32_t s = 0;
  for (int c = 0; c < CH; c++)
    for (int i = 0; i < 9; i++) s ^= a[c][i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int CH>
__global__ void __launch_bounds__(128) k_p256(uint32_t* out, uint32_t seed) {
  uint32_t a[CH][8], b[8];
  for (int c = 0; c < CH; c++)
    for (int i = 0; i < 8; i++) a[c][i] = seed * (i + 1) + threadIdx.x + 977 * c + blockIdx.x;
  for (int i = 0; i < 8; i++) b[i] = seed + 31 * i;
  for (int c = 0; c < CH; c++) a[c][7] &= 0x7fffffff;
  b[7] &= 0x7fffffff;
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int c = 0; c < CH; c++) P256p::mul(a[c], a[c], b);
  }
  uint32_t s = 0;
  for (int c = 0; c < CH; c++)
    for (int i = 0; i < 8; i++) s ^= a[c][i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class K>
double run(K kern, int blocks, uint32_t* out) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  kern<<<blocks, 128>>>(out, 12345u);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  kern<<<blocks, 128>>>(out, 12345u);
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
  printf("{\"sms\": %d", sms);
  for (int cps = 2; cps <= 8; cps += (cps < 4 ? 2 : (cps == 4 ? 1 : 3))) {   // resident CTAs per SM: 2, 4, 5, 8
    const int blocks = sms * cps;
    double ms1 = run(k_tom<1>, blocks, out), ms2 = run(k_tom<2>, blocks, out);
    double q1 = run(k_p256<1>, blocks, out), q2 = run(k_p256<2>, blocks, out);
    const double thr = (double)blocks * 128 * ITERS;
    printf(", \"tom_ch1_cta%d_gmacs\": %.1f, \"tom_ch2_cta%d_gmacs\": %.1f", cps, thr * 126 / (ms1 * 1e-3) / 1e9, cps,
           thr * 2 * 126 / (ms2 * 1e-3) / 1e9);
    printf(", \"p256_ch1_cta%d_gmacs\": %.1f, \"p256_ch2_cta%d_gmacs\": %.1f", cps, thr * 64 / (q1 * 1e-3) / 1e9, cps,
           thr * 2 * 64 / (q2 * 1e-3) / 1e9);
  }
  printf("}\n");
  return 0;
}
