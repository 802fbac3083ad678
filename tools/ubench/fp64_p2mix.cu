// Would a dual-path Poseidon2 pay?  INT32 warps run the production permutation (poseidon2.cuh) in a loop; the first
// `fp_warps` of every 8 run an exact FP64 modmul chain (6 DFMA-pipe ops each) as a stand-in for an FP64 permutation
// (which would need ~8136 multiply ops + ~2100 adds on that pipe).  Every warp runs for a fixed number of clocks.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../zktls_b200/csrc/poseidon2.cuh"
using namespace zkb;
#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)
constexpr int ILP = 6;
__device__ __forceinline__ double fmodmul(double a, double b) {
  const double INVP = 1.0 / 2013265921.0, MAGIC = 6755399441055744.0, PD = 2013265921.0;
  double h = a * b, l = fma(a, b, -h), q = fma(h, INVP, MAGIC) - MAGIC, r = fma(-q, PD, h);
  return r + l;
}
__global__ void __launch_bounds__(256) kern(unsigned long long* counts, uint32_t* sink, int fp_warps, long long duration, uint32_t seed) {
  const int warp = threadIdx.x >> 5;
  unsigned long long done = 0;
  const long long t0 = clock64();
  if (warp < fp_warps) {
    double a[ILP], c = (double)(seed | 1);
#pragma unroll
    for (int i = 0; i < ILP; ++i) a[i] = (double)(seed % 1000003u + threadIdx.x * 77 + i);
    while (clock64() - t0 < duration) {
#pragma unroll 1
      for (int it = 0; it < 64; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) a[i] = fmodmul(a[i], c);
      }
      done += 64 * ILP;
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += a[i];
    sink[blockIdx.x * blockDim.x + threadIdx.x] = (uint32_t)(long long)s;
  } else {
    uint32_t s[24];
#pragma unroll
    for (int i = 0; i < 24; ++i) s[i] = (seed + threadIdx.x * 7919u + i * 104729u) % P;
    while (clock64() - t0 < duration) {
      p2::permute(s, ZKB_P2_TABLES);
      done += 1;
    }
    uint32_t x = 0;
#pragma unroll
    for (int i = 0; i < 24; ++i) x ^= s[i];
    sink[blockIdx.x * blockDim.x + threadIdx.x] = x;
  }
  if ((threadIdx.x & 31) == 0) atomicAdd(&counts[warp < fp_warps ? 0 : 1], done * 32ull);
}
int main() {
  cudaDeviceProp p; CHECK(cudaGetDeviceProperties(&p, 0));
  int sms = p.multiProcessorCount;
  printf("%s, %d SMs, clock %d kHz\n", p.name, sms, p.clockRate);
  unsigned long long* counts; uint32_t* sink;
  CHECK(cudaMalloc(&counts, 16));
  for (int ctas = 4; ctas <= 6; ctas += 2) {
    CHECK(cudaMalloc(&sink, (size_t)sms * ctas * 256 * 4));
    for (int fp = 0; fp <= 4; ++fp) {
      const long long duration = 20000000;
      CHECK(cudaMemset(counts, 0, 16));
      kern<<<sms * ctas, 256>>>(counts, sink, fp, duration / 10, 12345u); CHECK(cudaDeviceSynchronize());
      CHECK(cudaMemset(counts, 0, 16));
      cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
      cudaEventRecord(e0); kern<<<sms * ctas, 256>>>(counts, sink, fp, duration, 12345u); cudaEventRecord(e1); CHECK(cudaDeviceSynchronize());
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      unsigned long long h[2]; CHECK(cudaMemcpy(h, counts, 16, cudaMemcpyDeviceToHost));
      double int_perm = h[1] / (ms * 1e-3), fp_mm = h[0] / (ms * 1e-3);
      double fp_perm_equiv = fp_mm / 1356.0 * (8136.0 / 10200.0);
      printf("warps/SM %2d fp64 warps %d/8: %7.3f ms  INT %6.3f Gperm/s   FP64 %6.3f T modmul/s (~%5.3f Gperm/s)   total ~%6.3f Gperm/s\n", ctas * 8, fp, ms,
             int_perm / 1e9, fp_mm / 1e12, fp_perm_equiv / 1e9, (int_perm + fp_perm_equiv) / 1e9);
    }
    cudaFree(sink);
  }
  return 0;
}
