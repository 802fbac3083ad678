// Microbenchmark: can the FP64 pipe carry BabyBear modmuls CONCURRENTLY with the INT32 (fma/alu pipe) Montgomery path?
// Each CTA has 8 warps; the first `fp_warps` of them run an exact double-precision modmul chain (6 DFMA-pipe ops per
// modmul), the rest the Montgomery "C" form (IMAD.WIDE + IMAD + IMAD.HI + 2 ALU).  Every warp runs for a fixed number of
// SM clocks and counts the modmuls it finished, so the split balances itself.  Prints modmul/clk/SM per class.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)
constexpr int ILP = 6;
constexpr uint32_t P = 2013265921u;

__device__ __forceinline__ uint32_t mont_c(uint32_t a, uint32_t b) {
  uint64_t t = (uint64_t)a * b;
  uint32_t m = (uint32_t)t * 0x88000001u;
  uint32_t r = (uint32_t)(t >> 32) - __umulhi(m, P);
  return min(r, r + P);
}
// exact a*b mod P on doubles holding integers; result in about (-P/2 - 2^9, P/2 + 2^9)
__device__ __forceinline__ double fmodmul(double a, double b) {
  const double INVP = 1.0 / 2013265921.0, MAGIC = 6755399441055744.0, PD = 2013265921.0;
  double h = a * b;
  double l = fma(a, b, -h);
  double q = fma(h, INVP, MAGIC) - MAGIC;
  double r = fma(-q, PD, h);
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
    uint32_t a[ILP], c = seed | 1;
#pragma unroll
    for (int i = 0; i < ILP; ++i) a[i] = seed + threadIdx.x * 77 + i;
    while (clock64() - t0 < duration) {
#pragma unroll 1
      for (int it = 0; it < 64; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) a[i] = mont_c(a[i], c);
      }
      done += 64 * ILP;
    }
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s ^= a[i];
    sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
  }
  if ((threadIdx.x & 31) == 0) atomicAdd(&counts[warp < fp_warps ? 0 : 1], done * 32ull);
}

// check the double chain against integer arithmetic
__global__ void check(uint32_t* bad) {
  uint32_t x = 123456789u + threadIdx.x * 7919u, c = 1999999999u - threadIdx.x;
  double a = (double)x; uint64_t ref = x;
  for (int i = 0; i < 1000; ++i) {
    a = fmodmul(a, (double)c);
    ref = ref * c % P;
    long long v = (long long)a; if (v < 0) v += P;
    if ((uint64_t)v != ref) { atomicAdd(bad, 1u); return; }
  }
}

int main() {
  cudaDeviceProp p; CHECK(cudaGetDeviceProperties(&p, 0));
  int sms = p.multiProcessorCount;
  printf("%s, %d SMs, clock %d kHz\n", p.name, sms, p.clockRate);
  uint32_t* bad; CHECK(cudaMalloc(&bad, 4)); CHECK(cudaMemset(bad, 0, 4));
  check<<<1, 256>>>(bad); uint32_t hb; CHECK(cudaMemcpy(&hb, bad, 4, cudaMemcpyDeviceToHost));
  printf("fp64 modmul exactness check: %s\n", hb ? "FAILED" : "ok");
  unsigned long long* counts; uint32_t* sink;
  CHECK(cudaMalloc(&counts, 16));
  for (int ctas = 2; ctas <= 6; ctas += 2) {     // resident CTAs per SM (8 warps each)
    CHECK(cudaMalloc(&sink, (size_t)sms * ctas * 256 * 4));
    for (int fp = 0; fp <= 8; ++fp) {
      const long long duration = 4000000;
      CHECK(cudaMemset(counts, 0, 16));
      kern<<<sms * ctas, 256>>>(counts, sink, fp, duration / 10, 12345u); CHECK(cudaDeviceSynchronize());
      CHECK(cudaMemset(counts, 0, 16));
      cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
      cudaEventRecord(e0); kern<<<sms * ctas, 256>>>(counts, sink, fp, duration, 12345u); cudaEventRecord(e1); CHECK(cudaDeviceSynchronize());
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      unsigned long long h[2]; CHECK(cudaMemcpy(h, counts, 16, cudaMemcpyDeviceToHost));
      double clk = ms * 1e-3 * p.clockRate * 1e3;
      printf("warps/SM %2d  fp64 warps %d/8: %7.3f ms  fp64 %6.2f  int %6.2f  total %6.2f modmul/clk/SM   %.3f T modmul/s\n", ctas * 8, fp, ms,
             h[0] / clk / sms, h[1] / clk / sms, (h[0] + h[1]) / clk / sms, (h[0] + h[1]) / (ms * 1e-3) / 1e12);
    }
    cudaFree(sink);
  }
  return 0;
}
