// Microbenchmark: code-generation variants of the BabyBear Montgomery multiply on sm_100a, each timed alone and with
// NALU extra modular additions per multiply (the Poseidon2 mix is ~1.1 modadd per modmul).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
constexpr uint32_t P = 2013265921u, PNI = 0x77ffffffu;
constexpr int ITERS = 2048, ILP = 6;

__device__ __forceinline__ uint32_t red(uint32_t x) { uint32_t y = x - P; return min(x, y); }
template <int V> __device__ __forceinline__ uint32_t mm(uint32_t a, uint32_t b) {
  if (V == 0) {        // plain C: ptxas picks WIDE + LO + HI
    uint64_t ab = (uint64_t)a * b; uint32_t m = (uint32_t)ab * PNI; return (uint32_t)((ab + (uint64_t)m * P) >> 32);
  } else if (V == 1) { // force the second multiply to stay IMAD.WIDE by consuming its (zero) low word
    uint32_t lo, hi, m, tl, th;
    asm("mul.wide.u32 {%0,%1}, %2, %3;" : "=r"(lo), "=r"(hi) : "r"(a), "r"(b));   // placeholder, replaced below
    return 0;
  } else if (V == 2) { // m on the ALU pipe: m = -(lo*Pinv), Pinv = 2^31 + 2^27 + 1
    uint64_t ab = (uint64_t)a * b; uint32_t lo = (uint32_t)ab;
    uint32_t m = 0u - (lo + (lo << 27) + (lo << 31));
    return (uint32_t)((ab + (uint64_t)m * P) >> 32);
  }
  return 0;
}
// V1 / V3 written with 64-bit PTX registers
__device__ __forceinline__ uint32_t mm_wide2(uint32_t a, uint32_t b) {
  uint64_t ab, t; uint32_t m, lo, hi;
  asm("mul.wide.u32 %0, %1, %2;" : "=l"(ab) : "r"(a), "r"(b));
  m = (uint32_t)ab * PNI;
  asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(t) : "r"(m), "r"(P), "l"(ab));
  asm("mov.b64 {%0,%1}, %2;" : "=r"(lo), "=r"(hi) : "l"(t));
  return hi + lo;      // lo == 0 mathematically; keeps both halves live so ptxas cannot narrow to IMAD.HI
}
__device__ __forceinline__ uint32_t mm_wide2_alum(uint32_t a, uint32_t b) {
  uint64_t ab, t; uint32_t m, lo, hi;
  asm("mul.wide.u32 %0, %1, %2;" : "=l"(ab) : "r"(a), "r"(b));
  uint32_t l = (uint32_t)ab;
  m = 0u - (l + (l << 27) + (l << 31));
  asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(t) : "r"(m), "r"(P), "l"(ab));
  asm("mov.b64 {%0,%1}, %2;" : "=r"(lo), "=r"(hi) : "l"(t));
  return hi + lo;
}
template <int V> __device__ __forceinline__ uint32_t mulv(uint32_t a, uint32_t b) {
  if (V == 1) return mm_wide2(a, b);
  if (V == 3) return mm_wide2_alum(a, b);
  return mm<V>(a, b);
}
template <int V, int NALU> __global__ void __launch_bounds__(256) kern(uint32_t* out, uint32_t seed) {
  uint32_t a[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) a[i] = (seed + threadIdx.x * 77 + i * 1234567u) % P;
  uint32_t c = (seed * 2654435761u) % P;
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) {
      uint32_t r = red(mulv<V>(a[i], c));
#pragma unroll
      for (int k = 0; k < NALU; ++k) r = red(r + a[(i + k + 1) % ILP]);
      a[i] = r;
    }
  }
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s ^= a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int V, int NALU> void run(const char* name) {
  uint32_t* out; int blocks = 148 * 8, threads = 256;
  cudaMalloc(&out, (size_t)blocks * threads * 4);
  kern<V, NALU><<<blocks, threads>>>(out, 12345); cudaDeviceSynchronize();
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0); kern<V, NALU><<<blocks, threads>>>(out, 12345); cudaEventRecord(e1); cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double muls = (double)blocks * threads * ITERS * ILP;
  uint32_t h; cudaMemcpy(&h, out, 4, cudaMemcpyDeviceToHost);
  printf("%-34s NALU=%d  %7.3f ms  %7.2f Gmodmul/s  %6.2f /clk/SM  (chk %08x)\n", name, NALU, ms, muls / ms / 1e6, muls / ms / 1e6 * 1e9 / 148 / 1.965e9 / 1e0 / 1e0, h);
  cudaFree(out);
}
int main() {
  run<0, 0>("C (WIDE+LO+HI)"); run<1, 0>("WIDE+LO+WIDE"); run<2, 0>("WIDE+alu-m+HI"); run<3, 0>("WIDE+alu-m+WIDE");
  run<0, 1>("C (WIDE+LO+HI)"); run<1, 1>("WIDE+LO+WIDE"); run<2, 1>("WIDE+alu-m+HI"); run<3, 1>("WIDE+alu-m+WIDE");
  run<0, 2>("C (WIDE+LO+HI)"); run<1, 2>("WIDE+LO+WIDE"); run<2, 2>("WIDE+alu-m+HI"); run<3, 2>("WIDE+alu-m+WIDE");
  return 0;
}
