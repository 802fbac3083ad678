// Integer-pipe microbenchmark v2 for sm_100a: every kernel is a grid of 148 * occ resident CTAs running a long loop of
// independent dependency chains; the rate is taken from CUDA events and from the SM clock measured in-kernel
// (clock64 / globaltimer), so it does not depend on the boost state.  Reports warp-instructions / clk / SMSP.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)
constexpr int ITERS = 20000, ILP = 8, THREADS = 256;
constexpr uint32_t P = 2013265921u;

template <int OP> __device__ __forceinline__ void step(uint32_t& a, uint32_t& b, uint32_t c, uint32_t d) {
  if (OP == 0) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a) : "r"(c), "r"(d));
  else if (OP == 1) asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(a) : "r"(c), "r"(d));
  else if (OP == 2) asm volatile("{.reg .u64 t; mul.wide.u32 t, %0, %2; mov.b64 {%0, %1}, t;}" : "+r"(a), "+r"(b) : "r"(c));
  else if (OP == 3) asm volatile("{.reg .u64 t, u; mov.b64 u, {%0, %1}; mad.wide.u32 t, %0, %2, u; mov.b64 {%0, %1}, t;}" : "+r"(a), "+r"(b) : "r"(c));
  else if (OP == 4) asm volatile("add.u32 %0, %0, %1;" : "+r"(a) : "r"(c));
  else if (OP == 5) asm volatile("{.reg .u32 t; add.u32 t, %0, %1; min.u32 %0, %0, t;}" : "+r"(a) : "r"(c));       // VIADDMNMX
  else if (OP == 6) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a) : "r"(c), "r"(d));
  else if (OP == 7) asm volatile("min.u32 %0, %0, %1;" : "+r"(a) : "r"(c));
  else if (OP == 8) asm volatile("shf.l.wrap.b32 %0, %0, %1, 7;" : "+r"(a) : "r"(c));
  else if (OP == 9) { asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a) : "r"(c), "r"(d)); asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(b) : "r"(c), "r"(d)); }   // IMAD || LOP3
  else if (OP == 10) { asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a) : "r"(c), "r"(d)); asm volatile("{.reg .u32 t; add.u32 t, %0, %1; min.u32 %0, %0, t;}" : "+r"(b) : "r"(c)); }   // IMAD || VIADDMNMX
  else if (OP == 11) { asm volatile("{.reg .u64 t; mul.wide.u32 t, %0, %2; mov.b64 {%0, %1}, t;}" : "+r"(a), "+r"(b) : "r"(c)); asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(b) : "r"(c), "r"(d)); }  // WIDE || LOP3
  else if (OP == 12) {   // montgomery, HI form, lazy
    uint64_t ab = (uint64_t)a * c; uint32_t m = (uint32_t)ab * 0x77ffffffu; a = (uint32_t)((ab + (uint64_t)m * P) >> 32);
  } else if (OP == 13) {  // montgomery canonical
    uint64_t ab = (uint64_t)a * c; uint32_t m = (uint32_t)ab * 0x77ffffffu; uint32_t r = (uint32_t)((ab + (uint64_t)m * P) >> 32); a = min(r, r - P);
  } else if (OP == 14) {  // shoup constant multiply: c = w, d = floor(w * 2^32 / P); result in [0, 2P)
    uint32_t q = __umulhi(a, d); a = a * c - q * P;
  } else if (OP == 15) {  // shoup with wide multiply for the quotient
    uint32_t q = (uint32_t)(((uint64_t)a * d) >> 32); a = a * c - q * P;
  } else if (OP == 16) {  // montgomery, subtractive form with two mul.hi: hi(ab) - hi(m P)
    uint32_t lo = a * c, hi = __umulhi(a, c); uint32_t m = lo * 0x88000001u; a = hi - __umulhi(m, P);
  } else if (OP == 17) {  // add on fma pipe + add on alu pipe
    asm volatile("mad.lo.u32 %0, %0, 1, %1;" : "+r"(a) : "r"(c)); asm volatile("add.u32 %0, %0, %1;" : "+r"(b) : "r"(c));
  } else if (OP == 18) {  // 64-bit add
    asm volatile("{.reg .u64 t, u; mov.b64 t, {%0, %1}; mov.b64 u, {%2, %3}; add.u64 t, t, u; mov.b64 {%0, %1}, t;}" : "+r"(a), "+r"(b) : "r"(c), "r"(d));
  }
}
template <int OP> __global__ void __launch_bounds__(THREADS) kern(uint32_t* out, uint32_t c, uint32_t d, unsigned long long* clk) {
  uint32_t a[ILP], b[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) { a[i] = c + threadIdx.x * 77 + i; b[i] = c * 3 + i; }
  unsigned long long g0, g1; long long t0 = clock64();
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(g0));
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) step<OP>(a[i], b[i], c, d);
  }
  long long t1 = clock64();
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(g1));
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += a[i] ^ b[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) { clk[0] = t1 - t0; clk[1] = g1 - g0; }
}
template <int OP> void run(const char* name, int instrs) {
  int occ = 0; CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern<OP>, THREADS, 0));
  int blocks = 148 * occ;
  uint32_t* out; unsigned long long* clk; unsigned long long h[2];
  CHECK(cudaMalloc(&out, (size_t)blocks * THREADS * 4)); CHECK(cudaMalloc(&clk, 16));
  kern<OP><<<blocks, THREADS>>>(out, 12345, 0x9e3779b9u, clk); CHECK(cudaDeviceSynchronize());
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0); kern<OP><<<blocks, THREADS>>>(out, 12345, 0x9e3779b9u, clk); cudaEventRecord(e1); CHECK(cudaDeviceSynchronize());
  float ms; cudaEventElapsedTime(&ms, e0, e1); CHECK(cudaMemcpy(h, clk, 16, cudaMemcpyDeviceToHost));
  double mhz = (double)h[0] / (double)h[1] * 1e3;
  double warp_steps_per_smsp = (double)occ * (THREADS / 32) / 4.0 * ITERS * ILP;     // per SMSP over the whole kernel
  double clk_per_step = (double)h[0] / warp_steps_per_smsp;
  printf("%-34s occ %2d  %7.3f ms  sm %6.0f MHz  %6.2f clk/warp-step/SMSP  (%d instr/step -> %5.2f clk/instr, %6.1f lane-steps/clk/SM)\n",
         name, occ, ms, mhz, clk_per_step, instrs, clk_per_step / instrs, 128.0 / clk_per_step);
  cudaFree(out); cudaFree(clk);
}
int main() {
  run<0>("IMAD lo", 1); run<1>("IMAD.HI", 1); run<2>("IMAD.WIDE (mul)", 1); run<3>("IMAD.WIDE (mad 64b acc)", 1); run<4>("add.u32", 1);
  run<5>("add+min (VIADDMNMX)", 1); run<6>("LOP3", 1); run<7>("min.u32", 1); run<8>("SHF", 1); run<18>("add.u64", 2);
  run<9>("IMAD + LOP3 (2 chains)", 2); run<10>("IMAD + VIADDMNMX (2 chains)", 2); run<11>("WIDE + LOP3", 2); run<17>("mad-add + add", 2);
  run<12>("mont lazy", 3); run<13>("mont canonical", 4); run<16>("mont 2x mul.hi", 5); run<14>("shoup (mul.hi)", 3); run<15>("shoup (mul.wide)", 3);
  return 0;
}
