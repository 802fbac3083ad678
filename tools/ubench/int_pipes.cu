// Microbenchmark: throughput of the integer instructions a BabyBear modmul is built from, on sm_100a.
// Each kernel runs ILP independent dependency chains per thread; reports lane-ops / clk / SM.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)
constexpr int ITERS = 4096, ILP = 8;

template <int OP> __device__ __forceinline__ void step(uint32_t& a, uint32_t& b, uint32_t c) {
  if (OP == 0) { a = a * c + b; }                                              // IMAD
  else if (OP == 1) { a = __umulhi(a, c) + b; }                                // IMAD.HI
  else if (OP == 2) { uint64_t w = (uint64_t)a * c + (((uint64_t)b << 32) | a); a = (uint32_t)w; b = (uint32_t)(w >> 32); }   // IMAD.WIDE
  else if (OP == 3) { a = a + b + c; }                                          // IADD3
  else if (OP == 4) { a = min(a, a - c); }                                      // IADD + IMNMX.U32
  else if (OP == 5) { a = (a ^ b) & c; }                                        // LOP3
  else if (OP == 6) { a = __funnelshift_l(a, b, 7) ; }                          // SHF
  else if (OP == 7) { asm volatile("{.reg .pred p; setp.lt.u32 p, %0, %1; selp.u32 %0, %0, %2, p;}" : "+r"(a) : "r"(c), "r"(b)); }   // ISETP+SEL
  else if (OP == 8) {   // canonical montgomery mul  (a = a*c mont)
    uint64_t ab = (uint64_t)a * c; uint32_t m = (uint32_t)ab * 0x77ffffffu; uint32_t r = (uint32_t)((ab + (uint64_t)m * 2013265921u) >> 32); uint32_t y = r - 2013265921u; a = min(r, y);
  } else if (OP == 9) {   // lazy montgomery mul (no final reduce)
    uint64_t ab = (uint64_t)a * c; uint32_t m = (uint32_t)ab * 0x77ffffffu; a = (uint32_t)((ab + (uint64_t)m * 2013265921u) >> 32);
  } else if (OP == 10) {  // signed montgomery: hi(ab) - hi(m*P)
    int64_t ab = (int64_t)(int32_t)a * (int32_t)c; int32_t m = (int32_t)((uint32_t)ab * 0x88000001u); int64_t t = ab - (int64_t)m * 2013265921; a = (uint32_t)(t >> 32);
  } else if (OP == 11) {  // mulhi-form montgomery: hi(a*c) - hi(m*P) with separate mul.lo / mul.hi
    uint32_t lo = a * c, hi = __umulhi(a, c); uint32_t m = lo * 0x88000001u; uint32_t t = __umulhi(m, 2013265921u); a = hi - t;
  } else if (OP == 12) {  // modular add canonical
    uint32_t s = a + c; a = min(s, s - 2013265921u);
  } else if (OP == 13) {  // 64-bit add
    uint64_t w = (((uint64_t)b << 32) | a) + (((uint64_t)c << 32) | c); a = (uint32_t)w; b = (uint32_t)(w >> 32);
  } else if (OP == 14) {  // IMAD used as an add: a = a*1 + c via mad (compiler may turn into IADD)
    asm volatile("mad.lo.u32 %0, %0, 1, %1;" : "+r"(a) : "r"(c));
  } else if (OP == 15) {  // double-precision FMA pipe
    double x = __hiloint2double(b, a); x = fma(x, 1.0000001, 3.0); a = __double2loint(x); b = __double2hiint(x);
  }
}
template <int OP> __global__ void kern(uint32_t* out, uint32_t seed, long long* cycles) {
  uint32_t a[ILP], b[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) { a[i] = seed + threadIdx.x * 77 + i; b[i] = seed * 3 + i; }
  uint32_t c = seed | 1;
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) step<OP>(a[i], b[i], c);
  }
  long long t1 = clock64();
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += a[i] ^ b[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}
template <int OP> int run(const char* name, int ops_per_step) {
  uint32_t* out; long long* cyc; long long h;
  int sms = 148, threads = 1024, blocks = sms * 2;     // 2048 threads / SM
  CHECK(cudaMalloc(&out, (size_t)blocks * threads * 4)); CHECK(cudaMalloc(&cyc, 8));
  kern<OP><<<blocks, threads>>>(out, 12345, cyc); CHECK(cudaDeviceSynchronize());
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0); kern<OP><<<blocks, threads>>>(out, 12345, cyc); cudaEventRecord(e1); CHECK(cudaDeviceSynchronize());
  float ms; cudaEventElapsedTime(&ms, e0, e1); CHECK(cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost));
  double steps = (double)blocks * threads * ITERS * ILP;
  // per SM per clk using the in-kernel clock of block 0 (all blocks run concurrently: 2 CTAs/SM resident)
  double per_clk_sm = steps / sms / (double)h;
  printf("%-28s %8.3f ms  %10lld clk  %7.2f steps/clk/SM  (%d instr-class ops per step)  %.2f Gstep/s\n", name, ms, h, per_clk_sm, ops_per_step, steps / ms / 1e6);
  cudaFree(out); cudaFree(cyc); return 0;
}
int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0); printf("%s, %d SMs, clock %d kHz\n", p.name, p.multiProcessorCount, p.clockRate);
  run<0>("IMAD lo", 1); run<1>("IMAD.HI", 1); run<2>("IMAD.WIDE (64b acc)", 1); run<3>("IADD3", 1); run<4>("IADD+IMNMX", 2); run<5>("LOP3", 1); run<6>("SHF", 1);
  run<7>("ISETP+SEL", 2); run<12>("modadd (IADD,IADD,IMNMX)", 3); run<13>("add.u64", 2); run<14>("mad.lo x,1,c", 1); run<15>("DFMA (+cvt)", 1);
  run<8>("mont canonical", 5); run<9>("mont lazy (wide)", 3); run<10>("mont signed (wide)", 3); run<11>("mont mulhi-form", 5);
  return 0;
}
