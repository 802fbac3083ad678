// IMMA (mma.sync m16n8k32 u8 x u8 -> s32) on sm_100a: fragment-layout check against the host, issue rate per SM, and whether it overlaps
// with IMAD.WIDE work on the multiplier pipe.  Question behind it: eval_check's random linear combination  sum_k v_k * mix^k  is a
// (points x terms) . (terms x 4) contraction of 31-bit integers; with v as its own 4 bytes (k = 4 term + byte) and the powers
// pre-multiplied by 2^(8 byte) and split into bytes it is an exact u8 GEMM with N = 16.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/imma_dot tools/ubench/imma_dot.cu && /tmp/imma_dot
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)
__device__ __forceinline__ void imma(int (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
// layout under test (PTX ISA, m16n8k32 .u8): lane = 4 g + t.  A regs: a0 = row g, k 4t..4t+3; a1 = row g+8, same k; a2 = row g, k 16+4t..;
// a3 = row g+8, k 16+4t...  B regs: b0 = k 4t..4t+3, col g; b1 = k 16+4t.., col g.  C: c0,c1 = row g, cols 2t, 2t+1; c2,c3 = row g+8.
__global__ void k_layout(const uint8_t* A /*16x32 row-major*/, const uint8_t* B /*32x8, B[k][n]*/, int* C /*16x8*/) {
  const int lane = threadIdx.x, g = lane >> 2, t = lane & 3;
  uint32_t a[4], b[2]; int c[4] = {0, 0, 0, 0};
  auto packA = [&](int row, int k0) { uint32_t v = 0; for (int i = 0; i < 4; ++i) v |= (uint32_t)A[row * 32 + k0 + i] << (8 * i); return v; };
  auto packB = [&](int k0, int col) { uint32_t v = 0; for (int i = 0; i < 4; ++i) v |= (uint32_t)B[(k0 + i) * 8 + col] << (8 * i); return v; };
  a[0] = packA(g, 4 * t); a[1] = packA(g + 8, 4 * t); a[2] = packA(g, 16 + 4 * t); a[3] = packA(g + 8, 16 + 4 * t);
  b[0] = packB(4 * t, g); b[1] = packB(16 + 4 * t, g);
  imma(c, a, b);
  C[g * 8 + 2 * t] = c[0]; C[g * 8 + 2 * t + 1] = c[1]; C[(g + 8) * 8 + 2 * t] = c[2]; C[(g + 8) * 8 + 2 * t + 1] = c[3];
}
template <int NI, int NW>
__global__ void __launch_bounds__(512) k_rate(int reps, int* out, uint32_t seed) {
  uint32_t a[4] = {seed + threadIdx.x, seed * 3 + 1, seed ^ 0x55u, seed + 7}, b[2] = {seed * 5, seed + 11};
  int c[4][4] = {};
  unsigned long long w[4] = {seed, seed + 1, seed + 2, seed + 3};
  uint32_t x = seed | 1u, y = threadIdx.x * 2654435761u + 1u;
  for (int r = 0; r < reps; ++r) {
#pragma unroll
    for (int i = 0; i < NI; ++i) imma(c[i & 3], a, b);
#pragma unroll
    for (int i = 0; i < NW; ++i) w[i & 3] += (unsigned long long)x * (y + i);      // IMAD.WIDE chains
    a[0] += 1; y += x;
  }
  int s = 0;
  for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + (int)(w[0] ^ w[1] ^ w[2] ^ w[3]) + (int)((w[0] ^ w[1] ^ w[2] ^ w[3]) >> 32);
}
template <int NI, int NW>
static void rate(const char* what, int warps, int* d_out, int sms, double mhz) {
  const int reps = 4096, blocks = sms;
  k_rate<NI, NW><<<blocks, warps * 32>>>(16, d_out, 3); CK(cudaDeviceSynchronize());
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0); k_rate<NI, NW><<<blocks, warps * 32>>>(reps, d_out, 3); cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double clk = ms * 1e-3 * mhz * 1e6;
  printf("%-34s warps/SM %2d: %7.3f ms  IMMA/clk/SM %6.3f  (u8 MAC/clk/SM %7.0f)  IMAD.WIDE lanes/clk/SM %6.2f\n", what, warps, ms,
         (double)reps * NI * warps / clk, (double)reps * NI * warps * 4096 / clk, (double)reps * NW * warps * 32 / clk);
}
int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  const double mhz = khz / 1000.0;
  printf("%s, %d SMs, %0.f MHz nominal\n", p.name, p.multiProcessorCount, mhz);
  std::vector<uint8_t> A(16 * 32), B(32 * 8); std::vector<int> C(16 * 8), R(16 * 8, 0);
  srand(7); for (auto& v : A) v = rand() & 255; for (auto& v : B) v = rand() & 255;
  for (int r = 0; r < 16; ++r) for (int n = 0; n < 8; ++n) for (int k = 0; k < 32; ++k) R[r * 8 + n] += (int)A[r * 32 + k] * (int)B[k * 8 + n];
  uint8_t *dA, *dB; int* dC; CK(cudaMalloc(&dA, A.size())); CK(cudaMalloc(&dB, B.size())); CK(cudaMalloc(&dC, C.size() * 4));
  CK(cudaMemcpy(dA, A.data(), A.size(), cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, B.data(), B.size(), cudaMemcpyHostToDevice));
  k_layout<<<1, 32>>>(dA, dB, dC); CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(C.data(), dC, C.size() * 4, cudaMemcpyDeviceToHost));
  int bad = 0; for (size_t i = 0; i < C.size(); ++i) bad += C[i] != R[i];
  printf("fragment layout (m16n8k32 u8.u8): %s (%d of 128 entries differ)\n", bad ? "MISMATCH" : "matches the host product", bad);
  int* d_out; CK(cudaMalloc(&d_out, (size_t)p.multiProcessorCount * 512 * 4));
  for (int warps : {4, 8, 16}) {
    rate<8, 0>("IMMA only (8 per rep)", warps, d_out, p.multiProcessorCount, mhz);
    rate<0, 16>("IMAD.WIDE only (16 per rep)", warps, d_out, p.multiProcessorCount, mhz);
    rate<4, 16>("4 IMMA + 16 IMAD.WIDE per rep", warps, d_out, p.multiProcessorCount, mhz);
    rate<1, 16>("1 IMMA + 16 IMAD.WIDE per rep", warps, d_out, p.multiProcessorCount, mhz);
  }
  return 0;
}
