// Poseidon2 permutation variants for sm_100a: throughput (permutations/s) of thread-per-state kernels, each checked
// against the host permutation.  Used to choose the formulation that goes into zktls_b200/csrc/poseidon2.cuh.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../../zktls_b200/csrc/poseidon2.cuh"
using namespace zkb;
#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__constant__ uint32_t c_zero = 0;   // opaque zero: a third add operand keeps an add on the ALU pipe (IADD3) instead of IMAD.IADD

struct ShoupTables { uint32_t d[24], dq[24]; };
constexpr ShoupTables make_shoup() {
  ShoupTables t{};
  for (int i = 0; i < 24; ++i) { t.d[i] = (uint32_t)(p2::DIAG_CANON[i] % P); t.dq[i] = (uint32_t)((((uint64_t)t.d[i]) << 32) / P); }
  return t;
}
__constant__ ShoupTables c_shoup = make_shoup();

// Montgomery product with the m*P step forced onto IMAD.WIDE (64-bit accumulate) instead of IMAD.HI: the low word of
// t + m*P is zero by construction; adding it to the high word keeps it live so ptxas cannot narrow the op to IMAD.HI.
__device__ __forceinline__ uint32_t mont_wide_lazy(uint32_t a, uint32_t b) {
  uint64_t t = (uint64_t)a * b;
  uint32_t m = (uint32_t)t * 0x77ffffffu;
  uint64_t s;
  asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(s) : "r"(m), "r"(P), "l"(t));
  return (uint32_t)(s >> 32) + (uint32_t)s;
}
__device__ __forceinline__ uint32_t sbox7_wide(uint32_t x) {
  uint32_t x2 = reduce_2p(mont_wide_lazy(x, x));
  uint32_t x4 = mont_wide_lazy(x2, x2);
  uint32_t x6 = mont_wide_lazy(x4, x2);
  return reduce_2p(mont_wide_lazy(x6, x));
}
// subtractive form: r = hi(t) - hi(m' P) in (-P, P), one conditional +P; lazy variant keeps the signed value
__device__ __forceinline__ uint32_t sbox7_sub(uint32_t x) {
  auto mm = [](uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a * b; uint32_t m = (uint32_t)t * 0x88000001u; uint32_t r = (uint32_t)(t >> 32) - __umulhi(m, P); return min(r, r + P); };
  uint32_t x2 = mm(x, x), x3 = mm(x2, x), x4 = mm(x2, x2);
  return mm(x3, x4);
}
// subtractive Montgomery with a PLAIN IMAD.HI (no 64-bit addend): canonical = 5 instr, lazy (0, 2P) = 4 instr (IADD3 hi - u + P)
__device__ __forceinline__ uint32_t mm_c(uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a * b; uint32_t m = (uint32_t)t * 0x88000001u; uint32_t r = (uint32_t)(t >> 32) - __umulhi(m, P); return min(r, r + P); }
__device__ __forceinline__ uint32_t mm_l(uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a * b; uint32_t m = (uint32_t)t * 0x88000001u; return (uint32_t)(t >> 32) - __umulhi(m, P) + P; }
__device__ __forceinline__ uint32_t sbox7_sub_lazy(uint32_t x) { uint32_t x2 = mm_c(x, x), x4 = mm_l(x2, x2), x6 = mm_l(x4, x2); return mm_c(x6, x); }
__device__ __forceinline__ uint32_t sbox7_sub_d3(uint32_t x) { uint32_t x2 = mm_c(x, x), x3 = mm_l(x2, x), x4 = mm_c(x2, x2); return mm_c(x3, x4); }
template <int SB> __device__ __forceinline__ uint32_t sbox_sel(uint32_t x) { return SB == 1 ? sbox7_wide(x) : SB == 2 ? sbox7_sub(x) : SB == 3 ? sbox7_sub_lazy(x) : SB == 4 ? sbox7_sub_d3(x) : p2::sbox7(x); }
template <int ALU> __device__ __forceinline__ uint32_t addm(uint32_t a, uint32_t b) {
  uint32_t s = ALU ? a + b + c_zero : a + b;
  return min(s, s - P);
}
template <int ALU> __device__ __forceinline__ uint32_t add_lazy(uint32_t a, uint32_t b) { return ALU ? a + b + c_zero : a + b; }

template <int ALU> __device__ __forceinline__ void m4v(uint32_t& x0, uint32_t& x1, uint32_t& x2, uint32_t& x3) {
  uint32_t t0 = addm<ALU>(x0, x1), t1 = addm<ALU>(x2, x3);
  uint32_t t2 = addm<ALU>(addm<ALU>(x1, x1), t1), t3 = addm<ALU>(addm<ALU>(x3, x3), t0);
  uint32_t t1_2 = addm<ALU>(t1, t1), t0_2 = addm<ALU>(t0, t0);
  uint32_t t4 = addm<ALU>(addm<ALU>(t1_2, t1_2), t3), t5 = addm<ALU>(addm<ALU>(t0_2, t0_2), t2);
  x0 = addm<ALU>(t3, t5); x1 = t5; x2 = addm<ALU>(t2, t4); x3 = t4;
}
template <int ALU> __device__ __forceinline__ void m_extv(uint32_t* s) {
#pragma unroll
  for (int c = 0; c < 6; ++c) m4v<ALU>(s[4 * c], s[4 * c + 1], s[4 * c + 2], s[4 * c + 3]);
  uint32_t sums[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    uint32_t a = addm<ALU>(s[k], s[4 + k]), b = addm<ALU>(s[8 + k], s[12 + k]), c = addm<ALU>(s[16 + k], s[20 + k]);
    sums[k] = addm<ALU>(addm<ALU>(a, b), c);
  }
#pragma unroll
  for (int i = 0; i < 24; ++i) s[i] = addm<ALU>(s[i], sums[i & 3]);
}

// V: bit0 = Shoup + lazy internal rounds; bit1 = ALU adds around the s-boxes / internal rounds; bit2 = ALU adds inside m_ext too
template <int V> __device__ __forceinline__ void permute_v(uint32_t* s) {
  constexpr int SH = V & 1, A1 = (V >> 1) & 1, A2 = (V >> 2) & 1;
  const auto& T = ZKB_P2_TABLES;
  m_extv<A2>(s);
#pragma unroll 1
  for (int r = 0; r < 4; ++r) {
#pragma unroll
    for (int i = 0; i < 24; ++i) s[i] = p2::sbox7(addm<A1>(s[i], T.ext[r * 24 + i]));
    m_extv<A2>(s);
  }
#pragma unroll 1
  for (int r = 0; r < 21; ++r) {
    if (SH) {
      s[0] = p2::sbox7(addm<A1>(reduce_2p(s[0]), T.in[r]));
      uint64_t a0 = 0, a1 = 0;
#pragma unroll
      for (int i = 0; i < 12; ++i) { a0 += s[i]; a1 += s[12 + i]; }
      uint32_t t0 = addm<0>(reduce_2p(reduce_2p((uint32_t)a0)), reduce_2p((uint32_t)(a0 >> 32) * R_MOD_P));
      uint32_t t1 = addm<0>(reduce_2p(reduce_2p((uint32_t)a1)), reduce_2p((uint32_t)(a1 >> 32) * R_MOD_P));
      uint32_t tot = addm<0>(t0, t1);
#pragma unroll
      for (int i = 0; i < 24; ++i) {
        uint32_t q = __umulhi(s[i], c_shoup.dq[i]);
        uint32_t rr = s[i] * c_shoup.d[i] - q * P;          // [0, 2P)
        s[i] = add_lazy<A1>(tot, reduce_2p(rr));            // [0, 2P)
      }
    } else {
      s[0] = p2::sbox7(addm<A1>(s[0], T.in[r]));
      uint64_t acc = 0;
#pragma unroll
      for (int i = 0; i < 24; ++i) acc += s[i];
      uint32_t lo = (uint32_t)acc, hi = (uint32_t)(acc >> 32);
      uint32_t tot = addm<0>(reduce_2p(reduce_2p(lo)), reduce_2p(hi * R_MOD_P));
#pragma unroll
      for (int i = 0; i < 24; ++i) s[i] = addm<A1>(tot, mont_mul(T.diag[i], s[i]));
    }
  }
  if (SH) {
#pragma unroll
    for (int i = 0; i < 24; ++i) s[i] = reduce_2p(s[i]);
  }
#pragma unroll 1
  for (int r = 4; r < 8; ++r) {
#pragma unroll
    for (int i = 0; i < 24; ++i) s[i] = p2::sbox7(addm<A1>(s[i], T.ext[r * 24 + i]));
    m_extv<A2>(s);
  }
}

template <int SB, int AL = 0> __device__ __forceinline__ void permute_sb(uint32_t* s) {
  const auto& T = ZKB_P2_TABLES;
  m_extv<AL>(s);
#pragma unroll 1
  for (int r = 0; r < 4; ++r) {
#pragma unroll
    for (int i = 0; i < 24; ++i) s[i] = sbox_sel<SB>(addm<AL>(s[i], T.ext[r * 24 + i]));
    m_extv<AL>(s);
  }
#pragma unroll 1
  for (int r = 0; r < 21; ++r) {
    s[0] = sbox_sel<SB>(addm<AL>(reduce_2p(s[0]), T.in[r]));
    uint64_t a0 = 0, a1 = 0;
#pragma unroll
    for (int i = 0; i < 12; ++i) { a0 += s[i]; a1 += s[12 + i]; }
    uint32_t t0 = addm<0>(reduce_2p(reduce_2p((uint32_t)a0)), reduce_2p((uint32_t)(a0 >> 32) * R_MOD_P));
    uint32_t t1 = addm<0>(reduce_2p(reduce_2p((uint32_t)a1)), reduce_2p((uint32_t)(a1 >> 32) * R_MOD_P));
    uint32_t tot = addm<0>(t0, t1);
#pragma unroll
    for (int i = 0; i < 24; ++i) {
      uint32_t q = __umulhi(s[i], c_shoup.dq[i]);
      uint32_t rr = s[i] * c_shoup.d[i] - q * P;
      s[i] = add_lazy<AL>(tot, reduce_2p(rr));
    }
  }
#pragma unroll
  for (int i = 0; i < 24; ++i) s[i] = reduce_2p(s[i]);
#pragma unroll 1
  for (int r = 4; r < 8; ++r) {
#pragma unroll
    for (int i = 0; i < 24; ++i) s[i] = sbox_sel<SB>(addm<AL>(s[i], T.ext[r * 24 + i]));
    m_extv<AL>(s);
  }
}

constexpr int REPS = 14;
#define DUAL_HERE
   // permutations per thread (a 224-column row)
template <int V, int BLOCK> __global__ void __launch_bounds__(BLOCK) kern(uint32_t* out, uint32_t seed) {
  uint32_t s[24];
  uint32_t gid = blockIdx.x * BLOCK + threadIdx.x;
#pragma unroll
  for (int i = 0; i < 24; ++i) s[i] = 0;
#pragma unroll 1
  for (int rep = 0; rep < REPS; ++rep) {
#pragma unroll
    for (int i = 0; i < 16; ++i) s[i] = (gid * 2654435761u + (uint32_t)(rep * 16 + i) * 40503u + seed) % P;
    if (V < 0) p2::permute(s, ZKB_P2_TABLES); else if (V >= 200) permute_sb<(V >= 200 ? V - 200 : 0), 1>(s); else if (V >= 100) permute_sb<(V >= 100 && V < 200 ? V - 100 : 0)>(s); else permute_v<(V < 0 || V >= 100 ? 0 : V)>(s);
  }
  uint4* o = reinterpret_cast<uint4*>(out + (size_t)gid * 8);
  o[0] = make_uint4(s[0], s[1], s[2], s[3]); o[1] = make_uint4(s[4], s[5], s[6], s[7]);
}

// ---- dual-state variant: two independent sponges per thread, half a round out of phase, so that the FMA-bound s-box
// layer of one state is scheduled against the ALU-bound linear layer of the other ----------------------------------
template <int A1> __device__ __forceinline__ void sb_layer(uint32_t* s, const uint32_t* rc) {
#pragma unroll
  for (int i = 0; i < 24; ++i) s[i] = p2::sbox7(addm<A1>(s[i], rc[i]));
}
template <int A1> __device__ __forceinline__ void int_s0(uint32_t* s, uint32_t rc) { s[0] = p2::sbox7(addm<A1>(reduce_2p(s[0]), rc)); }
template <int A1> __device__ __forceinline__ void int_lin(uint32_t* s) {
  uint64_t a0 = 0, a1 = 0;
#pragma unroll
  for (int i = 0; i < 12; ++i) { a0 += s[i]; a1 += s[12 + i]; }
  uint32_t t0 = addm<0>(reduce_2p(reduce_2p((uint32_t)a0)), reduce_2p((uint32_t)(a0 >> 32) * R_MOD_P));
  uint32_t t1 = addm<0>(reduce_2p(reduce_2p((uint32_t)a1)), reduce_2p((uint32_t)(a1 >> 32) * R_MOD_P));
  uint32_t tot = addm<0>(t0, t1);
#pragma unroll
  for (int i = 0; i < 24; ++i) {
    uint32_t q = __umulhi(s[i], c_shoup.dq[i]);
    uint32_t rr = s[i] * c_shoup.d[i] - q * P;
    s[i] = add_lazy<A1>(tot, reduce_2p(rr));
  }
}
template <int A1, int A2> __device__ __forceinline__ void permute_dual(uint32_t* a, uint32_t* b) {
  const auto& T = ZKB_P2_TABLES;
  m_extv<A2>(a);
#pragma unroll 1
  for (int r = 0; r < 4; ++r) {
    sb_layer<A1>(a, T.ext + r * 24); m_extv<A2>(b);
    m_extv<A2>(a); sb_layer<A1>(b, T.ext + r * 24);
  }
  m_extv<A2>(b);      // b has caught up: both are at the start of the internal rounds
#pragma unroll 1
  for (int r = 0; r < 21; ++r) {
    int_s0<A1>(a, T.in[r]);
    if (r > 0) int_lin<A1>(b);
    int_lin<A1>(a);
    int_s0<A1>(b, T.in[r]);
  }
  int_lin<A1>(b);
#pragma unroll
  for (int i = 0; i < 24; ++i) { a[i] = reduce_2p(a[i]); b[i] = reduce_2p(b[i]); }
  sb_layer<A1>(a, T.ext + 4 * 24);
#pragma unroll 1
  for (int r = 4; r < 8; ++r) {
    m_extv<A2>(a); sb_layer<A1>(b, T.ext + r * 24);
    if (r < 7) sb_layer<A1>(a, T.ext + (r + 1) * 24);
    m_extv<A2>(b);
  }
}
template <int A1, int A2, int BLOCK> __global__ void __launch_bounds__(BLOCK) kern_dual(uint32_t* out, uint32_t seed, uint32_t half) {
  uint32_t a[24], b[24];
  uint32_t gid = blockIdx.x * BLOCK + threadIdx.x, gid2 = gid + half;
#pragma unroll
  for (int i = 0; i < 24; ++i) a[i] = b[i] = 0;
#pragma unroll 1
  for (int rep = 0; rep < REPS; ++rep) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      a[i] = (gid * 2654435761u + (uint32_t)(rep * 16 + i) * 40503u + seed) % P;
      b[i] = (gid2 * 2654435761u + (uint32_t)(rep * 16 + i) * 40503u + seed) % P;
    }
    permute_dual<A1, A2>(a, b);
  }
  uint4* o = reinterpret_cast<uint4*>(out + (size_t)gid * 8);
  o[0] = make_uint4(a[0], a[1], a[2], a[3]); o[1] = make_uint4(a[4], a[5], a[6], a[7]);
  o = reinterpret_cast<uint4*>(out + (size_t)gid2 * 8);
  o[0] = make_uint4(b[0], b[1], b[2], b[3]); o[1] = make_uint4(b[4], b[5], b[6], b[7]);
}

static void host_ref(uint32_t gid, uint32_t seed, uint32_t* out8) {
  uint32_t s[24] = {0};
  for (int rep = 0; rep < REPS; ++rep) {
    for (int i = 0; i < 16; ++i) s[i] = (gid * 2654435761u + (uint32_t)(rep * 16 + i) * 40503u + seed) % P;
    p2::permute_host(s);
  }
  for (int i = 0; i < 8; ++i) out8[i] = s[i];
}
template <int V, int BLOCK> void run(const char* name) {
  const size_t threads = (size_t)1 << 22;
  uint32_t* out; CHECK(cudaMalloc(&out, threads * 32));
  int occ = 0; CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern<V, BLOCK>, BLOCK, 0));
  cudaFuncAttributes fa; CHECK(cudaFuncGetAttributes(&fa, kern<V, BLOCK>));
  kern<V, BLOCK><<<threads / BLOCK, BLOCK>>>(out, 7); CHECK(cudaDeviceSynchronize());
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  for (int i = 0; i < 3; ++i) kern<V, BLOCK><<<threads / BLOCK, BLOCK>>>(out, 7);
  cudaEventRecord(e1); CHECK(cudaDeviceSynchronize());
  float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 3;
  uint32_t h[8 * 4], ref[8]; bool ok = true;
  size_t probes[4] = {0, 12345, threads / 2 + 17, threads - 1};
  for (int p = 0; p < 4; ++p) {
    CHECK(cudaMemcpy(h, out + probes[p] * 8, 32, cudaMemcpyDeviceToHost));
    host_ref((uint32_t)probes[p], 7, ref);
    for (int i = 0; i < 8; ++i) ok = ok && (h[i] == ref[i]);
  }
  double perms = (double)threads * REPS;
  printf("%-44s block %4d regs %3d occ %2d  %8.3f ms  %7.3f Gperm/s  %6.3f T modmul/s  %s\n", name, BLOCK, fa.numRegs, occ, ms, perms / ms / 1e6, perms * 1356 / ms / 1e9, ok ? "OK" : "MISMATCH");
  cudaFree(out);
}

template <int A1, int A2, int BLOCK> void run_dual(const char* name) {
  const size_t threads = (size_t)1 << 22;
  uint32_t* out; CHECK(cudaMalloc(&out, threads * 32));
  int occ = 0; CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern_dual<A1, A2, BLOCK>, BLOCK, 0));
  cudaFuncAttributes fa; CHECK(cudaFuncGetAttributes(&fa, kern_dual<A1, A2, BLOCK>));
  kern_dual<A1, A2, BLOCK><<<threads / 2 / BLOCK, BLOCK>>>(out, 7, (uint32_t)(threads / 2)); CHECK(cudaDeviceSynchronize());
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  for (int i = 0; i < 3; ++i) kern_dual<A1, A2, BLOCK><<<threads / 2 / BLOCK, BLOCK>>>(out, 7, (uint32_t)(threads / 2));
  cudaEventRecord(e1); CHECK(cudaDeviceSynchronize());
  float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 3;
  uint32_t h[8], ref[8]; bool ok = true;
  size_t probes[4] = {0, 12345, threads / 2 + 17, threads - 1};
  for (int p = 0; p < 4; ++p) {
    CHECK(cudaMemcpy(h, out + probes[p] * 8, 32, cudaMemcpyDeviceToHost));
    host_ref((uint32_t)probes[p], 7, ref);
    for (int i = 0; i < 8; ++i) ok = ok && (h[i] == ref[i]);
  }
  double perms = (double)threads * REPS;
  printf("%-44s block %4d regs %3d occ %2d  %8.3f ms  %7.3f Gperm/s  %6.3f T modmul/s  %s\n", name, BLOCK, fa.numRegs, occ, ms, perms / ms / 1e6, perms * 1356 / ms / 1e9, ok ? "OK" : "MISMATCH");
  cudaFree(out);
}
int main() {
  run<-1, 128>("baseline (library permute)");
  run<100, 128>("sb0: library sbox in local permute");
  run<101, 128>("sb1: m*P via IMAD.WIDE (forced)");
  run<102, 128>("sb2: subtractive form, x3*x4 chain");
  run<103, 128>("sb3: subtractive, lazy x4/x6 (18 instr)");
  run<104, 128>("sb4: subtractive, depth-3 chain (19 instr)");
  run<103, 256>("sb3: subtractive, lazy x4/x6 (18 instr)");
  run<203, 128>("sb3 + all adds forced to IADD3 (ALU)");
  run<203, 256>("sb3 + all adds forced to IADD3 (ALU)");
  return 0;
  run_dual<0, 0, 128>("dual, shoup+lazy");
  run_dual<0, 0, 64>("dual, shoup+lazy");
  run_dual<0, 0, 256>("dual, shoup+lazy");
  run_dual<1, 0, 128>("dual, shoup+lazy, ALU adds near s-box");
  run_dual<1, 1, 128>("dual, shoup+lazy, ALU adds everywhere");

  run<-1, 128>("baseline (library permute)");
  run<-1, 256>("baseline (library permute)");
  run<0, 128>("v0 same formulation, local code");
  run<1, 128>("v1 shoup+lazy internal");
  run<2, 128>("v2 ALU adds near s-box");
  run<3, 128>("v3 shoup+lazy, ALU adds near s-box");
  run<6, 128>("v6 ALU adds everywhere");
  run<7, 128>("v7 shoup+lazy, ALU adds everywhere");
  run<3, 256>("v3 shoup+lazy, ALU adds near s-box");
  run<3, 64>("v3 shoup+lazy, ALU adds near s-box");
  return 0;
}
