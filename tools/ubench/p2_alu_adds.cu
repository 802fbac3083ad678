// Poseidon2 pipe-assignment experiment (sm_100a): ptxas turns ~13 % of the permutation's additions into IMAD.IADD, which
// issue on the multiplier pipe the s-boxes already saturate.  These variants keep every addition on the ALU pipe by
// making it a genuine 3-input IADD3: either `hi - u + P` (canonical product = reduce_2p of the lazy form) or `a + b + z`
// with z an opaque zero held in a REGISTER (kernel argument), not a constant-bank operand.
//   bit0: canonical Montgomery product as reduce_2p(lazy)         bit1: a + b + z in add_mod around the s-boxes
//   bit2: a + b + z inside the external linear layer              bit3: a + b + z in the internal rounds
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../../zktls_b200/csrc/poseidon2.cuh"
using namespace zkb;
#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

template <int Z> __device__ __forceinline__ uint32_t addz(uint32_t a, uint32_t b, uint32_t z) { return Z ? a + b + z : a + b; }
template <int Z> __device__ __forceinline__ uint32_t addm(uint32_t a, uint32_t b, uint32_t z) { return reduce_2p(addz<Z>(a, b, z)); }
template <int C> __device__ __forceinline__ uint32_t mmc(uint32_t a, uint32_t b) { return C ? reduce_2p(mont_mul_lazy(a, b)) : mont_mul(a, b); }
// bit4..: Montgomery factor m = lo * P^-1 with P^-1 = 2^31 + 2^27 + 1 as two shift-adds (LEA, ALU pipe) instead of one IMAD
// (fmaheavy pipe, measured 90 % busy in k_hash_rows).  NM = how many of the four products of an s-box use it.
__device__ __forceinline__ uint32_t m_alu(uint32_t lo) { uint32_t t = (lo << 27) + lo; return (lo << 31) + t; }
template <int ALUM> __device__ __forceinline__ uint32_t mm_lazy(uint32_t a, uint32_t b) {
  uint64_t t = (uint64_t)a * b;
  uint32_t m = ALUM ? m_alu((uint32_t)t) : (uint32_t)t * P_INV;
  return (uint32_t)(t >> 32) - mul_hi32(m, P) + P;
}
template <int ALUM> __device__ __forceinline__ uint32_t mm_canon(uint32_t a, uint32_t b) {
  uint64_t t = (uint64_t)a * b;
  uint32_t m = ALUM ? m_alu((uint32_t)t) : (uint32_t)t * P_INV;
  uint32_t r = (uint32_t)(t >> 32) - mul_hi32(m, P);
  uint32_t y = r + P;
  return y < r ? y : r;
}
template <int NM> __device__ __forceinline__ uint32_t sbox_m(uint32_t x) {
  uint32_t x2 = mm_canon<(NM >= 1)>(x, x), x4 = mm_lazy<(NM >= 2)>(x2, x2), x6 = mm_lazy<(NM >= 3)>(x4, x2);
  return mm_canon<(NM >= 4)>(x6, x);
}
template <int C> __device__ __forceinline__ uint32_t sbox(uint32_t x) {
  if (C >= 16) return sbox_m<C / 16>(x);
  uint32_t x2 = mmc<C>(x, x), x4 = mont_mul_lazy(x2, x2), x6 = mont_mul_lazy(x4, x2);
  return mmc<C>(x6, x);
}
template <int Z> __device__ __forceinline__ void m4z(uint32_t& x0, uint32_t& x1, uint32_t& x2, uint32_t& x3, uint32_t z) {
  uint32_t t0 = addm<Z>(x0, x1, z), t1 = addm<Z>(x2, x3, z);
  uint32_t t2 = addm<Z>(addm<Z>(x1, x1, z), t1, z), t3 = addm<Z>(addm<Z>(x3, x3, z), t0, z);
  uint32_t t1_2 = addm<Z>(t1, t1, z), t0_2 = addm<Z>(t0, t0, z);
  uint32_t t4 = addm<Z>(addm<Z>(t1_2, t1_2, z), t3, z), t5 = addm<Z>(addm<Z>(t0_2, t0_2, z), t2, z);
  x0 = addm<Z>(t3, t5, z); x1 = t5; x2 = addm<Z>(t2, t4, z); x3 = t4;
}
template <int Z> __device__ __forceinline__ void mext(uint32_t* s, uint32_t z) {
#pragma unroll
  for (int c = 0; c < 6; ++c) m4z<Z>(s[4 * c], s[4 * c + 1], s[4 * c + 2], s[4 * c + 3], z);
  uint32_t sums[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    uint32_t a = addm<Z>(s[k], s[4 + k], z), b = addm<Z>(s[8 + k], s[12 + k], z), c = addm<Z>(s[16 + k], s[20 + k], z);
    sums[k] = addm<Z>(addm<Z>(a, b, z), c, z);
  }
#pragma unroll
  for (int i = 0; i < 24; ++i) s[i] = addm<Z>(s[i], sums[i & 3], z);
}
template <int V> __device__ __forceinline__ void permute_v(uint32_t* s, uint32_t z) {
  constexpr int C = (V >= 16) ? (V & ~15) : (V & 1), Z1 = (V >> 1) & 1, Z2 = (V >> 2) & 1, Z3 = (V >> 3) & 1;      // V >= 16: V / 16 = products per s-box with the ALU m
  const auto& T = ZKB_P2_TABLES;
  mext<Z2>(s, z);
#pragma unroll 1
  for (int r = 0; r < 4; ++r) {
#pragma unroll
    for (int i = 0; i < 24; ++i) s[i] = sbox<C>(addm<Z1>(s[i], T.ext[r * 24 + i], z));
    mext<Z2>(s, z);
  }
#pragma unroll 1
  for (int r = 0; r < 21; ++r) {
    s[0] = sbox<C>(addm<Z1>(reduce_2p(s[0]), T.in[r], z));
    uint32_t tot = addm<Z3>(p2::sum12(s), p2::sum12(s + 12), z);
#pragma unroll
    for (int i = 0; i < 24; ++i) s[i] = addz<Z3>(tot, reduce_2p(p2::shoup_mul_lazy(s[i], T.diag[i], T.diag_q[i])), z);
  }
#pragma unroll
  for (int i = 0; i < 24; ++i) s[i] = reduce_2p(s[i]);
#pragma unroll 1
  for (int r = 4; r < 8; ++r) {
#pragma unroll
    for (int i = 0; i < 24; ++i) s[i] = sbox<C>(addm<Z1>(s[i], T.ext[r * 24 + i], z));
    mext<Z2>(s, z);
  }
}

// ---- second family: every chosen 2-input addition / subtraction written as min(a + b, ones) with `ones` = 0xffffffff held
// in a uniform register (kernel argument).  ptxas fuses that into ONE VIADDMNMX.U32 Rd, Ra, +-Rb, URones -- an ALU-pipe
// instruction -- and cannot turn it into IMAD.IADD.  W bits: 1 = the canonical products of the s-box (hi - u), 2 = round-constant
// additions, 4 = external linear layer, 8 = partial rounds.
template <int ON> __device__ __forceinline__ uint32_t addw(uint32_t a, uint32_t b, uint32_t ones) { return ON ? min(a + b, ones) : a + b; }
template <int ON> __device__ __forceinline__ uint32_t subw(uint32_t a, uint32_t b, uint32_t ones) { return ON ? min(a - b, ones) : a - b; }
template <int ON> __device__ __forceinline__ uint32_t addmw(uint32_t a, uint32_t b, uint32_t ones) { return reduce_2p(addw<ON>(a, b, ones)); }
template <int ON> __device__ __forceinline__ uint32_t mmw(uint32_t a, uint32_t b, uint32_t ones) {      // canonical Montgomery product
  uint64_t t = (uint64_t)a * b;
  uint32_t m = (uint32_t)t * P_INV;
  uint32_t r = subw<ON>((uint32_t)(t >> 32), mul_hi32(m, P), ones);
  uint32_t y = r + P;
  return y < r ? y : r;
}
template <int ON> __device__ __forceinline__ uint32_t sboxw(uint32_t x, uint32_t ones) {
  uint32_t x2 = mmw<ON>(x, x, ones), x4 = mont_mul_lazy(x2, x2), x6 = mont_mul_lazy(x4, x2);
  return mmw<ON>(x6, x, ones);
}
template <int ON> __device__ __forceinline__ void m4w(uint32_t& x0, uint32_t& x1, uint32_t& x2, uint32_t& x3, uint32_t o) {
  uint32_t t0 = addmw<ON>(x0, x1, o), t1 = addmw<ON>(x2, x3, o);
  uint32_t t2 = addmw<ON>(addmw<ON>(x1, x1, o), t1, o), t3 = addmw<ON>(addmw<ON>(x3, x3, o), t0, o);
  uint32_t t1_2 = addmw<ON>(t1, t1, o), t0_2 = addmw<ON>(t0, t0, o);
  uint32_t t4 = addmw<ON>(addmw<ON>(t1_2, t1_2, o), t3, o), t5 = addmw<ON>(addmw<ON>(t0_2, t0_2, o), t2, o);
  x0 = addmw<ON>(t3, t5, o); x1 = t5; x2 = addmw<ON>(t2, t4, o); x3 = t4;
}
// ON bits: 1 = the six 4x4 blocks, 2 = the column sums, 4 = the final 24 additions; 8 = only the first half of every 4x4 block
template <int ON> __device__ __forceinline__ void m4h(uint32_t& x0, uint32_t& x1, uint32_t& x2, uint32_t& x3, uint32_t o) {
  uint32_t t0 = addmw<1>(x0, x1, o), t1 = addmw<1>(x2, x3, o);
  uint32_t t2 = addmw<1>(addmw<1>(x1, x1, o), t1, o), t3 = addmw<1>(addmw<1>(x3, x3, o), t0, o);
  uint32_t t1_2 = addmw<0>(t1, t1, o), t0_2 = addmw<0>(t0, t0, o);
  uint32_t t4 = addmw<0>(addmw<0>(t1_2, t1_2, o), t3, o), t5 = addmw<0>(addmw<0>(t0_2, t0_2, o), t2, o);
  x0 = addmw<0>(t3, t5, o); x1 = t5; x2 = addmw<0>(t2, t4, o); x3 = t4;
}
template <int ON> __device__ __forceinline__ void mextw(uint32_t* s, uint32_t o) {
#pragma unroll
  for (int c = 0; c < 6; ++c) { if (ON & 8) m4h<1>(s[4 * c], s[4 * c + 1], s[4 * c + 2], s[4 * c + 3], o); else m4w<(ON & 1)>(s[4 * c], s[4 * c + 1], s[4 * c + 2], s[4 * c + 3], o); }
  uint32_t sums[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    uint32_t a = addmw<((ON >> 1) & 1)>(s[k], s[4 + k], o), b = addmw<((ON >> 1) & 1)>(s[8 + k], s[12 + k], o), c = addmw<((ON >> 1) & 1)>(s[16 + k], s[20 + k], o);
    sums[k] = addmw<((ON >> 1) & 1)>(addmw<((ON >> 1) & 1)>(a, b, o), c, o);
  }
#pragma unroll
  for (int i = 0; i < 24; ++i) s[i] = addmw<((ON >> 2) & 1)>(s[i], sums[i & 3], o);
}
template <int W> __device__ __forceinline__ void permute_w(uint32_t* s, uint32_t o) {
  constexpr int W1 = W & 1, W2 = (W >> 1) & 1, W4 = ((W >> 2) & 1) ? 7 : (W >> 4), W8 = (W >> 3) & 1;      // W >> 4: linear-layer sub-selection (mextw ON bits)
  const auto& T = ZKB_P2_TABLES;
  mextw<W4>(s, o);
#pragma unroll 1
  for (int r = 0; r < 4; ++r) {
#pragma unroll
    for (int i = 0; i < 24; ++i) s[i] = sboxw<W1>(addmw<W2>(s[i], T.ext[r * 24 + i], o), o);
    mextw<W4>(s, o);
  }
#pragma unroll 1
  for (int r = 0; r < 21; ++r) {
    s[0] = sboxw<W1>(addmw<W2>(reduce_2p(s[0]), T.in[r], o), o);
    uint32_t tot = addmw<W8>(p2::sum12(s), p2::sum12(s + 12), o);
#pragma unroll
    for (int i = 0; i < 24; ++i) s[i] = addw<W8>(tot, reduce_2p(p2::shoup_mul_lazy(s[i], T.diag[i], T.diag_q[i])), o);
  }
#pragma unroll
  for (int i = 0; i < 24; ++i) s[i] = reduce_2p(s[i]);
#pragma unroll 1
  for (int r = 4; r < 8; ++r) {
#pragma unroll
    for (int i = 0; i < 24; ++i) s[i] = sboxw<W1>(addmw<W2>(s[i], T.ext[r * 24 + i], o), o);
    mextw<W4>(s, o);
  }
}

// ---- third family: w11 plus the Montgomery factor m = lo * P^-1 = lo + (lo << 27) + (lo << 31) built from two shifts and two
// VIADDMNMX additions (4 ALU instructions instead of 1 IMAD) in the first NM products of every s-box.
template <int ON> __device__ __forceinline__ uint32_t m_sel(uint32_t lo, uint32_t ones) {
  if (!ON) return lo * P_INV;
  if (ON == 2) {      // (lo << 27) + (lo << 31) == ((lo ^ (lo << 4)) << 27) mod 2^32: an XOR that ptxas cannot re-fuse into an IMAD
    uint32_t k = (lo ^ (lo << 4)) << 27;
    return min(lo + k, ones);
  }
  uint32_t s1 = min(lo + (lo << 27), ones);
  return min(s1 + (lo << 31), ones);
}
template <int ON> __device__ __forceinline__ uint32_t mmx_canon(uint32_t a, uint32_t b, uint32_t ones) {
  uint64_t t = (uint64_t)a * b;
  uint32_t m = m_sel<ON>((uint32_t)t, ones);
  uint32_t r = min((uint32_t)(t >> 32) - mul_hi32(m, P), ones);
  uint32_t y = r + P;
  return y < r ? y : r;
}
template <int ON> __device__ __forceinline__ uint32_t mmx_lazy(uint32_t a, uint32_t b, uint32_t ones) {
  uint64_t t = (uint64_t)a * b;
  uint32_t m = m_sel<ON>((uint32_t)t, ones);
  return (uint32_t)(t >> 32) - mul_hi32(m, P) + P;
}
template <int NM> __device__ __forceinline__ uint32_t sboxx(uint32_t x, uint32_t ones) {
  constexpr int K = NM >= 10 ? 2 : 1, N = NM % 10;      // NM = 1x: the XOR form of m
  uint32_t x2 = mmx_canon<(N >= 1 ? K : 0)>(x, x, ones), x4 = mmx_lazy<(N >= 2 ? K : 0)>(x2, x2, ones), x6 = mmx_lazy<(N >= 3 ? K : 0)>(x4, x2, ones);
  return mmx_canon<(N >= 4 ? K : 0)>(x6, x, ones);
}
template <int NM> __device__ __forceinline__ void permute_x(uint32_t* s, uint32_t o) {
  const auto& T = ZKB_P2_TABLES;
  p2::m_ext(s);
#pragma unroll 1
  for (int r = 0; r < 4; ++r) {
#pragma unroll
    for (int i = 0; i < 24; ++i) s[i] = sboxx<NM>(addmw<1>(s[i], T.ext[r * 24 + i], o), o);
    p2::m_ext(s);
  }
#pragma unroll 1
  for (int r = 0; r < 21; ++r) {
    s[0] = sboxx<NM>(addmw<1>(reduce_2p(s[0]), T.in[r], o), o);
    uint32_t tot = addmw<1>(p2::sum12(s), p2::sum12(s + 12), o);
#pragma unroll
    for (int i = 0; i < 24; ++i) s[i] = addw<1>(tot, reduce_2p(p2::shoup_mul_lazy(s[i], T.diag[i], T.diag_q[i])), o);
  }
#pragma unroll
  for (int i = 0; i < 24; ++i) s[i] = reduce_2p(s[i]);
#pragma unroll 1
  for (int r = 4; r < 8; ++r) {
#pragma unroll
    for (int i = 0; i < 24; ++i) s[i] = sboxx<NM>(addmw<1>(s[i], T.ext[r * 24 + i], o), o);
    p2::m_ext(s);
  }
}

constexpr int REPS = 14;       // permutations per thread (a 224-column row)
template <int V, int BLOCK> __global__ void __launch_bounds__(BLOCK) kern(uint32_t* out, uint32_t seed, uint32_t z) {
  uint32_t s[24];
  uint32_t gid = blockIdx.x * BLOCK + threadIdx.x;
#pragma unroll
  for (int i = 0; i < 24; ++i) s[i] = 0;
#pragma unroll 1
  for (int rep = 0; rep < REPS; ++rep) {
#pragma unroll
    for (int i = 0; i < 16; ++i) s[i] = (gid * 2654435761u + (uint32_t)(rep * 16 + i) * 40503u + seed) % P;
    if (V < 0) p2::permute(s, ZKB_P2_TABLES); else if (V >= 2000) permute_x<(V >= 2000 ? V - 2000 : 0)>(s, ~z); else if (V >= 1000 && V < 2000) permute_w<(V >= 1000 && V < 2000 ? V - 1000 : 0)>(s, ~z); else permute_v<(V < 0 || V >= 1000 ? 0 : V)>(s, z);
  }
  uint4* o = reinterpret_cast<uint4*>(out + (size_t)gid * 8);
  o[0] = make_uint4(s[0], s[1], s[2], s[3]); o[1] = make_uint4(s[4], s[5], s[6], s[7]);
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
  kern<V, BLOCK><<<threads / BLOCK, BLOCK>>>(out, 7, 0); CHECK(cudaDeviceSynchronize());
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  for (int i = 0; i < 3; ++i) kern<V, BLOCK><<<threads / BLOCK, BLOCK>>>(out, 7, 0);
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
  printf("%-52s block %4d regs %3d occ %2d  %8.3f ms  %7.3f Gperm/s  %6.3f T modmul/s  %s\n", name, BLOCK, fa.numRegs, occ, ms, perms / ms / 1e6, perms * 1356 / ms / 1e9, ok ? "OK" : "MISMATCH");
  cudaFree(out);
}
int main(int argc, char** argv) {
  if (argc > 1) {      // short list: the candidates for the library
    run<-1, 128>("library permute");
    run<1003, 128>("w3  = w1 + w2");
    run<1011, 128>("w11 = w1 + w2 + w8");
    run<1009, 128>("w9  = w1 + w8");
    run<1003, 256>("w3  = w1 + w2");
    run<1011, 256>("w11 = w1 + w2 + w8");
    run<2000, 256>("x0 = w11 (local code)");
    run<2001, 256>("x1 = w11 + shift-add m in 1 of 4 products");
    run<2002, 256>("x2 = w11 + shift-add m in 2 of 4 products");
    run<2003, 256>("x3 = w11 + shift-add m in 3 of 4 products");
    run<2004, 256>("x4 = w11 + shift-add m in 4 of 4 products");
    run<2011, 256>("y1 = w11 + xor-shift m in 1 of 4 products");
    run<2012, 256>("y2 = w11 + xor-shift m in 2 of 4 products");
    run<2013, 256>("y3 = w11 + xor-shift m in 3 of 4 products");
    run<2014, 256>("y4 = w11 + xor-shift m in 4 of 4 products");
    if (argc > 2) return 0;
    run<1011 + 16 * 1, 256>("w11 + the 4x4 blocks");
    run<1011 + 16 * 2, 256>("w11 + column sums");
    run<1011 + 16 * 4, 256>("w11 + final 24 additions");
    run<1011 + 16 * 6, 256>("w11 + column sums + final additions");
    run<1011 + 16 * 8, 256>("w11 + first half of the 4x4 blocks");
    run<1011 + 16 * 12, 256>("w11 + first half of the 4x4 blocks + final additions");
    return 0;
  }
  run<-1, 128>("baseline (library permute)");
  run<0, 128>("v0 same formulation, local code");
  run<1, 128>("v1 canonical product = reduce_2p(lazy)");
  run<2, 128>("v2 a+b+z around s-boxes");
  run<3, 128>("v3 = v1 + v2");
  run<4, 128>("v4 a+b+z in external linear layer");
  run<7, 128>("v7 = v1 + v2 + v4");
  run<8, 128>("v8 a+b+z in internal rounds");
  run<15, 128>("v15 all");
  run<9, 128>("v9 = v1 + v8");
  run<11, 128>("v11 = v1 + v2 + v8");
  run<16 + 10, 128>("v26 = v10 + ALU m in 1 of 4 products");
  run<32 + 10, 128>("v42 = v10 + ALU m in 2 of 4 products");
  run<48 + 10, 128>("v58 = v10 + ALU m in 3 of 4 products");
  run<64 + 10, 128>("v74 = v10 + ALU m in 4 of 4 products");
  run<64 + 0, 128>("v64 = ALU m in 4 of 4 products, plain adds");
  run<32 + 0, 128>("v32 = ALU m in 2 of 4 products, plain adds");
  run<10, 128>("v10 = v2 + v8 (library form)");
  run<1000, 128>("w0  plain");
  run<1001, 128>("w1  s-box canonical subtract via VIADDMNMX");
  run<1002, 128>("w2  round-constant adds via VIADDMNMX");
  run<1004, 128>("w4  external linear layer via VIADDMNMX");
  run<1008, 128>("w8  partial rounds via VIADDMNMX");
  run<1003, 128>("w3  = w1 + w2");
  run<1007, 128>("w7  = w1 + w2 + w4");
  run<1015, 128>("w15 all");
  run<1013, 128>("w13 = w1 + w4 + w8");
  run<1005, 128>("w5  = w1 + w4");
  run<1015, 256>("w15 all");
  run<1007, 256>("w7");
  return 0;
}
