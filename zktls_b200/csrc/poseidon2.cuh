// Poseidon2 permutation over BabyBear (t = 24) for host and device.
// Replaces risc0-zkp 1.2.5 `core::hash::poseidon2::poseidon2_mix` and sppark's `poseidon2` device code
// (un-vendored).  The state lives in 24 registers per thread.
#pragma once
#include "field.cuh"
#include "poseidon2_consts.h"

namespace zkb { namespace p2 {

// Constant tables, built at compile time from the canonical values.  Round constants are Montgomery words.  The
// internal-layer diagonal is kept as PLAIN integers d together with dq = floor(d * 2^32 / P): multiplying a Montgomery
// word x by the constant d needs no Montgomery reduction -- x * d - mulhi(x, dq) * P lies in [0, 2P) for any 32-bit x
// (Shoup's precomputed-quotient multiplication), one mul.hi + two mul.lo instead of mul.wide + mul.lo + mul.hi.
struct ConstTables {
  uint32_t ext[8 * 24];
  uint32_t in[21];
  uint32_t diag[24];
  uint32_t diag_q[24];
};
constexpr ConstTables make_tables() {
  ConstTables t{};
  for (int i = 0; i < 8 * 24; ++i) t.ext[i] = mont_const(RC_EXT_CANON[i]);
  for (int i = 0; i < 21; ++i) t.in[i] = mont_const(RC_INT_CANON[i]);
  for (int i = 0; i < 24; ++i) { t.diag[i] = (uint32_t)(DIAG_CANON[i] % P); t.diag_q[i] = (uint32_t)(((uint64_t)t.diag[i] << 32) / P); }
  return t;
}
constexpr ConstTables HOST_TABLES = make_tables();

#if defined(__CUDACC__)
__constant__ ConstTables c_tables = make_tables();
#define ZKB_P2_TABLES (::zkb::p2::c_tables)
#endif

ZKB_HD uint32_t mulhi32(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
  return __umulhi(a, b);
#else
  return (uint32_t)(((uint64_t)a * b) >> 32);
#endif
}
// x * d mod P in [0, 2P) for any 32-bit x, given dq = floor(d * 2^32 / P)
ZKB_HD uint32_t shoup_mul_lazy(uint32_t x, uint32_t d, uint32_t dq) { return x * d - mulhi32(x, dq) * P; }
// sum of 12 words (each < 2^32) reduced to a canonical element: the 64-bit total is < 12 * 2^32, its high word h <= 11,
// and 2^32 == R_MOD_P (mod P) with 11 * R_MOD_P < 2P
ZKB_HD uint32_t sum12(const uint32_t* s) {
  uint64_t acc = 0;
#pragma unroll
  for (int i = 0; i < 12; ++i) acc += s[i];
  return add_mod(reduce_2p(reduce_2p((uint32_t)acc)), reduce_2p((uint32_t)(acc >> 32) * R_MOD_P));
}

// Pipe assignment (sm_100a).  ncu on k_hash_rows: the fmaheavy pipe (every IMAD form) is ~90 % busy and the ALU pipe is the
// other half of a balanced pair (VIADDMNMX takes two ALU issue slots), yet ptxas turns ~45 % of the plain two-input
// additions into IMAD.IADD.  `min(a + b, ones)` with ones = 0xffffffff held in a uniform register (a launch argument, so it
// cannot be folded) compiles to ONE VIADDMNMX.U32 Rd, Ra, +-Rb, URones: a two-input add that stays on the ALU pipe.  Used for
// the s-box's canonical products, the round-constant additions and the partial rounds (+6.5 % permutations/s); the
// external linear layer is left to ptxas -- forcing its 128 additions per round onto the ALU pipe overloads it (-9 %).
// Measurements: tools/ubench/p2_alu_adds.cu, profiles/r1_r_ubench_p2_alu_adds.txt.  On the host `ones` is a constant.
// (add_alu / mont_mul_alu live in field.cuh: the NTT butterflies use them too.)
ZKB_HD uint32_t sbox7(uint32_t x, uint32_t ones = 0xffffffffu) {
  uint32_t x2 = mont_mul_alu(x, x, ones);
  uint32_t x4 = mont_mul_lazy(x2, x2);     // < 2P
  uint32_t x6 = mont_mul_lazy(x4, x2);     // x4 < 2P, x2 < P
  return mont_mul_alu(x6, x, ones);        // x6 < 2P, x < P
}

// 4x4 MDS block [5 7 1 3; 4 6 1 1; 1 3 5 7; 1 1 4 6]
ZKB_HD void m4(uint32_t& x0, uint32_t& x1, uint32_t& x2, uint32_t& x3) {
  uint32_t t0 = add_mod(x0, x1), t1 = add_mod(x2, x3);
  uint32_t t2 = add_mod(add_mod(x1, x1), t1), t3 = add_mod(add_mod(x3, x3), t0);
  uint32_t t1_2 = add_mod(t1, t1), t0_2 = add_mod(t0, t0);
  uint32_t t4 = add_mod(add_mod(t1_2, t1_2), t3), t5 = add_mod(add_mod(t0_2, t0_2), t2);
  x0 = add_mod(t3, t5); x1 = t5; x2 = add_mod(t2, t4); x3 = t4;
}
ZKB_HD void m_ext(uint32_t* s) {
#pragma unroll
  for (int c = 0; c < 6; ++c) m4(s[4 * c], s[4 * c + 1], s[4 * c + 2], s[4 * c + 3]);
  uint32_t sums[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    uint32_t a = add_mod(s[k], s[4 + k]), b = add_mod(s[8 + k], s[12 + k]), c = add_mod(s[16 + k], s[20 + k]);
    sums[k] = add_mod(add_mod(a, b), c);
  }
#pragma unroll
  for (int i = 0; i < 24; ++i) s[i] = add_mod(s[i], sums[i & 3]);
}

template <typename Tables>
ZKB_HD void permute(uint32_t* s, const Tables& T, const uint32_t ones = 0xffffffffu) {
  m_ext(s);
#pragma unroll 1
  for (int r = 0; r < 4; ++r) {
#pragma unroll
    for (int i = 0; i < 24; ++i) s[i] = sbox7(reduce_2p(add_alu(s[i], T.ext[r * 24 + i], ones)), ones);
    m_ext(s);
  }
  // Partial rounds.  Cells are kept LAZY (in [0, 2P)) across these rounds: only cell 0 is made canonical for its
  // s-box, the row sum is taken over the lazy words, and each cell becomes tot + d_i * s_i with the product reduced
  // once to [0, P) -- so tot + product < 2P again with no second correction.
#pragma unroll 1
  for (int r = 0; r < 21; ++r) {
    s[0] = sbox7(reduce_2p(add_alu(reduce_2p(s[0]), T.in[r], ones)), ones);
    uint32_t tot = reduce_2p(add_alu(sum12(s), sum12(s + 12), ones));
#pragma unroll
    for (int i = 0; i < 24; ++i) s[i] = add_alu(tot, reduce_2p(shoup_mul_lazy(s[i], T.diag[i], T.diag_q[i])), ones);
  }
#pragma unroll
  for (int i = 0; i < 24; ++i) s[i] = reduce_2p(s[i]);
#pragma unroll 1
  for (int r = 4; r < 8; ++r) {
#pragma unroll
    for (int i = 0; i < 24; ++i) s[i] = sbox7(reduce_2p(add_alu(s[i], T.ext[r * 24 + i], ones)), ones);
    m_ext(s);
  }
}

inline void permute_host(uint32_t* s) { permute(s, HOST_TABLES); }

} }  // namespace zkb::p2
