// Poseidon2 permutation over BabyBear (t = 24) for host and device.
// Replaces risc0-zkp 1.2.5 `core::hash::poseidon2::poseidon2_mix` and sppark's `poseidon2` device code
// (un-vendored).  The state lives in 24 registers per thread.
#pragma once
#include "field.cuh"
#include "poseidon2_consts.h"

namespace zkb { namespace p2 {

// Montgomery-form constant tables, built at compile time from the canonical values.
struct ConstTables {
  uint32_t ext[8 * 24];
  uint32_t in[21];
  uint32_t diag[24];
};
constexpr ConstTables make_tables() {
  ConstTables t{};
  for (int i = 0; i < 8 * 24; ++i) t.ext[i] = mont_const(RC_EXT_CANON[i]);
  for (int i = 0; i < 21; ++i) t.in[i] = mont_const(RC_INT_CANON[i]);
  for (int i = 0; i < 24; ++i) t.diag[i] = mont_const(DIAG_CANON[i]);
  return t;
}
constexpr ConstTables HOST_TABLES = make_tables();

#if defined(__CUDACC__)
__constant__ ConstTables c_tables = make_tables();
#define ZKB_P2_TABLES (::zkb::p2::c_tables)
#endif

ZKB_HD uint32_t sbox7(uint32_t x) {
  uint32_t x2 = mont_mul(x, x);
  uint32_t x4 = mont_mul_lazy(x2, x2);     // < 2P
  uint32_t x6 = mont_mul_lazy(x4, x2);     // x4 < 2P, x2 < P
  return mont_mul(x6, x);                  // x6 < 2P, x < P
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
ZKB_HD void permute(uint32_t* s, const Tables& T) {
  m_ext(s);
#pragma unroll 1
  for (int r = 0; r < 4; ++r) {
#pragma unroll
    for (int i = 0; i < 24; ++i) s[i] = sbox7(add_mod(s[i], T.ext[r * 24 + i]));
    m_ext(s);
  }
#pragma unroll 1
  for (int r = 0; r < 21; ++r) {
    s[0] = sbox7(add_mod(s[0], T.in[r]));
    // tot = sum of all cells; 64-bit accumulate, one reduction
    uint64_t acc = 0;
#pragma unroll
    for (int i = 0; i < 24; ++i) acc += s[i];
    // acc < 24 P < 2^36: fold the high word with 2^32 == R_MOD_P (mod P); hi <= 11 so hi * R_MOD_P < 2P
    uint32_t lo = (uint32_t)acc, hi = (uint32_t)(acc >> 32);
    uint32_t tot = add_mod(reduce_2p(reduce_2p(lo)), reduce_2p(hi * R_MOD_P));
#pragma unroll
    for (int i = 0; i < 24; ++i) s[i] = add_mod(tot, mont_mul(T.diag[i], s[i]));
  }
#pragma unroll 1
  for (int r = 4; r < 8; ++r) {
#pragma unroll
    for (int i = 0; i < 24; ++i) s[i] = sbox7(add_mod(s[i], T.ext[r * 24 + i]));
    m_ext(s);
  }
}

inline void permute_host(uint32_t* s) { permute(s, HOST_TABLES); }

} }  // namespace zkb::p2
