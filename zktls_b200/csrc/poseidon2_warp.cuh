// Warp-cooperative Poseidon2 permutation (t = 24): cell i of the state lives in lane i of one warp (lanes 24..31 idle).
//
// For the LATENCY-bound top of a Merkle tree, where there are fewer permutations than the GPU has warps: the thread-per-state
// permutation of poseidon2.cuh is a 10 k-instruction dependent stream (~19 us), and a level with 2^11 outputs fills less than one
// SM's worth of threads.  Spread over 24 lanes the 24 s-boxes of a full round run side by side, the 4x4 MDS blocks and the column
// sums become a handful of shuffles, and the partial rounds' row sum is two REDUX (warp-wide integer add) instructions:
// ~4.5 k dependent cycles (~2.3 us) per permutation.  Same field arithmetic, same constants, bit-identical results
// (tests/test_hal_parity.py::test_merkle_*, tests/test_full_size_parity.py).
#pragma once
#include "poseidon2.cuh"

namespace zkb { namespace p2w {

constexpr unsigned FULL = 0xffffffffu;
// Montgomery forms of the M4 coefficients, row j = lane & 3 of [5 7 1 3; 4 6 1 1; 1 3 5 7; 1 1 4 6]
__device__ __forceinline__ void m4_row(int j, uint32_t (&c)[4]) {
  const uint32_t one = mont_const(1), three = mont_const(3), four = mont_const(4), five = mont_const(5), six = mont_const(6), seven = mont_const(7);
  if (j == 0) { c[0] = five; c[1] = seven; c[2] = one; c[3] = three; }
  else if (j == 1) { c[0] = four; c[1] = six; c[2] = one; c[3] = one; }
  else if (j == 2) { c[0] = one; c[1] = three; c[2] = five; c[3] = seven; }
  else { c[0] = one; c[1] = one; c[2] = four; c[3] = six; }
}
struct Lane {
  uint32_t lane, rc_ext[8], diag, diag_q, m4c[4];
  __device__ __forceinline__ void init() {
    lane = threadIdx.x & 31u;
    const uint32_t i = lane < 24u ? lane : 0u;
#pragma unroll
    for (int r = 0; r < 8; ++r) rc_ext[r] = lane < 24u ? ZKB_P2_TABLES.ext[r * 24 + i] : 0u;
    diag = ZKB_P2_TABLES.diag[i]; diag_q = ZKB_P2_TABLES.diag_q[i];
    m4_row((int)(lane & 3u), m4c);
  }
};
// external linear layer: s (canonical) -> M_ext(s) for this lane's cell (canonical)
__device__ __forceinline__ uint32_t m_ext(const Lane& L, uint32_t s) {
  const uint32_t base = L.lane & ~3u;
  uint64_t acc = 0;
#pragma unroll
  for (int k = 0; k < 4; ++k) acc += (uint64_t)L.m4c[k] * __shfl_sync(FULL, s, (int)(base + k));       // 4 P^2 < 2^64
  const uint32_t o = reduce_2p(mont_redc_lazy(acc));          // sum_k c_k x_k: the Montgomery factor of the coefficient cancels
  uint32_t tot = o;
#pragma unroll
  for (int k = 1; k < 6; ++k) { uint32_t src = L.lane + 4u * k; src = src >= 24u ? src - 24u : src; tot = add_mod(tot, __shfl_sync(FULL, o, (int)src)); }
  return L.lane < 24u ? add_mod(o, tot) : 0u;
}
// sum of the 24 cells (canonical words) mod P, in every lane
__device__ __forceinline__ uint32_t row_sum(uint32_t s) {
  const uint32_t lo = __reduce_add_sync(FULL, s & 0xffffu), hi = __reduce_add_sync(FULL, s >> 16);      // < 24 * 2^16 each: no overflow
  const uint64_t t = ((uint64_t)hi << 16) + lo;                                                        // < 2^36
  return add_mod(reduce_2p(reduce_2p((uint32_t)t)), reduce_2p((uint32_t)(t >> 32) * R_MOD_P));
}
// the whole permutation; s = this lane's cell (canonical; lanes >= 24 pass 0 and get 0 back)
__device__ __forceinline__ uint32_t permute(const Lane& L, uint32_t s) {
  s = m_ext(L, s);
#pragma unroll 1
  for (int r = 0; r < 4; ++r) s = m_ext(L, p2::sbox7(add_mod(s, L.rc_ext[r])));
#pragma unroll 1
  for (int r = 0; r < 21; ++r) {
    const uint32_t others = row_sum(L.lane == 0u ? 0u : s);                          // independent of the s-box below: overlaps its latency
    const uint32_t s0 = p2::sbox7(add_mod(s, ZKB_P2_TABLES.in[r]));                  // every lane computes, lane 0 keeps
    const uint32_t s0b = __shfl_sync(FULL, s0, 0);
    if (L.lane == 0u) s = s0;
    const uint32_t tot = add_mod(others, s0b);
    s = L.lane < 24u ? add_mod(tot, reduce_2p(p2::shoup_mul_lazy(s, L.diag, L.diag_q))) : 0u;
  }
#pragma unroll 1
  for (int r = 4; r < 8; ++r) s = m_ext(L, p2::sbox7(add_mod(s, L.rc_ext[r])));
  return s;
}

} }  // namespace zkb::p2w
