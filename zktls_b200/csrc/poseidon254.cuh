// Poseidon over the BN254 scalar field (t = 3, x^5, R_F = 8, R_P = 57): the `poseidon_254` hash suite's permutation.
//
// First slice of SURVEY.md 8(f)-4 (risc0-zkp 1.2.5 `core/hash/poseidon_254`, un-vendored; used by identity_p254 so that the
// Groth16 circuit can verify the last STARK: the instance is circomlib's `poseidon` with 2 inputs).  PINNED, unlike most of this
// repository's upstream facts: the 195 round constants and the Cauchy MDS matrix are REGENERATED here from the Poseidon reference
// procedure (Grain LFSR, field = 1, sbox = 0, n = 254, t = 3, R_F = 8, R_P = 57; matrix entries 1 / (x_i + y_j) from the same stream),
// and the result reproduces circomlib's public known answers poseidon([1,2]) and poseidon([3,4]) (tests/test_poseidon254.py; the
// library refuses to hash if its own start-up check of poseidon([1,2]) fails).
// Elements are 8 x u32 little-endian limbs in Montgomery form (R = 2^256) inside the permutation; digests hold the canonical value.
#pragma once
#include <cstdint>
#include "field.cuh"

namespace zkb { namespace p254 {

constexpr int T = 3, RF = 8, RP = 57, N_RC = T * (RF + RP);
struct Fr { uint32_t l[8]; };
// p = 21888242871839275222246405745257275088548364400416034343698204186575808495617
#define ZKB_P254_LIMBS {0xf0000001u, 0x43e1f593u, 0x79b97091u, 0x2833e848u, 0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u}
constexpr uint32_t P_INV_NEG = 0xefffffffu;                       // -p^-1 mod 2^32

ZKB_HD Fr r2() { Fr x = {{0xae216da7u, 0x1bb8e645u, 0xe35c59e3u, 0x53fe3ab1u, 0x53bb8085u, 0x8c49833du, 0x7f4e44a5u, 0x0216d0b1u}}; return x; }       // 2^512 mod p
ZKB_HD Fr one_m() { Fr x = {{0x4ffffffbu, 0xac96341cu, 0x9f60cd29u, 0x36fc7695u, 0x7879462eu, 0x666ea36fu, 0x9a07df2fu, 0x0e0a77c1u}}; return x; }    // 2^256 mod p
ZKB_HD uint32_t plimb(int i) { const uint32_t pl[8] = ZKB_P254_LIMBS; return pl[i]; }
ZKB_HD bool geq_p(const uint32_t* a) {
#pragma unroll
  for (int i = 7; i >= 0; --i) { uint32_t pi = plimb(i); if (a[i] != pi) return a[i] > pi; }
  return true;
}
ZKB_HD void sub_p(uint32_t* a) {
  uint64_t borrow = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) { uint64_t d = (uint64_t)a[i] - plimb(i) - borrow; a[i] = (uint32_t)d; borrow = (d >> 32) & 1u; }
}
ZKB_HD Fr add(const Fr& a, const Fr& b) {
  Fr r; uint64_t c = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) { c += (uint64_t)a.l[i] + b.l[i]; r.l[i] = (uint32_t)c; c >>= 32; }      // < 2p < 2^255: no carry out
  if (geq_p(r.l)) sub_p(r.l);
  return r;
}
// Montgomery product (CIOS, 32-bit limbs)
ZKB_HD Fr mul(const Fr& a, const Fr& b) {
  uint32_t t[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    uint64_t c = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) { uint64_t x = (uint64_t)a.l[j] * b.l[i] + t[j] + c; t[j] = (uint32_t)x; c = x >> 32; }
    uint64_t x = (uint64_t)t[8] + c; t[8] = (uint32_t)x; t[9] = (uint32_t)(x >> 32);
    const uint32_t m = t[0] * P_INV_NEG;
    x = (uint64_t)m * plimb(0) + t[0]; c = x >> 32;
#pragma unroll
    for (int j = 1; j < 8; ++j) { x = (uint64_t)m * plimb(j) + t[j] + c; t[j - 1] = (uint32_t)x; c = x >> 32; }
    x = (uint64_t)t[8] + c; t[7] = (uint32_t)x; t[8] = t[9] + (uint32_t)(x >> 32);
  }
  Fr r;
#pragma unroll
  for (int i = 0; i < 8; ++i) r.l[i] = t[i];
  if (t[8] || geq_p(r.l)) sub_p(r.l);
  return r;
}
ZKB_HD Fr to_mont(const Fr& canonical) { return mul(canonical, r2()); }
ZKB_HD Fr from_mont(const Fr& m) { Fr one = {{1, 0, 0, 0, 0, 0, 0, 0}}; return mul(m, one); }
ZKB_HD Fr pow5(const Fr& x) { Fr x2 = mul(x, x), x4 = mul(x2, x2); return mul(x4, x); }

struct Consts { Fr rc[N_RC]; Fr mds[T][T]; };      // Montgomery form

// one permutation; state in Montgomery form
template <typename C>
ZKB_HD void permute(Fr* s, const C& k) {
  int r = 0;
#pragma unroll 1
  for (int round = 0; round < RF + RP; ++round) {
#pragma unroll
    for (int i = 0; i < T; ++i) s[i] = add(s[i], k.rc[r + i]);
    r += T;
    if (round < RF / 2 || round >= RF / 2 + RP) {
#pragma unroll
      for (int i = 0; i < T; ++i) s[i] = pow5(s[i]);
    } else {
      s[0] = pow5(s[0]);
    }
    Fr o[T];
#pragma unroll
    for (int i = 0; i < T; ++i) o[i] = add(add(mul(k.mds[i][0], s[0]), mul(k.mds[i][1], s[1])), mul(k.mds[i][2], s[2]));
#pragma unroll
    for (int i = 0; i < T; ++i) s[i] = o[i];
  }
}

// digest (8 little-endian u32 words = a 256-bit integer, reduced mod p if needed) <-> Montgomery element
ZKB_HD Fr digest_to_fr(const uint32_t* w) {
  Fr x;
#pragma unroll
  for (int i = 0; i < 8; ++i) x.l[i] = w[i];
  while (geq_p(x.l)) sub_p(x.l);              // at most 5 times for an arbitrary 256-bit word; never for a digest this suite produced
  return to_mont(x);
}
ZKB_HD void fr_to_digest(uint32_t* w, const Fr& m) {
  Fr x = from_mont(m);
#pragma unroll
  for (int i = 0; i < 8; ++i) w[i] = x.l[i];
}

} }  // namespace zkb::p254
