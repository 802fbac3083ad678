// Host-side Poseidon2 sponge and Fiat-Shamir rng shared by the prover (prover.cu) and the verifier (verify.cpp).
// Stand-ins for risc0-zkp `core::hash::poseidon2::{unpadded_hash, Poseidon2HashFn, rng::Poseidon2Rng}`.
#pragma once
#include <cstring>
#include "field.cuh"
#include "poseidon2.cuh"

namespace zkb {

// ---- host Poseidon2 sponge + Fiat-Shamir rng (App. B.3, B.4) ----------------------------------------------
struct Digest { uint32_t w[8]; };

class HostSponge {
  uint32_t st[24] = {0}; int unmixed = 0; size_t count = 0;
 public:
  void absorb(uint32_t w) {
    st[unmixed++] = w; ++count;
    if (unmixed == p2::RATE) { p2::permute_host(st); unmixed = 0; }
  }
  Digest finish() {
    if (unmixed != 0 || count == 0) { for (int i = unmixed; i < p2::RATE; ++i) st[i] = 0; p2::permute_host(st); }
    Digest d; memcpy(d.w, st, 32); return d;
  }
};
static Digest hash_words(const uint32_t* w, size_t n) { HostSponge s; for (size_t i = 0; i < n; ++i) s.absorb(w[i]); return s.finish(); }

class HostRng {
  uint32_t cells[24] = {0}; int pool_used = 0;
 public:
  void mix(const Digest& d) {
    if (pool_used != 0) { p2::permute_host(cells); pool_used = 0; }
    for (int i = 0; i < 8; ++i) cells[i] = add_mod(cells[i], d.w[i]);
    p2::permute_host(cells);
  }
  Fp random_elem() {
    if (pool_used == p2::RATE) { p2::permute_host(cells); pool_used = 0; }
    return Fp::raw(cells[pool_used++]);
  }
  Fp4 random_ext_elem() { Fp a = random_elem(), b = random_elem(), c = random_elem(), d = random_elem(); return Fp4(a, b, c, d); }
  uint32_t random_bits(int bits) {
    uint32_t val = random_elem().as_u32();
    for (int i = 0; i < 3; ++i) { uint32_t nv = random_elem().as_u32(); if (val == 0) val = nv; }
    return val & (uint32_t)(((uint64_t)1 << bits) - 1);
  }
};

static inline Digest hash_pair_host(const Digest& a, const Digest& b) {
  uint32_t st[24] = {0};
  memcpy(st, a.w, 32); memcpy(st + 8, b.w, 32);
  p2::permute_host(st);
  Digest d; memcpy(d.w, st, 32); return d;
}

}  // namespace zkb
