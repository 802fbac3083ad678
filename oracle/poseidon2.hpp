// ORACLE (test infrastructure, NOT product code) -- parity unpinned vs. the Rust crates.
//
// CPU restatement of risc0-zkp 1.2.5 `core::hash::poseidon2::{mod.rs, consts.rs, rng.rs}`
// (pinned by /root/reference/Cargo.lock:5057-5085; un-vendored).  Algorithm per SURVEY.md App. B:
// BabyBear, t = 24, x^7, R_F = 8, R_P = 21, HorizenLabs instance.  Round constants are REGENERATED
// here from the Poseidon Grain LFSR (not copied from a table) and the permutation is pinned by the
// upstream known-answer test `poseidon2_test_vectors` (App. B.2) in tests/test_oracle_kat.py.
#pragma once
#include "field.hpp"
#include <vector>
#include <cstring>

namespace orc {

static const int CELLS = 24, CELLS_RATE = 16, CELLS_OUT = 8, ROUNDS_HALF_FULL = 4, ROUNDS_PARTIAL = 21;
static const int N_ROUND_CONSTANTS = 2 * ROUNDS_HALF_FULL * CELLS + ROUNDS_PARTIAL;   // 213

// M_INT_DIAG_HZN (diag(M_I) - 1), canonical values; not part of the Grain stream (App. B.1).
static const uint32_t M_INT_DIAG_HZN[CELLS] = {
    0x409133f0, 0x1667a8a1, 0x06a6c7b6, 0x6f53160e, 0x273b11d1, 0x03176c5d, 0x72f9bbf9, 0x73ceba91,
    0x5cdef81d, 0x01393285, 0x46daee06, 0x065d7ba6, 0x52d72d6f, 0x05dd05e0, 0x3bab4b63, 0x6ada3842,
    0x2fc5fbec, 0x770d61b0, 0x5715aae9, 0x03ef0e90, 0x75b6c770, 0x242adf5f, 0x00d0ca4c, 0x36c0e388};

// Poseidon Grain LFSR (80-bit), parameters field=1, sbox=0, n=31, t=24, R_F=8, R_P=21.
class GrainLfsr {
  uint8_t s[80];
  int clock() {
    int nb = s[62] ^ s[51] ^ s[38] ^ s[23] ^ s[13] ^ s[0];
    memmove(s, s + 1, 79);
    s[79] = (uint8_t)nb;
    return nb;
  }
  void put(int& pos, uint32_t val, int bits) {
    for (int i = bits - 1; i >= 0; --i) s[pos++] = (val >> i) & 1;
  }
 public:
  GrainLfsr() {
    int pos = 0;
    put(pos, 1, 2); put(pos, 0, 4); put(pos, 31, 12); put(pos, 24, 12); put(pos, 8, 10); put(pos, 21, 10);
    while (pos < 80) s[pos++] = 1;
    for (int i = 0; i < 160; ++i) clock();
  }
  int next_bit() {   // self-shrinking filter
    for (;;) {
      int a = clock();
      int b = clock();
      if (a) return b;
    }
  }
  uint32_t next_const() {
    for (;;) {
      uint32_t v = 0;
      for (int i = 0; i < 31; ++i) v = (v << 1) | (uint32_t)next_bit();
      if (v < P) return v;
    }
  }
};

struct Poseidon2Consts {
  uint32_t canonical[N_ROUND_CONSTANTS];   // as generated (for the sha256 / hex pin)
  Fp ext[2 * ROUNDS_HALF_FULL][CELLS];     // full rounds 0..3 then 4..7
  Fp in[ROUNDS_PARTIAL];
  Fp diag[CELLS];
  Poseidon2Consts() {
    GrainLfsr g;
    for (int i = 0; i < N_ROUND_CONSTANTS; ++i) canonical[i] = g.next_const();
    int k = 0;
    for (int r = 0; r < ROUNDS_HALF_FULL; ++r) for (int i = 0; i < CELLS; ++i) ext[r][i] = Fp::from(canonical[k++]);
    for (int r = 0; r < ROUNDS_PARTIAL; ++r) in[r] = Fp::from(canonical[k++]);
    for (int r = 0; r < ROUNDS_HALF_FULL; ++r) for (int i = 0; i < CELLS; ++i) ext[ROUNDS_HALF_FULL + r][i] = Fp::from(canonical[k++]);
    for (int i = 0; i < CELLS; ++i) diag[i] = Fp::from(M_INT_DIAG_HZN[i]);
  }
};
static inline const Poseidon2Consts& p2c() { static Poseidon2Consts c; return c; }

static inline Fp sbox7(Fp x) { Fp x2 = x * x, x4 = x2 * x2, x6 = x4 * x2; return x6 * x; }

static inline void m_ext(Fp* s) {
  Fp o[CELLS];
  for (int ch = 0; ch < CELLS / 4; ++ch) {
    Fp x0 = s[4 * ch], x1 = s[4 * ch + 1], x2 = s[4 * ch + 2], x3 = s[4 * ch + 3];
    Fp t0 = x0 + x1, t1 = x2 + x3;
    Fp t2 = x1 + x1 + t1, t3 = x3 + x3 + t0;
    Fp t1_4 = t1 + t1; t1_4 = t1_4 + t1_4;
    Fp t0_4 = t0 + t0; t0_4 = t0_4 + t0_4;
    Fp t4 = t1_4 + t3, t5 = t0_4 + t2;
    Fp t6 = t3 + t5, t7 = t2 + t4;
    o[4 * ch] = t6; o[4 * ch + 1] = t5; o[4 * ch + 2] = t7; o[4 * ch + 3] = t4;
  }
  Fp sums[4];
  for (int k = 0; k < 4; ++k) { Fp a; for (int ch = 0; ch < CELLS / 4; ++ch) a += o[4 * ch + k]; sums[k] = a; }
  for (int i = 0; i < CELLS; ++i) s[i] = o[i] + sums[i % 4];
}
static inline void m_int(Fp* s) {
  const Poseidon2Consts& c = p2c();
  Fp tot; for (int i = 0; i < CELLS; ++i) tot += s[i];
  for (int i = 0; i < CELLS; ++i) s[i] = tot + c.diag[i] * s[i];
}
// poseidon2_mix
static inline void poseidon2_mix(Fp* s) {
  const Poseidon2Consts& c = p2c();
  m_ext(s);
  for (int r = 0; r < ROUNDS_HALF_FULL; ++r) {
    for (int i = 0; i < CELLS; ++i) s[i] = sbox7(s[i] + c.ext[r][i]);
    m_ext(s);
  }
  for (int r = 0; r < ROUNDS_PARTIAL; ++r) {
    s[0] = sbox7(s[0] + c.in[r]);
    m_int(s);
  }
  for (int r = 0; r < ROUNDS_HALF_FULL; ++r) {
    for (int i = 0; i < CELLS; ++i) s[i] = sbox7(s[i] + c.ext[ROUNDS_HALF_FULL + r][i]);
    m_ext(s);
  }
}

struct Digest {
  uint32_t w[DIGEST_WORDS];
  bool operator==(const Digest& o) const { return memcmp(w, o.w, sizeof(w)) == 0; }
};

// unpadded_hash: rate-16 OVERWRITE-mode sponge, zero pad, count==0 -> one mix (App. B.3).
class Sponge {
  Fp st[CELLS]; int unmixed = 0; size_t count = 0;
 public:
  void absorb(Fp v) {
    st[unmixed] = v; ++unmixed; ++count;
    if (unmixed == CELLS_RATE) { poseidon2_mix(st); unmixed = 0; }
  }
  Digest finish() {
    if (unmixed != 0 || count == 0) {
      for (int i = unmixed; i < CELLS_RATE; ++i) st[i] = Fp();
      poseidon2_mix(st);
    }
    Digest d; for (int i = 0; i < CELLS_OUT; ++i) d.w[i] = st[i].v;
    return d;
  }
};
static inline Digest hash_elem_slice(const Fp* s, size_t n) { Sponge sp; for (size_t i = 0; i < n; ++i) sp.absorb(s[i]); return sp.finish(); }
static inline Digest hash_ext_elem_slice(const Fp4* s, size_t n) {
  Sponge sp; for (size_t i = 0; i < n; ++i) for (int j = 0; j < 4; ++j) sp.absorb(s[i].c[j]); return sp.finish();
}
static inline Digest hash_pair(const Digest& a, const Digest& b) {
  Sponge sp;
  for (int i = 0; i < 8; ++i) sp.absorb(Fp::raw(a.w[i]));
  for (int i = 0; i < 8; ++i) sp.absorb(Fp::raw(b.w[i]));
  return sp.finish();
}

// Poseidon2Rng (Fiat-Shamir), App. B.4.
class Poseidon2Rng {
  Fp cells[CELLS]; int pool_used = 0;
 public:
  void mix(const Digest& d) {
    if (pool_used != 0) { poseidon2_mix(cells); pool_used = 0; }
    for (int i = 0; i < CELLS_OUT; ++i) cells[i] += Fp::raw(d.w[i]);
    poseidon2_mix(cells);
  }
  Fp random_elem() {
    if (pool_used == CELLS_RATE) { poseidon2_mix(cells); pool_used = 0; }
    return cells[pool_used++];
  }
  Fp4 random_ext_elem() { Fp a = random_elem(), b = random_elem(), c = random_elem(), d = random_elem(); return Fp4(a, b, c, d); }
  uint32_t random_bits(int bits) {
    uint32_t val = random_elem().as_u32();
    for (int i = 0; i < 3; ++i) { uint32_t nv = random_elem().as_u32(); if (val == 0) val = nv; }
    return val & (uint32_t)(((uint64_t)1 << bits) - 1);
  }
};

}  // namespace orc
