// ORACLE (test infrastructure, NOT product code) -- parity unpinned vs. the Rust crates.
//
// CPU restatement of risc0-core 1.2.5 `field::baby_bear::{Elem, ExtElem}` (pinned by
// /root/reference/Cargo.lock:5008-5017; reached from the reference only through
// /root/reference/crates/guest-prover-r0/src/prover.rs:90).  The crate source is not vendored in
// the reference tree, so this follows the published algorithm as written out in SURVEY.md App. A
// and is pinned by the known-answer values of SURVEY.md App. A/F (tests/test_oracle_kat.py).
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use
// anything under oracle/.
#pragma once
#include <cstdint>
#include <cstddef>

namespace orc {

static const uint32_t P = 2013265921u;        // 15 * 2^27 + 1
static const uint32_t M_INV = 0x88000001u;    // P * M_INV == 1 mod 2^32
static const uint32_t R2 = 1172168163u;       // 2^64 mod P
static const uint32_t INVALID = 0xffffffffu;

// Montgomery product exactly as App. A.
static inline uint32_t mont_mul(uint32_t a, uint32_t b) {
  uint64_t o64 = (uint64_t)a * b;
  uint32_t low = 0u - (uint32_t)o64;
  uint32_t red = M_INV * low;
  o64 += (uint64_t)red * P;
  uint32_t ret = (uint32_t)(o64 >> 32);
  return ret >= P ? ret - P : ret;
}
static inline uint32_t f_add(uint32_t a, uint32_t b) { uint32_t s = a + b; return s >= P ? s - P : s; }
static inline uint32_t f_sub(uint32_t a, uint32_t b) { uint32_t d = a - b; return a < b ? d + P : d; }

// Fp: stored Montgomery word in `v` (always canonical, < P).
struct Fp {
  uint32_t v;
  Fp() : v(0) {}
  static Fp raw(uint32_t w) { Fp r; r.v = w; return r; }
  static Fp from(uint32_t x) { return raw(mont_mul(R2, x % P)); }   // encode canonical value
  uint32_t as_u32() const { return mont_mul(1, v); }               // decode
  Fp operator+(Fp o) const { return raw(f_add(v, o.v)); }
  Fp operator-(Fp o) const { return raw(f_sub(v, o.v)); }
  Fp operator*(Fp o) const { return raw(mont_mul(v, o.v)); }
  Fp operator-() const { return raw(f_sub(0, v)); }
  Fp& operator+=(Fp o) { v = f_add(v, o.v); return *this; }
  Fp& operator-=(Fp o) { v = f_sub(v, o.v); return *this; }
  Fp& operator*=(Fp o) { v = mont_mul(v, o.v); return *this; }
  bool operator==(Fp o) const { return v == o.v; }
  bool operator!=(Fp o) const { return v != o.v; }
};

static inline Fp f_pow(Fp x, uint64_t e) {
  Fp r = Fp::from(1);
  while (e) { if (e & 1) r *= x; x *= x; e >>= 1; }
  return r;
}
static inline Fp f_inv(Fp x) { return f_pow(x, P - 2); }   // inv(0) == 0

// Fp4 = Fp[x] / (x^4 + 11)
struct Fp4 {
  Fp c[4];
  Fp4() {}
  Fp4(Fp a0, Fp a1, Fp a2, Fp a3) { c[0] = a0; c[1] = a1; c[2] = a2; c[3] = a3; }
  static Fp4 from_base(Fp a) { return Fp4(a, Fp(), Fp(), Fp()); }
  static Fp4 zero() { return Fp4(); }
  static Fp4 one() { return from_base(Fp::from(1)); }
  Fp4 operator+(const Fp4& o) const { return Fp4(c[0] + o.c[0], c[1] + o.c[1], c[2] + o.c[2], c[3] + o.c[3]); }
  Fp4 operator-(const Fp4& o) const { return Fp4(c[0] - o.c[0], c[1] - o.c[1], c[2] - o.c[2], c[3] - o.c[3]); }
  Fp4 operator*(Fp s) const { return Fp4(c[0] * s, c[1] * s, c[2] * s, c[3] * s); }
  Fp4 operator*(const Fp4& o) const {
    const Fp NBETA = Fp::from(P - 11);
    const Fp* a = c; const Fp* b = o.c;
    return Fp4(a[0] * b[0] + NBETA * (a[1] * b[3] + a[2] * b[2] + a[3] * b[1]),
               a[0] * b[1] + a[1] * b[0] + NBETA * (a[2] * b[3] + a[3] * b[2]),
               a[0] * b[2] + a[1] * b[1] + a[2] * b[0] + NBETA * (a[3] * b[3]),
               a[0] * b[3] + a[1] * b[2] + a[2] * b[1] + a[3] * b[0]);
  }
  Fp4& operator+=(const Fp4& o) { *this = *this + o; return *this; }
  Fp4& operator-=(const Fp4& o) { *this = *this - o; return *this; }
  Fp4& operator*=(const Fp4& o) { *this = *this * o; return *this; }
  bool operator==(const Fp4& o) const { return c[0] == o.c[0] && c[1] == o.c[1] && c[2] == o.c[2] && c[3] == o.c[3]; }
  bool operator!=(const Fp4& o) const { return !(*this == o); }
};

static inline Fp4 f4_inv(const Fp4& x) {
  const Fp BETA = Fp::from(11), NBETA = Fp::from(P - 11);
  const Fp* a = x.c;
  Fp two = Fp::from(2);
  Fp b0 = a[0] * a[0] + BETA * (a[1] * (a[3] * two) - a[2] * a[2]);
  Fp b2 = a[0] * (a[2] * two) - a[1] * a[1] + BETA * (a[3] * a[3]);
  Fp cc = b0 * b0 + BETA * b2 * b2;
  Fp ic = f_inv(cc);
  b0 *= ic; b2 *= ic;
  return Fp4(a[0] * b0 + BETA * a[2] * b2,
             -(a[1] * b0) + NBETA * a[3] * b2,
             -(a[0] * b2) + a[2] * b0,
             a[1] * b2 - a[3] * b0);
}
static inline Fp4 f4_pow(Fp4 x, uint64_t e) {
  Fp4 r = Fp4::one();
  while (e) { if (e & 1) r *= x; x *= x; e >>= 1; }
  return r;
}

// Roots of unity: ROU_FWD[27] = 137, ROU_FWD[i] = 137^(2^(27-i)); ROU_REV = inverses.
static const int MAX_ROU_PO2 = 27;
struct RouTables {
  Fp fwd[MAX_ROU_PO2 + 1], rev[MAX_ROU_PO2 + 1];
  RouTables() {
    fwd[MAX_ROU_PO2] = Fp::from(137);
    for (int i = MAX_ROU_PO2 - 1; i >= 0; --i) fwd[i] = fwd[i + 1] * fwd[i + 1];
    for (int i = 0; i <= MAX_ROU_PO2; ++i) rev[i] = f_inv(fwd[i]);
  }
};
static inline const RouTables& rou() { static RouTables t; return t; }

static inline uint32_t bit_rev(uint32_t x, int bits) {
  uint32_t r = 0;
  for (int i = 0; i < bits; ++i) { r = (r << 1) | (x & 1); x >>= 1; }
  return r;
}
static inline int log2_exact(size_t n) { int k = 0; while (((size_t)1 << k) < n) ++k; return k; }

// Protocol constants (risc0-zkp/src/lib.rs).
static const size_t INV_RATE = 4, QUERIES = 50, FRI_FOLD = 16, FRI_FOLD_PO2 = 4, FRI_MIN_DEGREE = 256;
static const size_t EXT_SIZE = 4, CHECK_SIZE = 16, DIGEST_WORDS = 8;

}  // namespace orc
