// BabyBear field arithmetic for the B200 backend (host + device).
//
// Replaces risc0-core 1.2.5 `field::baby_bear::{Elem, ExtElem}` (Cargo.lock pin at
// /root/reference/Cargo.lock:5008-5017) and risc0-sys `cxx/{fp.h,fpext.h}`: P = 15*2^27+1, elements are
// Montgomery words (R = 2^32), Fp4 = Fp[x]/(x^4+11).  Every value stored to memory is canonical (< P), so
// results are bit-identical with any other correct implementation.
#pragma once
#include <cstdint>
#include <cstddef>

#if defined(__CUDACC__)
#define ZKB_HD __host__ __device__ __forceinline__
#define ZKB_D __device__ __forceinline__
#else
#define ZKB_HD inline
#define ZKB_D inline
#endif

namespace zkb {

constexpr uint32_t P = 2013265921u;          // 0x78000001
constexpr uint32_t P_NEG_INV = 0x77ffffffu;  // -P^{-1} mod 2^32
constexpr uint32_t R_MOD_P = 0x0ffffffeu;    // Montgomery form of 1
constexpr uint32_t R2_MOD_P = 1172168163u;   // 2^64 mod P
constexpr uint32_t INVALID = 0xffffffffu;

constexpr uint32_t P_INV = 0x88000001u;      // P^{-1} mod 2^32 (= 2^31 + 2^27 + 1)
ZKB_HD uint32_t mul_hi32(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
  return __umulhi(a, b);
#else
  return (uint32_t)(((uint64_t)a * b) >> 32);
#endif
}
// Montgomery reduction of a 64-bit value t < P * 2^32, SUBTRACTIVE form: with m = lo(t) * P^{-1} the low words of t and
// m * P agree, so t / 2^32 = hi(t) - hi(m * P) exactly, a value in (-P, P).  On sm_100a this is IMAD + a plain IMAD.HI + one
// IADD3; the additive form (t + m' P) >> 32 compiles to an IMAD.HI with a 64-bit addend, which measures ~4 % slower over
// a whole Poseidon2 permutation (profiles/r1_ubench_p2_variants.txt).
//   mont_redc_lazy: hi - u + P, in (0, 2P)  (one 3-input add);   reduce_2p brings it to [0, P).
ZKB_HD uint32_t mont_redc_lazy(uint64_t t) {
  uint32_t m = (uint32_t)t * P_INV;
  return (uint32_t)(t >> 32) - mul_hi32(m, P) + P;
}
ZKB_HD uint32_t reduce_2p(uint32_t x) {      // [0, 2P) -> [0, P)
  uint32_t y = x - P;
  return y < x ? y : x;                        // unsigned min: VIADDMNMX.U32
}
// canonical product; valid when a * b < P * 2^32 (e.g. a < 2P, b < P): r = hi - u in (-P, P), then one conditional + P
ZKB_HD uint32_t mont_mul(uint32_t a, uint32_t b) {
  uint64_t t = (uint64_t)a * b;
  uint32_t m = (uint32_t)t * P_INV;
  uint32_t r = (uint32_t)(t >> 32) - mul_hi32(m, P);
  uint32_t y = r + P;
  return y < r ? y : r;                        // r "negative" (wrapped)  <=>  r + P wraps back below r
}
// product left in (0, 2P); same precondition
ZKB_HD uint32_t mont_mul_lazy(uint32_t a, uint32_t b) { return mont_redc_lazy((uint64_t)a * b); }
ZKB_HD uint32_t add_mod(uint32_t a, uint32_t b) { return reduce_2p(a + b); }
// ALU-pinned forms (see the pipe-assignment note in poseidon2.cuh): `ones` must be 0xffffffff and must reach the kernel as a
// launch argument, so that min(a + b, ones) survives as ONE VIADDMNMX instead of being folded into an add ptxas may turn into IMAD.IADD.
ZKB_HD uint32_t umin32(uint32_t a, uint32_t b) { return a < b ? a : b; }
ZKB_HD uint32_t add_alu(uint32_t a, uint32_t b, uint32_t ones) { return umin32(a + b, ones); }
ZKB_HD uint32_t mont_mul_alu(uint32_t a, uint32_t b, uint32_t ones) {      // mont_mul with the final subtraction on the ALU pipe
  uint64_t t = (uint64_t)a * b;
  uint32_t m = (uint32_t)t * P_INV;
  uint32_t r = umin32((uint32_t)(t >> 32) - mul_hi32(m, P), ones);
  uint32_t y = r + P;
  return y < r ? y : r;
}

ZKB_HD uint32_t sub_mod(uint32_t a, uint32_t b) {
  uint32_t d = a - b;
  uint32_t e = d + P;
  return e < d ? e : d;                        // a < b  <=>  d wrapped  <=>  d + P wraps back below d
}

struct Fp {
  uint32_t v;   // Montgomery word, canonical
  ZKB_HD Fp() : v(0) {}
  ZKB_HD explicit Fp(uint32_t raw) : v(raw) {}
  static ZKB_HD Fp raw(uint32_t w) { return Fp(w); }
  static ZKB_HD Fp from(uint32_t x) { return Fp(mont_mul(R2_MOD_P, x % P)); }
  static ZKB_HD Fp one() { return Fp(R_MOD_P); }
  ZKB_HD uint32_t as_u32() const { return mont_mul(1u, v); }
  ZKB_HD Fp operator+(Fp o) const { return Fp(add_mod(v, o.v)); }
  ZKB_HD Fp operator-(Fp o) const { return Fp(sub_mod(v, o.v)); }
  ZKB_HD Fp operator*(Fp o) const { return Fp(mont_mul(v, o.v)); }
  ZKB_HD Fp operator-() const { return Fp(sub_mod(0u, v)); }
  ZKB_HD Fp& operator+=(Fp o) { v = add_mod(v, o.v); return *this; }
  ZKB_HD Fp& operator-=(Fp o) { v = sub_mod(v, o.v); return *this; }
  ZKB_HD Fp& operator*=(Fp o) { v = mont_mul(v, o.v); return *this; }
  ZKB_HD bool operator==(Fp o) const { return v == o.v; }
  ZKB_HD bool operator!=(Fp o) const { return v != o.v; }
};

ZKB_HD Fp pow(Fp x, uint64_t e) {
  Fp r = Fp::one();
  while (e) { if (e & 1) r *= x; x *= x; e >>= 1; }
  return r;
}
ZKB_HD Fp inv(Fp x) { return pow(x, P - 2); }

// Montgomery constants for 11 and P-11 (computed: 11 * 2^32 mod P, (P-11) * 2^32 mod P)
constexpr uint32_t mont_const(uint64_t x) { return (uint32_t)(((x % P) << 32) % P); }
constexpr uint32_t BETA = mont_const(11);
constexpr uint32_t NBETA = mont_const(P - 11);

struct Fp4 {
  Fp c[4];
  ZKB_HD Fp4() {}
  ZKB_HD Fp4(Fp a0, Fp a1, Fp a2, Fp a3) { c[0] = a0; c[1] = a1; c[2] = a2; c[3] = a3; }
  static ZKB_HD Fp4 from_base(Fp a) { return Fp4(a, Fp(), Fp(), Fp()); }
  static ZKB_HD Fp4 zero() { return Fp4(); }
  static ZKB_HD Fp4 one() { return from_base(Fp::one()); }
  static ZKB_HD Fp4 raw(uint32_t a, uint32_t b, uint32_t c_, uint32_t d) { return Fp4(Fp::raw(a), Fp::raw(b), Fp::raw(c_), Fp::raw(d)); }
  static ZKB_HD Fp4 load(const uint32_t* w) { return Fp4(Fp(w[0]), Fp(w[1]), Fp(w[2]), Fp(w[3])); }
  ZKB_HD void store(uint32_t* w) const { w[0] = c[0].v; w[1] = c[1].v; w[2] = c[2].v; w[3] = c[3].v; }
  ZKB_HD Fp4 operator+(const Fp4& o) const { return Fp4(c[0] + o.c[0], c[1] + o.c[1], c[2] + o.c[2], c[3] + o.c[3]); }
  ZKB_HD Fp4 operator-(const Fp4& o) const { return Fp4(c[0] - o.c[0], c[1] - o.c[1], c[2] - o.c[2], c[3] - o.c[3]); }
  ZKB_HD Fp4 operator*(Fp s) const { return Fp4(c[0] * s, c[1] * s, c[2] * s, c[3] * s); }
  ZKB_HD Fp4 operator*(const Fp4& o) const {
    const Fp nb(NBETA);
    const Fp* a = c; const Fp* b = o.c;
    return Fp4(a[0] * b[0] + nb * (a[1] * b[3] + a[2] * b[2] + a[3] * b[1]),
               a[0] * b[1] + a[1] * b[0] + nb * (a[2] * b[3] + a[3] * b[2]),
               a[0] * b[2] + a[1] * b[1] + a[2] * b[0] + nb * (a[3] * b[3]),
               a[0] * b[3] + a[1] * b[2] + a[2] * b[1] + a[3] * b[0]);
  }
  ZKB_HD Fp4& operator+=(const Fp4& o) { *this = *this + o; return *this; }
  ZKB_HD Fp4& operator-=(const Fp4& o) { *this = *this - o; return *this; }
  ZKB_HD Fp4& operator*=(const Fp4& o) { *this = *this * o; return *this; }
  ZKB_HD bool operator==(const Fp4& o) const { return c[0] == o.c[0] && c[1] == o.c[1] && c[2] == o.c[2] && c[3] == o.c[3]; }
  ZKB_HD bool operator!=(const Fp4& o) const { return !(*this == o); }
};

ZKB_HD Fp4 inv(const Fp4& x) {
  const Fp beta(BETA), nbeta(NBETA);
  const Fp* a = x.c;
  Fp b0 = a[0] * a[0] + beta * (a[1] * (a[3] + a[3]) - a[2] * a[2]);
  Fp b2 = a[0] * (a[2] + a[2]) - a[1] * a[1] + beta * (a[3] * a[3]);
  Fp cc = b0 * b0 + beta * b2 * b2;
  Fp ic = inv(cc);
  b0 *= ic; b2 *= ic;
  return Fp4(a[0] * b0 + beta * a[2] * b2, -(a[1] * b0) + nbeta * a[3] * b2, -(a[0] * b2) + a[2] * b0, a[1] * b2 - a[3] * b0);
}
ZKB_HD Fp4 pow(Fp4 x, uint64_t e) {
  Fp4 r = Fp4::one();
  while (e) { if (e & 1) r *= x; x *= x; e >>= 1; }
  return r;
}

ZKB_HD uint32_t bit_rev32(uint32_t x, int bits) {
#if defined(__CUDA_ARCH__)
  return bits == 0 ? 0u : (__brev(x) >> (32 - bits));
#else
  uint32_t r = 0;
  for (int i = 0; i < bits; ++i) { r = (r << 1) | (x & 1); x >>= 1; }
  return r;
#endif
}

// Protocol constants (risc0-zkp/src/lib.rs).
constexpr int MAX_ROU_PO2 = 27;
constexpr size_t INV_RATE = 4, QUERIES = 50, FRI_FOLD = 16, FRI_FOLD_PO2 = 4, FRI_MIN_DEGREE = 256;
constexpr size_t EXT_SIZE = 4, CHECK_SIZE = 16, DIGEST_WORDS = 8;
constexpr int MIN_PO2 = 1, MAX_PO2 = 26;   // NTT sizes accepted by the operators (MAX_CYCLES_PO2 24 + 2 expand bits)

#if defined(__CUDACC__)
ZKB_HD Fp4 ld4(const uint4& v) { return Fp4::raw(v.x, v.y, v.z, v.w); }
ZKB_HD uint4 st4(const Fp4& r) { return make_uint4(r.c[0].v, r.c[1].v, r.c[2].v, r.c[3].v); }
#endif

}  // namespace zkb
