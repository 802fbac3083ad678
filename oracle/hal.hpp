// ORACLE (test infrastructure, NOT product code) -- parity unpinned vs. the Rust crates.
//
// CPU restatement of risc0-zkp 1.2.5 `hal::cpu::CpuHal` and `core::{ntt.rs, poly.rs}` (un-vendored;
// /root/reference/Cargo.lock:5057-5085), operator by operator as specified in SURVEY.md App. C.
// Like CpuHal it is a scalar-Montgomery implementation parallelised with one thread per column
// (NTT family) or over rows (hashing, mixing, folding) -- rayon there, OpenMP here -- which is
// also what makes it the "port" CPU baseline of bench.py.
#pragma once
#include "field.hpp"
#include "poseidon2.hpp"
#include <vector>
#include <algorithm>

namespace orc {

// ---- core::ntt ----------------------------------------------------------------------------------
// rev_butterfly_N: DIF, natural in -> bit-reversed out, twiddles ROU_REV[N]^i (App. C.1).
static void rev_butterfly(Fp* io, int n_bits) {
  if (n_bits == 0) return;
  size_t half = (size_t)1 << (n_bits - 1);
  Fp step = rou().rev[n_bits], cur = Fp::from(1);
  for (size_t i = 0; i < half; ++i) {
    Fp a = io[i], b = io[i + half];
    io[i] = a + b;
    io[i + half] = (a - b) * cur;
    cur *= step;
  }
  rev_butterfly(io, n_bits - 1);
  rev_butterfly(io + half, n_bits - 1);
}
// fwd_butterfly_N skipping levels <= expand_bits: DIT, bit-reversed in -> natural out (App. C.3).
static void fwd_butterfly(Fp* io, int n_bits, int expand_bits) {
  if (n_bits == expand_bits) return;
  size_t half = (size_t)1 << (n_bits - 1);
  fwd_butterfly(io, n_bits - 1, expand_bits);
  fwd_butterfly(io + half, n_bits - 1, expand_bits);
  Fp step = rou().fwd[n_bits], cur = Fp::from(1);
  for (size_t i = 0; i < half; ++i) {
    Fp a = io[i], b = io[i + half] * cur;
    io[i] = a + b;
    io[i + half] = a - b;
    cur *= step;
  }
}
static void interpolate_ntt(Fp* io, size_t n) {
  int k = log2_exact(n);
  rev_butterfly(io, k);
  Fp norm = f_inv(Fp::from((uint32_t)n));
  for (size_t i = 0; i < n; ++i) io[i] *= norm;
}
static void bit_reverse(Fp* io, size_t n) {
  int k = log2_exact(n);
  for (size_t i = 0; i < n; ++i) { size_t r = bit_rev((uint32_t)i, k); if (i < r) std::swap(io[i], io[r]); }
}

// ---- Hal operators --------------------------------------------------------------------------------
static void batch_interpolate_ntt(Fp* io, size_t count, size_t n) {
#pragma omp parallel for schedule(dynamic, 1)
  for (size_t c = 0; c < count; ++c) interpolate_ntt(io + c * n, n);
}
static void zk_shift(Fp* io, size_t count, size_t n) {
  int k = log2_exact(n);
  Fp three = Fp::from(3);
#pragma omp parallel for schedule(static)
  for (size_t i = 0; i < n; ++i) {
    Fp s = f_pow(three, bit_rev((uint32_t)i, k));
    for (size_t c = 0; c < count; ++c) io[c * n + i] *= s;
  }
}
static void batch_expand(Fp* out, const Fp* in, size_t count, size_t n, int expand_bits) {
  size_t big = n << expand_bits;
#pragma omp parallel for schedule(dynamic, 1)
  for (size_t c = 0; c < count; ++c)
    for (size_t i = 0; i < big; ++i) out[c * big + i] = in[c * n + (i >> expand_bits)];
}
static void batch_evaluate_ntt(Fp* io, size_t count, size_t big, int expand_bits) {
  int k = log2_exact(big);
#pragma omp parallel for schedule(dynamic, 1)
  for (size_t c = 0; c < count; ++c) fwd_butterfly(io + c * big, k, expand_bits);
}
static void batch_expand_into_evaluate_ntt(Fp* out, const Fp* in, size_t count, size_t n, int expand_bits) {
  size_t big = n << expand_bits;
  int k = log2_exact(big);
#pragma omp parallel for schedule(dynamic, 1)
  for (size_t c = 0; c < count; ++c) {
    Fp* o = out + c * big;
    for (size_t i = 0; i < big; ++i) o[i] = in[c * n + (i >> expand_bits)];
    fwd_butterfly(o, k, expand_bits);
  }
}
static void batch_bit_reverse(Fp* io, size_t count, size_t n) {
#pragma omp parallel for schedule(dynamic, 1)
  for (size_t c = 0; c < count; ++c) bit_reverse(io + c * n, n);
}
static void hash_rows(Digest* out, const Fp* matrix, size_t rows, size_t cols) {
#pragma omp parallel for schedule(static)
  for (size_t r = 0; r < rows; ++r) {
    Sponge sp;
    for (size_t c = 0; c < cols; ++c) sp.absorb(matrix[c * rows + r]);
    out[r] = sp.finish();
  }
}
static void hash_fold(Digest* io, size_t input_size, size_t output_size) {
#pragma omp parallel for schedule(static)
  for (size_t i = 0; i < output_size; ++i) io[output_size + i] = hash_pair(io[input_size + 2 * i], io[input_size + 2 * i + 1]);
}
static void batch_evaluate_any(const Fp* coeffs, size_t poly_count, size_t n, const uint32_t* which, const Fp4* xs, Fp4* out, size_t eval_count) {
  (void)poly_count;
#pragma omp parallel for schedule(dynamic, 1)
  for (size_t j = 0; j < eval_count; ++j) {
    const Fp* p = coeffs + (size_t)which[j] * n;
    Fp4 tot, cur = Fp4::one(), x = xs[j];
    for (size_t i = 0; i < n; ++i) { tot += cur * p[i]; cur *= x; }
    out[j] = tot;
  }
}
static void mix_poly_coeffs(Fp4* out, Fp4 mix_start, Fp4 mix, const Fp* in, const uint32_t* combos, size_t input_size, size_t count) {
#pragma omp parallel for schedule(static)
  for (size_t idx = 0; idx < count; ++idx) {
    Fp4 cur = mix_start;
    for (size_t i = 0; i < input_size; ++i) {
      out[(size_t)combos[i] * count + idx] += cur * in[i * count + idx];
      cur *= mix;
    }
  }
}
static void eltwise_sum_extelem(Fp* out, const Fp4* in, size_t count, size_t to_add) {
#pragma omp parallel for schedule(static)
  for (size_t idx = 0; idx < count; ++idx) {
    Fp4 tot;
    for (size_t k = 0; k < to_add; ++k) tot += in[k * count + idx];
    for (int j = 0; j < 4; ++j) out[j * count + idx] = tot.c[j];
  }
}
static void fri_fold(Fp* out, const Fp* in, Fp4 mix, size_t m) {   // m = out.len/4, in.len = 64*m
#pragma omp parallel for schedule(static)
  for (size_t idx = 0; idx < m; ++idx) {
    Fp4 tot, cur = Fp4::one();
    for (size_t i = 0; i < FRI_FOLD; ++i) {
      size_t r = bit_rev((uint32_t)i, FRI_FOLD_PO2) * m + idx;
      Fp4 v(in[0 * 16 * m + r], in[1 * 16 * m + r], in[2 * 16 * m + r], in[3 * 16 * m + r]);
      tot += cur * v;
      cur *= mix;
    }
    for (int j = 0; j < 4; ++j) out[j * m + idx] = tot.c[j];
  }
}
static void eltwise_add_elem(Fp* o, const Fp* a, const Fp* b, size_t n) { for (size_t i = 0; i < n; ++i) o[i] = a[i] + b[i]; }
static void eltwise_copy_elem(Fp* o, const Fp* a, size_t n) { for (size_t i = 0; i < n; ++i) o[i] = a[i]; }
static void eltwise_zeroize_elem(Fp* x, size_t n) { for (size_t i = 0; i < n; ++i) if (x[i].v == INVALID) x[i].v = 0; }
static void gather_sample(Fp* dst, const Fp* src, size_t idx, size_t size, size_t stride) { for (size_t i = 0; i < size; ++i) dst[i] = src[idx + i * stride]; }
static void prefix_products(Fp4* io, size_t n) { for (size_t i = 1; i < n; ++i) io[i] *= io[i - 1]; }
// Hal::scatter(into, index, offsets, values) (witgen helper, App. C): row r owns entries index[r] .. index[r+1]
static void scatter(Fp* into, const uint32_t* index, size_t n_rows, const uint32_t* offsets, const Fp* values) {
  for (size_t r = 0; r < n_rows; ++r)
    for (uint32_t k = index[r]; k < index[r + 1]; ++k) into[offsets[k]] = values[k];
}

// ---- core::poly -----------------------------------------------------------------------------------
static Fp4 poly_eval(const Fp4* coeffs, size_t n, Fp4 x) {
  Fp4 tot, cur = Fp4::one();
  for (size_t i = 0; i < n; ++i) { tot += coeffs[i] * cur; cur *= x; }
  return tot;
}
// p(x) /= (x - z); returns the remainder (App. C.13).
static Fp4 poly_divide(Fp4* p, size_t n, Fp4 z) {
  Fp4 cur;
  for (size_t i = n; i-- > 0;) { Fp4 next = z * cur + p[i]; p[i] = cur; cur = next; }
  return cur;
}
static void poly_interpolate(Fp4* out, const Fp4* x, const Fp4* fx, size_t size) {
  if (size == 1) { out[0] = fx[0]; return; }
  if (size == 2) {
    out[1] = (fx[0] - fx[1]) * f4_inv(x[0] - x[1]);
    out[0] = fx[0] - out[1] * x[0];
    return;
  }
  std::vector<Fp4> ft(size + 1);     // ft = prod (x - x_i)
  ft[0] = Fp4::one();
  for (size_t i = 0; i < size; ++i) {
    for (size_t j = i + 1; j >= 1; --j) ft[j] = ft[j - 1] - x[i] * ft[j];
    ft[0] = Fp4::zero() - x[i] * ft[0];
  }
  for (size_t i = 0; i < size; ++i) out[i] = Fp4::zero();
  for (size_t i = 0; i < size; ++i) {
    std::vector<Fp4> fr(ft);
    poly_divide(fr.data(), size + 1, x[i]);
    Fp4 fr_xi = poly_eval(fr.data(), size, x[i]);
    Fp4 mul = fx[i] * f4_inv(fr_xi);
    for (size_t j = 0; j < size; ++j) out[j] += mul * fr[j];
  }
}

}  // namespace orc
