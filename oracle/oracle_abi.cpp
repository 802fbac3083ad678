// ORACLE (test infrastructure, NOT product code) -- parity unpinned vs. the Rust crates.
//
// C entry points over the oracle (ctypes-loaded by oracle/oracle.py from tests/, smoke() and the
// cpu_baseline / --impl reference legs of bench.py only).  All Fp are Montgomery u32 words, Fp4 = 4
// consecutive words, Digest = 8 words, matrices column-major -- the same conventions as
// include/zkb200.h so that test code can hand identical buffers to both sides.
#include "prover.hpp"
#include <cstdlib>
#include <cstring>
#include <string>
#include <omp.h>

using namespace orc;

static Fp4 ld4(const uint32_t* w) { return Fp4(Fp::raw(w[0]), Fp::raw(w[1]), Fp::raw(w[2]), Fp::raw(w[3])); }
static char* dup_err(const std::exception& e) { return strdup(e.what()); }
#define ORC_TRY try {
#define ORC_END } catch (const std::exception& e) { return dup_err(e); } return nullptr;

extern "C" {

int orc_num_threads() { return omp_get_max_threads(); }
void orc_set_num_threads(int n) { omp_set_num_threads(n); }

// -- field / constants -----------------------------------------------------------------------------
uint32_t orc_fp_encode(uint32_t x) { return Fp::from(x).v; }
uint32_t orc_fp_decode(uint32_t w) { return Fp::raw(w).as_u32(); }
uint32_t orc_fp_mul(uint32_t a, uint32_t b) { return mont_mul(a, b); }
uint32_t orc_fp_add(uint32_t a, uint32_t b) { return f_add(a, b); }
uint32_t orc_fp_sub(uint32_t a, uint32_t b) { return f_sub(a, b); }
uint32_t orc_fp_inv(uint32_t a) { return f_inv(Fp::raw(a)).v; }
void orc_fp4_mul(uint32_t* out, const uint32_t* a, const uint32_t* b) { Fp4 r = ld4(a) * ld4(b); memcpy(out, &r, 16); }
void orc_fp4_inv(uint32_t* out, const uint32_t* a) { Fp4 r = f4_inv(ld4(a)); memcpy(out, &r, 16); }
void orc_rou_tables(uint32_t* fwd_canonical, uint32_t* rev_canonical) {
  for (int i = 0; i <= MAX_ROU_PO2; ++i) { fwd_canonical[i] = rou().fwd[i].as_u32(); rev_canonical[i] = rou().rev[i].as_u32(); }
}
void orc_poseidon2_round_constants(uint32_t* canonical213) { memcpy(canonical213, p2c().canonical, sizeof(p2c().canonical)); }
void orc_poseidon2_mix(uint32_t* cells24) { poseidon2_mix((Fp*)cells24); }
void orc_hash_elem_slice(uint32_t* digest, const uint32_t* elems, size_t n) { Digest d = hash_elem_slice((const Fp*)elems, n); memcpy(digest, d.w, 32); }
void orc_hash_pair(uint32_t* digest, const uint32_t* a, const uint32_t* b) { Digest d = hash_pair(*(const Digest*)a, *(const Digest*)b); memcpy(digest, d.w, 32); }

// -- Poseidon2Rng ----------------------------------------------------------------------------------
void* orc_rng_new() { return new Poseidon2Rng(); }
void orc_rng_free(void* r) { delete (Poseidon2Rng*)r; }
void orc_rng_mix(void* r, const uint32_t* digest) { ((Poseidon2Rng*)r)->mix(*(const Digest*)digest); }
uint32_t orc_rng_random_elem(void* r) { return ((Poseidon2Rng*)r)->random_elem().v; }
uint32_t orc_rng_random_bits(void* r, int bits) { return ((Poseidon2Rng*)r)->random_bits(bits); }

// -- Hal operators -----------------------------------------------------------------------------------
void orc_batch_interpolate_ntt(uint32_t* io, size_t count, int po2) { batch_interpolate_ntt((Fp*)io, count, (size_t)1 << po2); }
void orc_zk_shift(uint32_t* io, size_t count, int po2) { zk_shift((Fp*)io, count, (size_t)1 << po2); }
void orc_batch_expand(uint32_t* out, const uint32_t* in, size_t count, int in_po2, int expand_bits) { batch_expand((Fp*)out, (const Fp*)in, count, (size_t)1 << in_po2, expand_bits); }
void orc_batch_evaluate_ntt(uint32_t* io, size_t count, int po2, int expand_bits) { batch_evaluate_ntt((Fp*)io, count, (size_t)1 << po2, expand_bits); }
void orc_batch_expand_into_evaluate_ntt(uint32_t* out, const uint32_t* in, size_t count, int in_po2, int expand_bits) {
  batch_expand_into_evaluate_ntt((Fp*)out, (const Fp*)in, count, (size_t)1 << in_po2, expand_bits);
}
void orc_batch_bit_reverse(uint32_t* io, size_t count, int po2) { batch_bit_reverse((Fp*)io, count, (size_t)1 << po2); }
void orc_hash_rows(uint32_t* out, const uint32_t* matrix, size_t rows, size_t cols) { hash_rows((Digest*)out, (const Fp*)matrix, rows, cols); }
void orc_hash_fold(uint32_t* io, size_t input_size, size_t output_size) { hash_fold((Digest*)io, input_size, output_size); }
void orc_merkle_build(uint32_t* nodes, size_t rows) {
  for (size_t l = (size_t)log2_exact(rows); l-- > 0;) hash_fold((Digest*)nodes, (size_t)2 << l, (size_t)1 << l);
}
void orc_batch_evaluate_any(const uint32_t* coeffs, size_t poly_count, int po2, const uint32_t* which, const uint32_t* xs, uint32_t* out, size_t eval_count) {
  batch_evaluate_any((const Fp*)coeffs, poly_count, (size_t)1 << po2, which, (const Fp4*)xs, (Fp4*)out, eval_count);
}
void orc_mix_poly_coeffs(uint32_t* out, const uint32_t* mix_start, const uint32_t* mix, const uint32_t* in, const uint32_t* combos, size_t input_size, size_t count) {
  mix_poly_coeffs((Fp4*)out, ld4(mix_start), ld4(mix), (const Fp*)in, combos, input_size, count);
}
void orc_eltwise_sum_extelem(uint32_t* out, const uint32_t* in, size_t count, size_t to_add) { eltwise_sum_extelem((Fp*)out, (const Fp4*)in, count, to_add); }
void orc_fri_fold(uint32_t* out, const uint32_t* in, const uint32_t* mix, size_t out_count) { fri_fold((Fp*)out, (const Fp*)in, ld4(mix), out_count); }
void orc_eltwise_add_elem(uint32_t* o, const uint32_t* a, const uint32_t* b, size_t n) { eltwise_add_elem((Fp*)o, (const Fp*)a, (const Fp*)b, n); }
void orc_eltwise_copy_elem(uint32_t* o, const uint32_t* a, size_t n) { eltwise_copy_elem((Fp*)o, (const Fp*)a, n); }
void orc_eltwise_zeroize_elem(uint32_t* x, size_t n) { eltwise_zeroize_elem((Fp*)x, n); }
void orc_gather_sample(uint32_t* dst, const uint32_t* src, size_t idx, size_t size, size_t stride) { gather_sample((Fp*)dst, (const Fp*)src, idx, size, stride); }
void orc_prefix_products(uint32_t* io, size_t n) { prefix_products((Fp4*)io, n); }
void orc_scatter(uint32_t* into, const uint32_t* index, size_t n_rows, const uint32_t* offsets, const uint32_t* values) { scatter((Fp*)into, index, n_rows, offsets, (const Fp*)values); }
// returns the remainder in rem[4]
void orc_poly_divide(uint32_t* p, size_t n, const uint32_t* z, uint32_t* rem) { Fp4 r = poly_divide((Fp4*)p, n, ld4(z)); memcpy(rem, &r, 16); }
void orc_poly_interpolate(uint32_t* out, const uint32_t* x, const uint32_t* fx, size_t size) { poly_interpolate((Fp4*)out, (const Fp4*)x, (const Fp4*)fx, size); }

// -- circuit + prover -------------------------------------------------------------------------------
struct OrcProver {
  Circuit circuit;
  std::unique_ptr<Prover> prover;
  std::vector<Fp> io, mix;
};

const char* orc_eval_check(uint32_t* check, const uint32_t* blob, size_t blob_words, const uint32_t* accum, const uint32_t* code, const uint32_t* data,
                           const uint32_t* mix_g, const uint32_t* out_g, const uint32_t* poly_mix, int po2) {
  ORC_TRY
  Circuit c = Circuit::parse(blob, blob_words);
  const Fp* groups[3] = {(const Fp*)accum, (const Fp*)code, (const Fp*)data};
  eval_check((Fp*)check, c, groups, (const Fp*)mix_g, (const Fp*)out_g, ld4(poly_mix), po2);
  ORC_END
}

const char* orc_accumulate(const uint32_t* blob, size_t blob_words, uint32_t* accum, const uint32_t* code, const uint32_t* data,
                           const uint32_t* mix_g, const uint32_t* out_g, int po2) {
  ORC_TRY
  Circuit c = Circuit::parse(blob, blob_words);
  accumulate(c, (Fp*)accum, (const Fp*)code, (const Fp*)data, (const Fp*)mix_g, (const Fp*)out_g, po2);
  ORC_END
}

const char* orc_prover_new(const uint32_t* blob, size_t blob_words, void** out) {
  ORC_TRY
  std::unique_ptr<OrcProver> p(new OrcProver);
  p->circuit = Circuit::parse(blob, blob_words);
  p->prover.reset(new Prover(p->circuit));
  *out = p.release();
  ORC_END
}
void orc_prover_free(void* h) { delete (OrcProver*)h; }

// code/data: group_size x 2^po2 column-major evaluations; io: out_size elems; mix_out: mix_size elems.
const char* orc_segment_begin(void* h, int po2, const uint32_t* io, const uint32_t* code, const uint32_t* data, uint32_t* mix_out) {
  ORC_TRY
  OrcProver* p = (OrcProver*)h;
  const Circuit& c = p->circuit;
  size_t n = (size_t)1 << po2;
  p->io.assign((const Fp*)io, (const Fp*)io + c.out_size);
  p->mix.resize(c.mix_size);
  std::vector<Fp> cv((const Fp*)code, (const Fp*)code + (size_t)c.group_size[1] * n), dv((const Fp*)data, (const Fp*)data + (size_t)c.group_size[2] * n);
  segment_begin(*p->prover, po2, p->io.data(), std::move(cv), std::move(dv), p->mix.data());
  memcpy(mix_out, p->mix.data(), 4 * c.mix_size);
  ORC_END
}
const char* orc_segment_finish(void* h, const uint32_t* accum) {
  ORC_TRY
  OrcProver* p = (OrcProver*)h;
  size_t n = p->prover->n;
  std::vector<Fp> av((const Fp*)accum, (const Fp*)accum + (size_t)p->circuit.group_size[0] * n);
  segment_finish(*p->prover, p->io.data(), p->mix.data(), std::move(av));
  ORC_END
}
size_t orc_seal_words(void* h) { return ((OrcProver*)h)->prover->iop.proof.size(); }
void orc_seal_copy(void* h, uint32_t* out) { auto& s = ((OrcProver*)h)->prover->iop.proof; memcpy(out, s.data(), 4 * s.size()); }
size_t orc_root_count(void* h) { return ((OrcProver*)h)->prover->roots.size(); }
void orc_roots_copy(void* h, uint32_t* out) { auto& r = ((OrcProver*)h)->prover->roots; memcpy(out, r.data(), 32 * r.size()); }

}  // extern "C"
