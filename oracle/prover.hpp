// ORACLE (test infrastructure, NOT product code) -- parity unpinned vs. the Rust crates.
//
// CPU restatement of the risc0-zkp 1.2.5 prover orchestration (`prove/{write_iop, poly_group,
// merkle, fri, prover}.rs`, `merkle.rs`, `taps.rs`) and of the circuit-side `eval_check` /
// `prove_segment` driver (risc0-circuit-rv32im/src/prove/{mod.rs, hal/cpu.rs}); all un-vendored
// (/root/reference/Cargo.lock:5057-5085, 4961-4964).  Order of operations per SURVEY.md App. D --
// the order is what makes Merkle roots, FRI commitments and seals comparable.
#pragma once
#include "hal.hpp"
#include "circuit.hpp"
#include <memory>

namespace orc {

// ---- WriteIOP (D.0) --------------------------------------------------------------------------------
struct WriteIOP {
  std::vector<uint32_t> proof;
  Poseidon2Rng rng;
  void write_fp(const Fp* s, size_t n) { for (size_t i = 0; i < n; ++i) proof.push_back(s[i].v); }
  void write_fp4(const Fp4* s, size_t n) { for (size_t i = 0; i < n; ++i) for (int j = 0; j < 4; ++j) proof.push_back(s[i].c[j].v); }
  void write_digests(const Digest* d, size_t n) { for (size_t i = 0; i < n; ++i) for (int j = 0; j < 8; ++j) proof.push_back(d[i].w[j]); }
  void commit(const Digest& d) { rng.mix(d); }
};

// ---- MerkleTreeParams / MerkleTreeProver (D.2, D.5) ----------------------------------------------
struct MerkleParams {
  size_t row_size, col_size, queries, layers, top_layer, top_size;
  MerkleParams(size_t rows, size_t cols, size_t q) : row_size(rows), col_size(cols), queries(q) {
    layers = (size_t)log2_exact(rows);
    top_layer = 0;
    for (size_t i = 1; i < layers; ++i) { if (((size_t)1 << i) > queries) break; top_layer = i; }
    top_size = (size_t)1 << top_layer;
  }
};
struct MerkleTreeProver {
  MerkleParams params;
  const Fp* matrix;                 // cols x rows, column-major (not owned)
  std::vector<Digest> nodes;        // heap-indexed, 2*rows
  MerkleTreeProver(const Fp* m, size_t rows, size_t cols, size_t queries) : params(rows, cols, queries), matrix(m), nodes(2 * rows) {
    hash_rows(nodes.data() + rows, matrix, rows, cols);
    for (size_t l = params.layers; l-- > 0;) hash_fold(nodes.data(), (size_t)2 << l, (size_t)1 << l);
  }
  const Digest& root() const { return nodes[1]; }
  void commit(WriteIOP& iop) const {
    iop.write_digests(nodes.data() + params.top_size, params.top_size);
    iop.commit(root());
  }
  void prove(WriteIOP& iop, size_t idx) const {
    std::vector<Fp> row(params.col_size);
    gather_sample(row.data(), matrix, idx, params.col_size, params.row_size);
    iop.write_fp(row.data(), row.size());
    for (size_t i = idx + params.row_size; i >= 2 * params.top_size; i >>= 1) iop.write_digests(&nodes[i ^ 1], 1);
  }
};

// ---- PolyGroup -------------------------------------------------------------------------------------
struct PolyGroup {
  size_t count, n;
  std::vector<Fp> coeffs;       // count x n, natural order after construction
  std::vector<Fp> evaluated;    // count x 4n
  std::unique_ptr<MerkleTreeProver> merkle;
  // takes bit-reversed coefficients
  PolyGroup(std::vector<Fp>&& c, size_t count_, size_t n_) : count(count_), n(n_), coeffs(std::move(c)), evaluated(count_ * n_ * INV_RATE) {
    batch_expand_into_evaluate_ntt(evaluated.data(), coeffs.data(), count, n, 2);
    batch_bit_reverse(coeffs.data(), count, n);
    merkle.reset(new MerkleTreeProver(evaluated.data(), n * INV_RATE, count, QUERIES));
  }
};

// ---- eval_check (App. C.12) ---------------------------------------------------------------------
static void eval_check(Fp* check, const Circuit& c, const Fp* const groups[3], const Fp* mix_g, const Fp* out_g, Fp4 poly_mix, int po2) {
  size_t n = (size_t)1 << po2, domain = n * INV_RATE;
  Fp rou_d = rou().fwd[po2 + 2];
  Fp three = Fp::from(3), one = Fp::from(1);
#pragma omp parallel
  {
    std::vector<Fp> fp_vars; std::vector<MixState> mix_vars;
#pragma omp for schedule(static)
    for (size_t cyc = 0; cyc < domain; ++cyc) {
      auto get = [&](uint32_t tap_idx) {
        const Tap& t = c.taps[tap_idx];
        size_t row = (cyc + domain - (size_t)INV_RATE * t.back) & (domain - 1);
        return groups[t.group][(size_t)t.column * domain + row];
      };
      Fp4 tot = poly_ext<Fp>(c, poly_mix, get, mix_g, out_g, fp_vars, mix_vars);
      Fp x = f_pow(rou_d, cyc);
      Fp y = f_pow(three * x, n);
      Fp4 ret = tot * f_inv(y - one);
      for (int j = 0; j < 4; ++j) check[j * domain + cyc] = ret.c[j];
    }
  }
}

// ---- fri_prove (D.4) -----------------------------------------------------------------------------
struct FriRound {
  std::vector<Fp> evaluated;
  std::unique_ptr<MerkleTreeProver> merkle;
  size_t domain;
};
template <typename Inner>
static void fri_prove(WriteIOP& iop, std::vector<Fp> coeffs /* 4 planar x len, bit-reversed */, Inner inner, std::vector<Digest>* roots) {
  size_t len = coeffs.size() / EXT_SIZE;
  size_t orig_domain = len * INV_RATE;
  std::vector<std::unique_ptr<FriRound>> rounds;
  while (len > FRI_MIN_DEGREE) {
    std::unique_ptr<FriRound> r(new FriRound);
    r->domain = len * INV_RATE;
    r->evaluated.resize(EXT_SIZE * r->domain);
    batch_expand_into_evaluate_ntt(r->evaluated.data(), coeffs.data(), EXT_SIZE, len, 2);
    r->merkle.reset(new MerkleTreeProver(r->evaluated.data(), r->domain / FRI_FOLD, FRI_FOLD * EXT_SIZE, QUERIES));
    r->merkle->commit(iop);
    if (roots) roots->push_back(r->merkle->root());
    Fp4 fold_mix = iop.rng.random_ext_elem();
    std::vector<Fp> out(coeffs.size() / FRI_FOLD);
    fri_fold(out.data(), coeffs.data(), fold_mix, out.size() / EXT_SIZE);
    coeffs.swap(out);
    len /= FRI_FOLD;
    rounds.push_back(std::move(r));
  }
  std::vector<Fp> fin(coeffs);
  batch_bit_reverse(fin.data(), EXT_SIZE, len);
  iop.write_fp(fin.data(), fin.size());
  iop.commit(hash_elem_slice(fin.data(), fin.size()));
  int bits = log2_exact(orig_domain);
  for (size_t q = 0; q < QUERIES; ++q) {
    size_t pos = iop.rng.random_bits(bits);
    inner(iop, pos);
    for (auto& r : rounds) {
      size_t group = pos % (r->domain / FRI_FOLD);
      r->merkle->prove(iop, group);
      pos = group;
    }
  }
}

// ---- Prover (commit_group / finalize) + segment driver (D.2, D.3) -----------------------------------
struct Prover {
  const Circuit& circuit;
  WriteIOP iop;
  int po2 = 0; size_t n = 0;
  std::unique_ptr<PolyGroup> groups[3];
  std::unique_ptr<PolyGroup> check_group;
  std::vector<Digest> roots;       // commit order: group commits, check, FRI rounds (for root-level parity checks)
  explicit Prover(const Circuit& c) : circuit(c) {}
  void set_po2(int p) { po2 = p; n = (size_t)1 << p; }

  // trace: group_size[g] x n column-major evaluations (consumed)
  void commit_group(uint32_t g, std::vector<Fp>&& trace) {
    size_t cols = circuit.group_size[g];
    batch_interpolate_ntt(trace.data(), cols, n);
    zk_shift(trace.data(), cols, n);
    groups[g].reset(new PolyGroup(std::move(trace), cols, n));
    groups[g]->merkle->commit(iop);
    roots.push_back(groups[g]->merkle->root());
  }

  void finalize(const Fp* mix_g, const Fp* out_g) {
    const Circuit& c = circuit;
    size_t domain = n * INV_RATE;
    // 1. check polynomial
    Fp4 poly_mix = iop.rng.random_ext_elem();
    std::vector<Fp> check(EXT_SIZE * domain);
    const Fp* ev[3] = {groups[0]->evaluated.data(), groups[1]->evaluated.data(), groups[2]->evaluated.data()};
    eval_check(check.data(), c, ev, mix_g, out_g, poly_mix, po2);
    batch_interpolate_ntt(check.data(), EXT_SIZE, domain);
    check_group.reset(new PolyGroup(std::move(check), CHECK_SIZE, n));
    check_group->merkle->commit(iop);
    roots.push_back(check_group->merkle->root());
    // 2. DEEP evaluations
    Fp4 z = iop.rng.random_ext_elem();
    Fp back_one = rou().rev[po2];
    size_t tap_size = c.tap_size();
    std::vector<Fp4> all_xs(tap_size), eval_u(tap_size);
    for (uint32_t g = 0; g < 3; ++g) {
      size_t b = c.group_tap_begin(g), e = c.group_tap_end(g);
      std::vector<uint32_t> which;
      for (size_t t = b; t < e; ++t) { which.push_back(c.taps[t].column); all_xs[t] = z * f_pow(back_one, c.taps[t].back); }
      batch_evaluate_any(groups[g]->coeffs.data(), c.group_size[g], n, which.data(), all_xs.data() + b, eval_u.data() + b, e - b);
    }
    // 3. coeff_u
    std::vector<Fp4> coeff_u(tap_size + CHECK_SIZE);
    for (const Register& r : c.regs) poly_interpolate(&coeff_u[r.tap_pos], &all_xs[r.tap_pos], &eval_u[r.tap_pos], r.size);
    Fp4 z_pow = f4_pow(z, EXT_SIZE);
    {
      std::vector<uint32_t> which(CHECK_SIZE); std::vector<Fp4> xs(CHECK_SIZE, z_pow);
      for (size_t i = 0; i < CHECK_SIZE; ++i) which[i] = (uint32_t)i;
      batch_evaluate_any(check_group->coeffs.data(), CHECK_SIZE, n, which.data(), xs.data(), &coeff_u[tap_size], CHECK_SIZE);
    }
    // 4.
    iop.write_fp4(coeff_u.data(), coeff_u.size());
    iop.commit(hash_ext_elem_slice(coeff_u.data(), coeff_u.size()));
    Fp4 mix = iop.rng.random_ext_elem();
    // 5. combos
    size_t combos_size = c.combos_size();
    std::vector<Fp4> combos((combos_size + 1) * n);
    Fp4 cur = Fp4::one();
    {
      size_t reg_i = 0;
      for (uint32_t g = 0; g < 3; ++g) {
        std::vector<uint32_t> ids;
        while (reg_i < c.regs.size() && c.regs[reg_i].group == g) ids.push_back(c.regs[reg_i++].combo_id);
        if (ids.size() != c.group_size[g]) throw std::runtime_error("every column of a group needs at least one tap");
        mix_poly_coeffs(combos.data(), cur, mix, groups[g]->coeffs.data(), ids.data(), ids.size(), n);
        cur *= f4_pow(mix, ids.size());
      }
      std::vector<uint32_t> ids(CHECK_SIZE, (uint32_t)combos_size);
      mix_poly_coeffs(combos.data(), cur, mix, check_group->coeffs.data(), ids.data(), CHECK_SIZE, n);
    }
    // 6. subtract the interpolants and divide out the evaluation points
    cur = Fp4::one();
    for (const Register& r : c.regs) {
      for (size_t i = 0; i < r.size; ++i) combos[n * r.combo_id + i] -= cur * coeff_u[r.tap_pos + i];
      cur *= mix;
    }
    for (size_t i = 0; i < CHECK_SIZE; ++i) { combos[n * combos_size] -= cur * coeff_u[tap_size + i]; cur *= mix; }
    for (size_t ci = 0; ci < combos_size; ++ci)
      for (uint32_t back : c.combos[ci]) {
        Fp4 rem = poly_divide(&combos[ci * n], n, z * f_pow(back_one, back));
        if (rem != Fp4::zero()) throw std::runtime_error("combo division left a remainder");
      }
    if (poly_divide(&combos[combos_size * n], n, z_pow) != Fp4::zero()) throw std::runtime_error("check combo division left a remainder");
    // 7. FRI
    std::vector<Fp> fin(EXT_SIZE * n);
    eltwise_sum_extelem(fin.data(), combos.data(), n, combos_size + 1);
    batch_bit_reverse(fin.data(), EXT_SIZE, n);
    fri_prove(iop, std::move(fin), [&](WriteIOP& io, size_t idx) {
      for (uint32_t g = 0; g < 3; ++g) groups[g]->merkle->prove(io, idx);
      check_group->merkle->prove(io, idx);
    }, &roots);
  }
};

static const char PROOF_SYSTEM_INFO[17] = "RISC0_STARK:v1__";

static Digest hash_protocol_info(const uint8_t* info) {
  Fp e[16]; for (int i = 0; i < 16; ++i) e[i] = Fp::from(info[i]);
  return hash_elem_slice(e, 16);
}

// Segment driver, first half: commits the header, `code` and `data`, then draws the `mix` globals.
static void segment_begin(Prover& p, int po2, const Fp* io, std::vector<Fp>&& code, std::vector<Fp>&& data, Fp* mix_out) {
  const Circuit& c = p.circuit;
  p.iop.commit(hash_protocol_info((const uint8_t*)PROOF_SYSTEM_INFO));
  p.iop.commit(hash_protocol_info(c.info));
  std::vector<Fp> hdr(io, io + c.out_size);
  hdr.push_back(Fp::raw((uint32_t)po2));      // Elem::from_u32_slice(&[po2]): a raw reinterpretation, NOT the Montgomery encoding of po2
  p.iop.commit(hash_elem_slice(hdr.data(), hdr.size()));
  p.iop.write_fp(hdr.data(), hdr.size());
  p.set_po2(po2);
  p.commit_group(1, std::move(code));
  p.commit_group(2, std::move(data));
  for (uint32_t i = 0; i < c.mix_size; ++i) mix_out[i] = p.iop.rng.random_elem();
}
// Second half: commits `accum` (produced by the caller's accumulate step from `mix`) and finalizes.
static void segment_finish(Prover& p, const Fp* io, const Fp* mix, std::vector<Fp>&& accum) {
  p.commit_group(0, std::move(accum));
  p.finalize(mix, io);
}

}  // namespace orc
