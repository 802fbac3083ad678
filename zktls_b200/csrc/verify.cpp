// CPU verifier for segment seals: zkb_verify_segment.
//
// Restates risc0-zkp 1.2.5 `verify::{mod.rs (Verifier::verify), fri.rs (fri_verify), merkle.rs (MerkleTreeVerifier),
// read_iop.rs (ReadIOP)}` (un-vendored; pinned at /root/reference/Cargo.lock:5057-5085; the reference reaches it through
// `receipt.verify` inside the call at /root/reference/crates/guest-prover-r0/src/prover.rs:90).  SURVEY.md 8(f)-1: this turns
// "the seal equals the oracle's seal" into "the seal VERIFIES".  Host-only, like the reference's verifier; it shares no
// code with the prover except the field, the Poseidon2 permutation and the circuit parser.
//
// Checks, in transcript order: header commitments, the four group Merkle top layers, the constraint
// relation at the DEEP point z (poly_ext over the tap evaluations recovered from coeff_u against the check polynomial),
// then 50 FRI queries: every row's Merkle path, the DEEP quotient combination at the query point, every fold, and the
// final low-degree polynomial.
#include "common.cuh"
#include "circuit.hpp"
#include "transcript.hpp"

namespace zkb {
namespace {

struct VerifyError : std::runtime_error { using std::runtime_error::runtime_error; };
#define VFY(cond, msg) do { if (!(cond)) throw VerifyError(std::string("zkb200: invalid proof: ") + (msg)); } while (0)

struct ReadIOP {
  const uint32_t* p; size_t left;
  HostRng rng;
  ReadIOP(const uint32_t* seal, size_t n) : p(seal), left(n) {}
  const uint32_t* read(size_t n) {
    VFY(n <= left, "seal too short");
    const uint32_t* r = p; p += n; left -= n; return r;
  }
  // field elements must be canonical Montgomery words (ReadIOP::read_field_elem_slice validates with `is_valid`)
  const uint32_t* read_fp(size_t n) {
    const uint32_t* r = read(n);
    for (size_t i = 0; i < n; ++i) VFY(r[i] < P, "non-canonical field element in seal");
    return r;
  }
  void commit(const Digest& d) { rng.mix(d); }
};

static Fp rou_fwd(int po2) { return pow(Fp::from(137), (uint64_t)1 << (MAX_ROU_PO2 - po2)); }
static Fp rou_rev(int po2) { return inv(rou_fwd(po2)); }
static Fp4 poly_eval(const Fp4* coeffs, size_t n, const Fp4& x) {
  Fp4 tot, cur = Fp4::one();
  for (size_t i = 0; i < n; ++i) { tot += coeffs[i] * cur; cur *= x; }
  return tot;
}

// merkle.rs MerkleTreeParams + verify/merkle.rs MerkleTreeVerifier
struct MerkleVerifier {
  size_t rows, cols, layers, top_layer, top_size;
  std::vector<Digest> top;     // heap-indexed [1, 2*top_size)
  MerkleVerifier(ReadIOP& iop, size_t rows_, size_t cols_) : rows(rows_), cols(cols_) {
    layers = 0; while (((size_t)1 << layers) < rows) ++layers;
    VFY(((size_t)1 << layers) == rows, "merkle rows not a power of two");
    top_layer = 0;
    for (size_t i = 1; i < layers; ++i) { if (((size_t)1 << i) > QUERIES) break; top_layer = i; }
    top_size = (size_t)1 << top_layer;
    top.resize(2 * top_size);
    const uint32_t* w = iop.read(top_size * 8);
    for (size_t i = 0; i < top_size; ++i) memcpy(top[top_size + i].w, w + 8 * i, 32);
    for (size_t i = top_size; i-- > 1;) top[i] = hash_pair_host(top[2 * i], top[2 * i + 1]);
    iop.commit(root());
  }
  const Digest& root() const { return top[1]; }
  // reads the row and its path, checks it against the committed top layer, returns the row
  const uint32_t* verify(ReadIOP& iop, size_t idx) const {
    VFY(idx < rows, "merkle index out of range");
    const uint32_t* row = iop.read_fp(cols);
    Digest cur = hash_words(row, cols);
    idx += rows;
    while (idx >= 2 * top_size) {
      Digest other; memcpy(other.w, iop.read(8), 32);
      cur = (idx & 1) ? hash_pair_host(other, cur) : hash_pair_host(cur, other);
      idx >>= 1;
    }
    VFY(memcmp(top[idx].w, cur.w, 32) == 0, "merkle path does not match the committed top layer");
    return row;
  }
};

// in-place size-16 inverse NTT on Fp4 (DIF, natural -> bit-reversed, scaled 1/16), then bit reversal: the coefficients of
// the degree-15 interpolant (core/ntt.rs interpolate_ntt + bit_reverse, as used by verify/fri.rs fold_eval)
static void interpolate16(Fp4* io) {
  for (int bits = 4; bits >= 1; --bits) {
    const size_t half = (size_t)1 << (bits - 1);
    const Fp step = rou_rev(bits);
    for (size_t blk = 0; blk < 16; blk += 2 * half) {
      Fp cur = Fp::one();
      for (size_t i = 0; i < half; ++i) {
        Fp4 a = io[blk + i], b = io[blk + i + half];
        io[blk + i] = a + b;
        io[blk + i + half] = (a - b) * cur;
        cur *= step;
      }
    }
  }
  const Fp norm = inv(Fp::from(16));
  Fp4 tmp[16];
  for (uint32_t i = 0; i < 16; ++i) tmp[bit_rev32(i, 4)] = io[i] * norm;
  for (int i = 0; i < 16; ++i) io[i] = tmp[i];
}

static const char PROOF_SYSTEM_INFO[17] = "RISC0_STARK:v1__";
static Digest hash_protocol_info(const uint8_t* info) {
  uint32_t e[16];
  for (int i = 0; i < 16; ++i) e[i] = Fp::from(info[i]).v;
  return hash_words(e, 16);
}

// check_code: risc0-zkp `Verifier::verify(.., check_code: impl Fn(u32, &Digest) -> Result<..>)` -- the code group's Merkle root IS the
// program being proven (its "control ID" for this po2).  Without this binding a prover may choose the code trace freely; in
// the SYN family an all-zero selector column makes every constraint vanish, so a seal for ANY claimed io would verify.
struct ControlCheck {
  const uint32_t* ids; size_t n;      // n entries of 9 words: po2, code root[8]
  uint32_t* out;                      // 9 words (po2, code root) handed back to the caller, or nullptr
  void operator()(uint32_t po2, const Digest& root) const {
    if (out) { out[0] = po2; memcpy(out + 1, root.w, 32); }
    if (n == 0) { VFY(out != nullptr, "no control ID to check the code root against"); return; }
    for (size_t i = 0; i < n; ++i)
      if (ids[9 * i] == po2 && memcmp(ids + 9 * i + 1, root.w, 32) == 0) return;
    VFY(false, "code root is not a control ID of this circuit for this po2 (check_code)");
  }
};

static void verify_segment(const CircuitDef& c, const uint32_t* seal, size_t seal_words, const ControlCheck& check_code) {
  ReadIOP iop(seal, seal_words);
  // header (App. D.2)
  iop.commit(hash_protocol_info((const uint8_t*)PROOF_SYSTEM_INFO));
  iop.commit(hash_protocol_info(c.info));
  const uint32_t* hdr = iop.read_fp(c.out_size + 1);
  iop.commit(hash_words(hdr, c.out_size + 1));
  const uint32_t* out_g = hdr;
  const uint32_t po2 = hdr[c.out_size];      // the raw word (`to_u32_words()[0]`), as the prover wrote it
  VFY(po2 >= 1 && po2 + 2 <= (uint32_t)MAX_PO2, "po2 out of range");
  const size_t n = (size_t)1 << po2, domain = n * INV_RATE;
  // group commitments: code, data, then the mix globals, then accum (commit order of prove_segment)
  MerkleVerifier code_m(iop, domain, c.group_size[GROUP_CODE]);
  check_code(po2, code_m.root());
  MerkleVerifier data_m(iop, domain, c.group_size[GROUP_DATA]);
  std::vector<uint32_t> mix_g(c.mix_size);
  for (uint32_t i = 0; i < c.mix_size; ++i) mix_g[i] = iop.rng.random_elem().v;
  MerkleVerifier accum_m(iop, domain, c.group_size[GROUP_ACCUM]);
  const Fp4 poly_mix = iop.rng.random_ext_elem();
  MerkleVerifier check_m(iop, domain, CHECK_SIZE);
  const Fp4 z = iop.rng.random_ext_elem();
  const Fp back_one = rou_rev((int)po2);
  // coeff_u: per-register interpolants of the tap evaluations, then the 16 check evaluations at z^4
  const size_t tap_size = c.tap_size();
  const uint32_t* cu_words = iop.read_fp((tap_size + CHECK_SIZE) * 4);
  std::vector<Fp4> coeff_u(tap_size + CHECK_SIZE);
  for (size_t i = 0; i < coeff_u.size(); ++i) coeff_u[i] = Fp4::load(cu_words + 4 * i);
  iop.commit(hash_words(cu_words, (tap_size + CHECK_SIZE) * 4));
  // tap evaluations at z * back_one^back from the interpolants
  std::vector<Fp4> eval_u(tap_size);
  for (const RegisterDef& r : c.regs)
    for (uint32_t i = 0; i < r.size; ++i) {
      Fp4 x = z * pow(back_one, c.taps[r.tap_pos + i].back);
      eval_u[r.tap_pos + i] = poly_eval(&coeff_u[r.tap_pos], r.size, x);
    }
  // the constraint relation at z
  const Fp4 result = poly_ext_host(c, poly_mix, eval_u.data(), mix_g.data(), out_g);
  Fp4 check;
  {
    static const int remap[4] = {0, 2, 1, 3};
    Fp4 zi = Fp4::one();
    for (int i = 0; i < 4; ++i) {
      for (int j = 0; j < 4; ++j) {
        Fp4 e; e.c[j] = Fp::one();
        check += coeff_u[tap_size + remap[i] + 4 * j] * zi * e;
      }
      zi *= z;
    }
    check *= pow(z * Fp::from(3), n) - Fp4::one();
  }
  VFY(check == result, "constraint polynomial does not match the check polynomial at the DEEP point");
  // DEEP combination
  const Fp4 mix = iop.rng.random_ext_elem();
  const size_t combos_size = c.combos_size();
  std::vector<size_t> combo_begin(combos_size + 2, 0);
  for (size_t i = 0; i < combos_size; ++i) combo_begin[i + 1] = combo_begin[i] + c.combos[i].size();
  combo_begin[combos_size + 1] = combo_begin[combos_size] + 1;
  std::vector<Fp4> combo_u(combo_begin[combos_size + 1]);
  std::vector<Fp4> mix_pows;     // one power per register, then one per check column
  {
    Fp4 cur = Fp4::one();
    for (const RegisterDef& r : c.regs) {
      for (uint32_t i = 0; i < r.size; ++i) combo_u[combo_begin[r.combo_id] + i] += cur * coeff_u[r.tap_pos + i];
      mix_pows.push_back(cur);
      cur *= mix;
    }
    for (size_t i = 0; i < CHECK_SIZE; ++i) {
      combo_u[combo_begin[combos_size]] += cur * coeff_u[tap_size + i];
      mix_pows.push_back(cur);
      cur *= mix;
    }
  }
  const Fp4 z_pow = pow(z, INV_RATE);
  // FRI (verify/fri.rs)
  struct Round { size_t dom; MerkleVerifier m; Fp4 mix; };
  std::vector<Round> rounds;
  size_t degree = n, dom = domain;
  while (degree > FRI_MIN_DEGREE) {
    MerkleVerifier m(iop, dom / FRI_FOLD, FRI_FOLD * EXT_SIZE);
    Fp4 fm = iop.rng.random_ext_elem();
    rounds.push_back(Round{dom, std::move(m), fm});
    dom /= FRI_FOLD; degree /= FRI_FOLD;
  }
  const uint32_t* fin = iop.read_fp(EXT_SIZE * degree);
  iop.commit(hash_words(fin, EXT_SIZE * degree));
  std::vector<Fp4> final_poly(degree);
  for (size_t i = 0; i < degree; ++i) final_poly[i] = Fp4::raw(fin[i], fin[degree + i], fin[2 * degree + i], fin[3 * degree + i]);
  int dom_po2 = 0; while (((size_t)1 << dom_po2) < domain) ++dom_po2;
  int fin_po2 = 0; while (((size_t)1 << fin_po2) < dom) ++fin_po2;
  const Fp gen_final = rou_fwd(fin_po2), gen_domain = rou_fwd(dom_po2);
  const MerkleVerifier* group_m[3] = {&accum_m, &code_m, &data_m};
  std::vector<Fp4> tot(combos_size + 1);
  for (size_t q = 0; q < QUERIES; ++q) {
    size_t pos = iop.rng.random_bits(dom_po2);
    // inner: rows of the four group trees at `pos`, combined into the DEEP quotient value at x = w^pos
    const Fp x = pow(gen_domain, pos);
    const uint32_t* rows[3];
    for (int g = 0; g < 3; ++g) rows[g] = group_m[g]->verify(iop, pos);
    const uint32_t* check_row = check_m.verify(iop, pos);
    for (auto& t : tot) t = Fp4::zero();
    for (size_t r = 0; r < c.regs.size(); ++r) tot[c.regs[r].combo_id] += mix_pows[r] * Fp::raw(rows[c.regs[r].group][c.regs[r].column]);
    for (size_t i = 0; i < CHECK_SIZE; ++i) tot[combos_size] += mix_pows[c.regs.size() + i] * Fp::raw(check_row[i]);
    Fp4 goal;
    const Fp4 x4 = Fp4::from_base(x);
    for (size_t i = 0; i < combos_size; ++i) {
      Fp4 num = tot[i] - poly_eval(&combo_u[combo_begin[i]], c.combos[i].size(), x4);
      Fp4 divisor = Fp4::one();
      for (uint32_t back : c.combos[i]) divisor *= x4 - z * pow(back_one, back);
      goal += num * inv(divisor);
    }
    goal += (tot[combos_size] - combo_u[combo_begin[combos_size]]) * inv(x4 - z_pow);
    // fold rounds
    for (const Round& r : rounds) {
      const size_t groups = r.dom / FRI_FOLD;
      const size_t quot = pos / groups, group = pos % groups;
      const uint32_t* data = r.m.verify(iop, group);
      Fp4 ext[FRI_FOLD];
      for (size_t i = 0; i < FRI_FOLD; ++i) ext[i] = Fp4::raw(data[i], data[FRI_FOLD + i], data[2 * FRI_FOLD + i], data[3 * FRI_FOLD + i]);
      VFY(ext[quot] == goal, "FRI query value does not match the previous round");
      int rpo2 = 0; while (((size_t)1 << rpo2) < r.dom) ++rpo2;
      const Fp inv_wk = pow(rou_rev(rpo2), group);
      interpolate16(ext);
      goal = poly_eval(ext, FRI_FOLD, r.mix * inv_wk);
      pos = group;
    }
    const Fp4 fx = poly_eval(final_poly.data(), degree, Fp4::from_base(pow(gen_final, pos)));
    VFY(fx == goal, "final FRI polynomial does not match the folded query");
  }
  VFY(iop.left == 0, "trailing words after the proof");
}

}  // namespace
}  // namespace zkb

using namespace zkb;

extern "C" {

zkb_err zkb_verify_segment(const uint32_t* h_circuit, size_t circuit_words, const uint32_t* h_seal, size_t seal_words,
                           const uint32_t* h_control_ids, size_t n_control_ids, uint32_t* h_out_po2_code_root) {
  ZKB_API_BEGIN
  ZKB_REQUIRE(h_circuit && h_seal, "null argument");
  ZKB_REQUIRE(h_control_ids || n_control_ids == 0, "null control ID table");
  ZKB_REQUIRE(n_control_ids > 0 || h_out_po2_code_root,
              "zkb_verify_segment needs the circuit's control IDs (po2, code root), or an output buffer so the caller can check the code root itself");
  CircuitDef c = CircuitDef::parse(h_circuit, circuit_words);
  verify_segment(c, h_seal, seal_words, ControlCheck{h_control_ids, n_control_ids, h_out_po2_code_root});
  ZKB_API_END
}

}  // extern "C"
