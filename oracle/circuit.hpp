// ORACLE (test infrastructure, NOT product code) -- parity unpinned vs. the Rust crates.
//
// Circuit description consumed by the oracle prover: TapSet (risc0-zkp/src/taps.rs) and the
// PolyExtStep program (risc0-zkp/src/adapter.rs), SURVEY.md App. C.12 / D.1.  The rv32im circuit's
// generated `poly_fp` is not obtainable offline, so circuits arrive as data: a u32
// blob whose layout is documented in DESIGN.md ("circuit blob") and parsed independently here and
// in the product (zktls_b200/csrc/circuit.hpp).
#pragma once
#include "field.hpp"
#include <vector>
#include <stdexcept>
#include <algorithm>
#include <string>

namespace orc {

enum StepOp : uint32_t { OP_CONST = 0, OP_GET = 1, OP_GET_GLOBAL = 2, OP_ADD = 3, OP_SUB = 4, OP_MUL = 5, OP_TRUE = 6, OP_AND_EQZ = 7, OP_AND_COND = 8 };
static const uint32_t CIRCUIT_MAGIC = 0x5a4b4331u;
static const size_t CIRCUIT_HEADER_WORDS = 16;

struct Tap { uint32_t group, column, back; };
struct Register { uint32_t group, column, tap_pos, size, combo_id; };   // taps [tap_pos, tap_pos+size)
struct Step { uint32_t op, a, b, c; };

struct Circuit {
  uint32_t group_size[3];     // accum(0), code(1), data(2)
  uint32_t mix_size, out_size, ret;
  uint8_t info[16];
  std::vector<Tap> taps;
  std::vector<Register> regs;
  std::vector<std::vector<uint32_t>> combos;   // lexicographically sorted distinct back-lists
  std::vector<Step> steps;
  size_t n_fp_vars = 0, n_mix_vars = 0;

  size_t tap_size() const { return taps.size(); }
  size_t combos_size() const { return combos.size(); }
  size_t group_tap_begin(uint32_t g) const { size_t i = 0; while (i < taps.size() && taps[i].group < g) ++i; return i; }
  size_t group_tap_end(uint32_t g) const { size_t i = group_tap_begin(g); while (i < taps.size() && taps[i].group == g) ++i; return i; }

  static Circuit parse(const uint32_t* w, size_t len) {
    if (len < CIRCUIT_HEADER_WORDS || w[0] != CIRCUIT_MAGIC) throw std::runtime_error("circuit blob: bad magic");
    Circuit c;
    for (int g = 0; g < 3; ++g) c.group_size[g] = w[1 + g];
    c.mix_size = w[4]; c.out_size = w[5];
    size_t n_taps = w[6], n_steps = w[7];
    c.ret = w[8];
    for (int i = 0; i < 16; ++i) c.info[i] = (uint8_t)(w[12 + i / 4] >> (8 * (i % 4)));
    if (len != CIRCUIT_HEADER_WORDS + 3 * n_taps + 4 * n_steps) throw std::runtime_error("circuit blob: bad length");
    const uint32_t* p = w + CIRCUIT_HEADER_WORDS;
    for (size_t i = 0; i < n_taps; ++i, p += 3) {
      Tap t{p[0], p[1], p[2]};
      if (t.group > 2 || t.column >= c.group_size[t.group]) throw std::runtime_error("circuit blob: tap out of range");
      if (i) {
        const Tap& q = c.taps.back();
        bool ok = (q.group < t.group) || (q.group == t.group && (q.column < t.column || (q.column == t.column && q.back < t.back)));
        if (!ok) throw std::runtime_error("circuit blob: taps not sorted");
      }
      c.taps.push_back(t);
    }
    // registers + combos
    std::vector<std::vector<uint32_t>> reg_backs;
    for (size_t i = 0; i < n_taps;) {
      size_t j = i; std::vector<uint32_t> backs;
      while (j < n_taps && c.taps[j].group == c.taps[i].group && c.taps[j].column == c.taps[i].column) backs.push_back(c.taps[j++].back);
      c.regs.push_back(Register{c.taps[i].group, c.taps[i].column, (uint32_t)i, (uint32_t)(j - i), 0});
      reg_backs.push_back(backs);
      i = j;
    }
    c.combos = reg_backs;
    std::sort(c.combos.begin(), c.combos.end());
    c.combos.erase(std::unique(c.combos.begin(), c.combos.end()), c.combos.end());
    for (size_t r = 0; r < c.regs.size(); ++r)
      c.regs[r].combo_id = (uint32_t)(std::lower_bound(c.combos.begin(), c.combos.end(), reg_backs[r]) - c.combos.begin());
    for (size_t i = 0; i < n_steps; ++i, p += 4) {
      Step s{p[0], p[1], p[2], p[3]};
      switch (s.op) {
        case OP_CONST: break;
        case OP_GET: if (s.a >= n_taps) throw std::runtime_error("circuit blob: Get out of range"); break;
        case OP_GET_GLOBAL: if (s.a > 1 || s.b >= (s.a == 0 ? c.mix_size : c.out_size)) throw std::runtime_error("circuit blob: GetGlobal out of range"); break;
        case OP_ADD: case OP_SUB: case OP_MUL: if (s.a >= c.n_fp_vars || s.b >= c.n_fp_vars) throw std::runtime_error("circuit blob: fp operand out of range"); break;
        case OP_TRUE: break;
        case OP_AND_EQZ: if (s.a >= c.n_mix_vars || s.b >= c.n_fp_vars) throw std::runtime_error("circuit blob: AndEqz operand out of range"); break;
        case OP_AND_COND: if (s.a >= c.n_mix_vars || s.b >= c.n_fp_vars || s.c >= c.n_mix_vars) throw std::runtime_error("circuit blob: AndCond operand out of range"); break;
        default: throw std::runtime_error("circuit blob: bad opcode");
      }
      if (s.op <= OP_MUL) ++c.n_fp_vars; else ++c.n_mix_vars;
      c.steps.push_back(s);
    }
    if (c.ret >= c.n_mix_vars) throw std::runtime_error("circuit blob: ret out of range");
    return c;
  }
};

struct MixState { Fp4 tot, mul; };

// PolyExtStep interpreter, generic in the tap value type: Fp in the prover (evaluations on the LDE
// domain), Fp4 in the verifier (values derived from coeff_u).  Semantics per adapter.rs (App. C.12).
template <typename V, typename GetTap>
static Fp4 poly_ext(const Circuit& c, const Fp4& poly_mix, GetTap get_tap, const Fp* mix_globals, const Fp* out_globals,
                    std::vector<V>& fp_vars, std::vector<MixState>& mix_vars);

template <typename V> struct ValOps;
template <> struct ValOps<Fp> {
  static Fp from_fp(Fp x) { return x; }
  static Fp4 scale(const Fp4& m, Fp v) { return m * v; }
};
template <> struct ValOps<Fp4> {
  static Fp4 from_fp(Fp x) { return Fp4::from_base(x); }
  static Fp4 scale(const Fp4& m, const Fp4& v) { return m * v; }
};

template <typename V, typename GetTap>
static Fp4 poly_ext(const Circuit& c, const Fp4& poly_mix, GetTap get_tap, const Fp* mix_globals, const Fp* out_globals,
                    std::vector<V>& fp_vars, std::vector<MixState>& mix_vars) {
  fp_vars.clear(); mix_vars.clear();
  for (const Step& s : c.steps) {
    switch (s.op) {
      case OP_CONST: fp_vars.push_back(ValOps<V>::from_fp(Fp::from(s.a))); break;
      case OP_GET: fp_vars.push_back(get_tap(s.a)); break;
      case OP_GET_GLOBAL: fp_vars.push_back(ValOps<V>::from_fp(s.a == 0 ? mix_globals[s.b] : out_globals[s.b])); break;
      case OP_ADD: fp_vars.push_back(fp_vars[s.a] + fp_vars[s.b]); break;
      case OP_SUB: fp_vars.push_back(fp_vars[s.a] - fp_vars[s.b]); break;
      case OP_MUL: fp_vars.push_back(fp_vars[s.a] * fp_vars[s.b]); break;
      case OP_TRUE: mix_vars.push_back(MixState{Fp4::zero(), Fp4::one()}); break;
      case OP_AND_EQZ: {
        MixState x = mix_vars[s.a];
        mix_vars.push_back(MixState{x.tot + ValOps<V>::scale(x.mul, fp_vars[s.b]), x.mul * poly_mix});
        break;
      }
      case OP_AND_COND: {
        MixState x = mix_vars[s.a], y = mix_vars[s.c];
        mix_vars.push_back(MixState{x.tot + ValOps<V>::scale(y.tot * x.mul, fp_vars[s.b]), x.mul * y.mul});
        break;
      }
    }
  }
  return mix_vars[c.ret].tot;
}

}  // namespace orc
