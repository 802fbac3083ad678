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
  std::vector<Step> wsteps;      // witness ("accumulate") program, DESIGN.md "circuit blob"

  size_t tap_size() const { return taps.size(); }
  size_t combos_size() const { return combos.size(); }
  size_t group_tap_begin(uint32_t g) const { size_t i = 0; while (i < taps.size() && taps[i].group < g) ++i; return i; }
  size_t group_tap_end(uint32_t g) const { size_t i = group_tap_begin(g); while (i < taps.size() && taps[i].group == g) ++i; return i; }

  static Circuit parse(const uint32_t* w, size_t len) {
    if (len < CIRCUIT_HEADER_WORDS || w[0] != CIRCUIT_MAGIC) throw std::runtime_error("circuit blob: bad magic");
    Circuit c;
    for (int g = 0; g < 3; ++g) c.group_size[g] = w[1 + g];
    c.mix_size = w[4]; c.out_size = w[5];
    size_t n_taps = w[6], n_steps = w[7], n_wsteps = w[11];
    c.ret = w[8];
    for (int i = 0; i < 16; ++i) c.info[i] = (uint8_t)(w[12 + i / 4] >> (8 * (i % 4)));
    if (len != CIRCUIT_HEADER_WORDS + 3 * n_taps + 4 * n_steps + 4 * n_wsteps) throw std::runtime_error("circuit blob: bad length");
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
    for (size_t i = 0; i < n_wsteps; ++i, p += 4) c.wsteps.push_back(Step{p[0], p[1], p[2], p[3]});      // witness program: checked where it is run
    return c;
  }
};

// Witness program interpreter: CircuitHal::accumulate as data (risc0-circuit-rv32im `prove/hal/cpu.rs` accumulate: the circuit's
// generated step_compute_accum run over every cycle, then Hal::prefix_products).  Phases (separated by Barrier / PrefixProduct) run
// one after the other, rows within a phase independently; a phase reads accum columns only as earlier phases (or the caller) left them.
enum : uint32_t { W_CONST = 0, W_GET = 1, W_GET_GLOBAL = 2, W_ADD = 3, W_SUB = 4, W_MUL = 5, W_SET = 6, W_BARRIER = 7, W_PREFIX_PRODUCT = 8 };
static void accumulate(const Circuit& c, Fp* accum, const Fp* code, const Fp* data, const Fp* mix_g, const Fp* out_g, int po2) {
  const size_t n = (size_t)1 << po2;
  const Fp* groups[3] = {accum, code, data};
  size_t lo = 0, first_val = 0;
  while (lo < c.wsteps.size()) {
    size_t hi = lo;
    while (hi < c.wsteps.size() && c.wsteps[hi].op != W_BARRIER && c.wsteps[hi].op != W_PREFIX_PRODUCT) ++hi;
    size_t n_vals = 0;
    for (size_t k = lo; k < hi; ++k) n_vals += c.wsteps[k].op <= W_MUL;
    // sets are buffered per phase so that a row never sees another row's write of the same phase, whatever the parser allowed
    std::vector<std::pair<size_t, Fp>> writes;
#pragma omp parallel
    {
      std::vector<Fp> v(n_vals);
      std::vector<std::pair<size_t, Fp>> mine;
#pragma omp for schedule(static)
      for (long long row = 0; row < (long long)n; ++row) {
        size_t vi = 0;
        for (size_t k = lo; k < hi; ++k) {
          const Step& s = c.wsteps[k];
          auto val = [&](uint32_t id) -> Fp { if (id < first_val || id - first_val >= vi) throw std::runtime_error("witness operand outside its phase"); return v[id - first_val]; };
          switch (s.op) {
            case W_CONST: v[vi++] = Fp::from(s.a); break;
            case W_GET: v[vi++] = groups[s.a][(size_t)s.b * n + (((size_t)row + n - (s.c % n)) & (n - 1))]; break;
            case W_GET_GLOBAL: v[vi++] = s.a == 0 ? mix_g[s.b] : out_g[s.b]; break;
            case W_ADD: v[vi++] = val(s.a) + val(s.b); break;
            case W_SUB: v[vi++] = val(s.a) - val(s.b); break;
            case W_MUL: v[vi++] = val(s.a) * val(s.b); break;
            case W_SET: if (s.c == 0xffffffffu || val(s.c).v != 0) mine.emplace_back((size_t)s.a * n + (size_t)row, val(s.b)); break;
            default: throw std::runtime_error("bad witness opcode");
          }
        }
      }
#pragma omp critical
      writes.insert(writes.end(), mine.begin(), mine.end());
    }
    for (auto& w : writes) accum[w.first] = w.second;
    first_val += n_vals;
    if (hi < c.wsteps.size() && c.wsteps[hi].op == W_PREFIX_PRODUCT) {      // 4 planar accum columns = one Fp4 column: inclusive product scan
      Fp* col = accum + (size_t)c.wsteps[hi].a * n;
      Fp4 cur = Fp4::one();
      for (size_t i = 0; i < n; ++i) {
        cur *= Fp4(col[i], col[n + i], col[2 * n + i], col[3 * n + i]);
        col[i] = cur.c[0]; col[n + i] = cur.c[1]; col[2 * n + i] = cur.c[2]; col[3 * n + i] = cur.c[3];
      }
    }
    lo = hi + 1;
  }
}

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
