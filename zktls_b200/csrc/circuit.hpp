// Circuit description for the B200 backend: TapSet + PolyExtStep program, delivered as data.
//
// Mirrors risc0-zkp 1.2.5 `taps::TapSet` and `adapter::{PolyExtStep, PolyExtStepDef}` (un-vendored;
// SURVEY.md App. C.12, D.1).  In the reference the constraint polynomial is generated C++/CUDA (`poly_fp`);
// here the same information arrives as a u32 blob (layout in DESIGN.md, "circuit blob") so that the rv32im
// definition can be dropped in as data once a Rust host exists.
#pragma once
#include <algorithm>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <tuple>
#include <vector>
#include "field.cuh"

namespace zkb {

enum : uint32_t { PX_CONST = 0, PX_GET = 1, PX_GET_GLOBAL = 2, PX_ADD = 3, PX_SUB = 4, PX_MUL = 5, PX_TRUE = 6, PX_AND_EQZ = 7, PX_AND_COND = 8 };
constexpr uint32_t CIRCUIT_MAGIC = 0x5a4b4331u;   // "ZKC1"
constexpr size_t CIRCUIT_HEADER_WORDS = 16;
constexpr int GROUP_ACCUM = 0, GROUP_CODE = 1, GROUP_DATA = 2, NUM_GROUPS = 3;

struct TapDef { uint32_t group, column, back; };
constexpr uint32_t MAX_TAP_BACK = 1u << 12;
struct RegisterDef { uint32_t group, column, tap_pos, size, combo_id; };
struct StepDef { uint32_t op, a, b, c; };
// witness ("accumulate") program: CircuitHal::accumulate as data -- the value opcodes of PolyExtStep + Set / Barrier / PrefixProduct
enum : uint32_t { WX_CONST = 0, WX_GET = 1, WX_GET_GLOBAL = 2, WX_ADD = 3, WX_SUB = 4, WX_MUL = 5, WX_SET = 6, WX_BARRIER = 7, WX_PREFIX_PRODUCT = 8 };
constexpr uint32_t WX_ALWAYS = 0xffffffffu;

struct CircuitDef {
  uint32_t group_size[NUM_GROUPS] = {0, 0, 0};
  uint32_t mix_size = 0, out_size = 0, ret = 0;
  uint8_t info[16] = {0};
  std::vector<TapDef> taps;
  std::vector<RegisterDef> regs;
  std::vector<std::vector<uint32_t>> combos;
  std::vector<StepDef> steps;
  uint32_t n_fp_vars = 0, n_mix_vars = 0;
  std::vector<StepDef> wsteps;      // witness program of the accum group (may be empty: the circuit then has no device-side accumulate)

  size_t tap_size() const { return taps.size(); }
  size_t combos_size() const { return combos.size(); }
  size_t group_tap_begin(uint32_t g) const { size_t i = 0; while (i < taps.size() && taps[i].group < g) ++i; return i; }
  size_t group_tap_end(uint32_t g) const { size_t i = group_tap_begin(g); while (i < taps.size() && taps[i].group == g) ++i; return i; }

  static void fail(const std::string& m) { throw std::runtime_error("zkb200: circuit blob: " + m); }

  static CircuitDef parse(const uint32_t* w, size_t len) {
    if (!w || len < CIRCUIT_HEADER_WORDS) fail("too short");
    if (w[0] != CIRCUIT_MAGIC) fail("bad magic");
    CircuitDef c;
    for (int g = 0; g < NUM_GROUPS; ++g) c.group_size[g] = w[1 + g];
    c.mix_size = w[4]; c.out_size = w[5];
    const size_t n_taps = w[6], n_steps = w[7];
    c.ret = w[8];
    for (int i = 0; i < 16; ++i) c.info[i] = (uint8_t)(w[12 + i / 4] >> (8 * (i % 4)));
    const size_t n_wsteps = w[11];
    if (len != CIRCUIT_HEADER_WORDS + 3 * n_taps + 4 * n_steps + 4 * n_wsteps) fail("length does not match the header counts");
    const uint32_t* p = w + CIRCUIT_HEADER_WORDS;
    c.taps.reserve(n_taps);
    for (size_t i = 0; i < n_taps; ++i, p += 3) {
      TapDef t{p[0], p[1], p[2]};
      if (t.group >= NUM_GROUPS || t.column >= c.group_size[t.group]) fail("tap out of range");
      if (t.back >= MAX_TAP_BACK) fail("tap reaches too far back");      // halo = 4 * back words per staged column; rv32im taps reach back <= 5
      if (i) {
        const TapDef& q = c.taps.back();
        if (std::tie(q.group, q.column, q.back) >= std::tie(t.group, t.column, t.back)) fail("taps must be strictly sorted by (group, column, back)");
      }
      c.taps.push_back(t);
    }
    // registers = runs of taps on one column; combos = sorted distinct back-lists
    std::vector<std::vector<uint32_t>> backs;
    for (size_t i = 0; i < n_taps;) {
      size_t j = i;
      std::vector<uint32_t> b;
      while (j < n_taps && c.taps[j].group == c.taps[i].group && c.taps[j].column == c.taps[i].column) b.push_back(c.taps[j++].back);
      c.regs.push_back(RegisterDef{c.taps[i].group, c.taps[i].column, (uint32_t)i, (uint32_t)(j - i), 0});
      backs.push_back(std::move(b));
      i = j;
    }
    c.combos = backs;
    std::sort(c.combos.begin(), c.combos.end());
    c.combos.erase(std::unique(c.combos.begin(), c.combos.end()), c.combos.end());
    for (size_t r = 0; r < c.regs.size(); ++r)
      c.regs[r].combo_id = (uint32_t)(std::lower_bound(c.combos.begin(), c.combos.end(), backs[r]) - c.combos.begin());
    // every column of every group must be covered by a register, in order (mix_poly_coeffs walks columns)
    {
      size_t r = 0;
      for (uint32_t g = 0; g < NUM_GROUPS; ++g)
        for (uint32_t col = 0; col < c.group_size[g]; ++col, ++r)
          if (r >= c.regs.size() || c.regs[r].group != g || c.regs[r].column != col) fail("every column needs at least one tap");
    }
    c.steps.reserve(n_steps);
    for (size_t i = 0; i < n_steps; ++i, p += 4) {
      StepDef s{p[0], p[1], p[2], p[3]};
      switch (s.op) {
        case PX_CONST: if (s.a >= P) fail("Const not canonical"); break;
        case PX_GET: if (s.a >= n_taps) fail("Get out of range"); break;
        case PX_GET_GLOBAL: if (s.a > 1 || s.b >= (s.a == 0 ? c.mix_size : c.out_size)) fail("GetGlobal out of range"); break;
        case PX_ADD: case PX_SUB: case PX_MUL: if (s.a >= c.n_fp_vars || s.b >= c.n_fp_vars) fail("operand refers to a later value"); break;
        case PX_TRUE: break;
        case PX_AND_EQZ: if (s.a >= c.n_mix_vars || s.b >= c.n_fp_vars) fail("AndEqz operand out of range"); break;
        case PX_AND_COND: if (s.a >= c.n_mix_vars || s.b >= c.n_fp_vars || s.c >= c.n_mix_vars) fail("AndCond operand out of range"); break;
        default: fail("unknown opcode");
      }
      if (s.op <= PX_MUL) ++c.n_fp_vars; else ++c.n_mix_vars;
      c.steps.push_back(s);
    }
    if (c.ret >= c.n_mix_vars) fail("ret out of range");
    // witness program: values are numbered in definition order and live inside one phase (between barriers); a phase must not
    // read an accum column it writes (rows run in parallel)
    {
      uint32_t n_vals = 0, phase_first = 0;
      std::vector<char> set_here(c.group_size[GROUP_ACCUM], 0), got_here(c.group_size[GROUP_ACCUM], 0);
      auto val_ok = [&](uint32_t v) { return v >= phase_first && v < n_vals; };
      for (size_t i = 0; i < n_wsteps; ++i, p += 4) {
        StepDef s{p[0], p[1], p[2], p[3]};
        switch (s.op) {
          case WX_CONST: if (s.a >= P) fail("witness Const not canonical"); ++n_vals; break;
          case WX_GET:
            if (s.a >= NUM_GROUPS || s.b >= c.group_size[s.a] || s.c >= MAX_TAP_BACK) fail("witness Get out of range");
            if (s.a == GROUP_ACCUM) { if (set_here[s.b]) fail("witness phase reads an accum column it writes"); got_here[s.b] = 1; }
            ++n_vals; break;
          case WX_GET_GLOBAL: if (s.a > 1 || s.b >= (s.a == 0 ? c.mix_size : c.out_size)) fail("witness GetGlobal out of range"); ++n_vals; break;
          case WX_ADD: case WX_SUB: case WX_MUL: if (!val_ok(s.a) || !val_ok(s.b)) fail("witness operand outside its phase"); ++n_vals; break;
          case WX_SET:
            if (s.a >= c.group_size[GROUP_ACCUM] || !val_ok(s.b) || (s.c != WX_ALWAYS && !val_ok(s.c))) fail("witness Set out of range");
            if (got_here[s.a]) fail("witness phase writes an accum column it reads");
            set_here[s.a] = 1; break;
          case WX_PREFIX_PRODUCT: if ((size_t)s.a + 4 > c.group_size[GROUP_ACCUM]) fail("witness PrefixProduct out of range");     // fall through: acts as a barrier
          case WX_BARRIER: phase_first = n_vals; std::fill(set_here.begin(), set_here.end(), 0); std::fill(got_here.begin(), got_here.end(), 0); break;
          default: fail("unknown witness opcode");
        }
        c.wsteps.push_back(s);
      }
    }
    return c;
  }
};

// Host interpreter over Fp4 tap values -- the verifier's use of the constraint system (adapter.rs poly_ext with
// ExtElem inputs).  `u[tap]` are the tap evaluations at the DEEP point.
inline Fp4 poly_ext_host(const CircuitDef& c, const Fp4& poly_mix, const Fp4* u, const uint32_t* mix_g, const uint32_t* out_g) {
  struct Mix { Fp4 tot, mul; };
  std::vector<Fp4> fp; fp.reserve(c.n_fp_vars);
  std::vector<Mix> mx; mx.reserve(c.n_mix_vars);
  for (const StepDef& s : c.steps) {
    switch (s.op) {
      case PX_CONST: fp.push_back(Fp4::from_base(Fp::from(s.a))); break;
      case PX_GET: fp.push_back(u[s.a]); break;
      case PX_GET_GLOBAL: fp.push_back(Fp4::from_base(Fp::raw(s.a == 0 ? mix_g[s.b] : out_g[s.b]))); break;
      case PX_ADD: fp.push_back(fp[s.a] + fp[s.b]); break;
      case PX_SUB: fp.push_back(fp[s.a] - fp[s.b]); break;
      case PX_MUL: fp.push_back(fp[s.a] * fp[s.b]); break;
      case PX_TRUE: mx.push_back(Mix{Fp4::zero(), Fp4::one()}); break;
      case PX_AND_EQZ: { Mix x = mx[s.a]; mx.push_back(Mix{x.tot + x.mul * fp[s.b], x.mul * poly_mix}); break; }
      case PX_AND_COND: { Mix x = mx[s.a], y = mx[s.c]; mx.push_back(Mix{x.tot + fp[s.b] * y.tot * x.mul, x.mul * y.mul}); break; }
    }
  }
  return mx[c.ret].tot;
}

}  // namespace zkb
