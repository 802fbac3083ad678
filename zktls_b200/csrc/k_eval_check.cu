// eval_check: the constraint ("check") polynomial over the LDE domain (sm_100a).
//
// Stand-in for `CircuitHal::eval_check` (risc0-circuit-rv32im/src/prove/hal/{cpu,cuda}.rs + the generated
// `poly_fp`):   check[j*4n + c] = ( poly_fp(c) * ((3 w_4n^c)^n - 1)^-1 )[j].
// The constraint system is data (circuit.hpp).  The host lowers the PolyExtStep list to a small register VM:
//   * GetGlobal and Const become immediates, Get becomes a (group, column, back) operand fetched straight from the
//     column-major LDE matrix -- consecutive threads read consecutive rows, so every fetch is one coalesced line;
//   * Add/Sub/Mul results live in per-thread slots in shared memory ([slot][thread], conflict-free), allocated by
//     liveness so a few slots suffice; mix states keep only `tot` (their `mul` is a static power of poly_mix, looked up
//     in a precomputed table, exactly as the generated reference code indexes `poly_mix[k]`);
//   * (3x)^n takes 4 values (c mod 4), so the division is one Fp4 x Fp multiply by a launch-time constant.
// Control flow is uniform across the grid (every thread runs the same program), so there is no divergence.
#include "common.cuh"
#include "circuit.hpp"
#include <tuple>

namespace zkb {

enum : uint32_t { K_CONST = 0, K_TAP = 1, K_TEMP = 2 };
enum : uint32_t { VM_ADD = 0, VM_SUB = 1, VM_MUL = 2, VM_EQZ = 3, VM_COND = 4 };
constexpr uint32_t MIX_ZERO = 0xff;
constexpr int EC_BLOCK = 256;

struct EvalProgram {
  std::vector<uint4> code;
  uint32_t n_slots = 0, n_mix_slots = 0, n_powers = 0;
  uint32_t ret_slot = MIX_ZERO;
};

// Lower the PolyExtStep program (see header comment).  Throws on programs that need more live temporaries than
// fit in shared memory.
static EvalProgram lower(const CircuitDef& c, const uint32_t* mix_g, const uint32_t* out_g) {
  struct Val { uint32_t kind, payload; };
  const size_t n = c.steps.size();
  // last use of every fp / mix var
  std::vector<int64_t> fp_last(c.n_fp_vars, -1), mx_last(c.n_mix_vars, -1);
  {
    uint32_t fi = 0, mi = 0;
    for (size_t i = 0; i < n; ++i) {
      const StepDef& s = c.steps[i];
      switch (s.op) {
        case PX_ADD: case PX_SUB: case PX_MUL: fp_last[s.a] = fp_last[s.b] = (int64_t)i; break;
        case PX_AND_EQZ: mx_last[s.a] = (int64_t)i; fp_last[s.b] = (int64_t)i; break;
        case PX_AND_COND: mx_last[s.a] = mx_last[s.c] = (int64_t)i; fp_last[s.b] = (int64_t)i; break;
        default: break;
      }
      if (s.op <= PX_MUL) ++fi; else ++mi;
    }
    mx_last[c.ret] = (int64_t)n;   // live to the end
  }
  EvalProgram prog;
  std::vector<Val> fp(c.n_fp_vars);
  std::vector<uint32_t> mx_slot(c.n_mix_vars, MIX_ZERO), mx_pow(c.n_mix_vars, 0);
  std::vector<uint32_t> free_slots, free_mix;
  auto alloc = [&](std::vector<uint32_t>& fl, uint32_t& count) { if (!fl.empty()) { uint32_t s = fl.back(); fl.pop_back(); return s; } return count++; };
  auto release_fp = [&](uint32_t var, size_t i) { if (fp[var].kind == K_TEMP && fp_last[var] == (int64_t)i) { free_slots.push_back(fp[var].payload); fp_last[var] = -2; } };
  auto release_mx = [&](uint32_t var, size_t i) { if (mx_slot[var] != MIX_ZERO && mx_last[var] == (int64_t)i) { free_mix.push_back(mx_slot[var]); mx_last[var] = -2; } };
  uint32_t fi = 0, mi = 0;
  for (size_t i = 0; i < n; ++i) {
    const StepDef& s = c.steps[i];
    switch (s.op) {
      case PX_CONST: fp[fi++] = Val{K_CONST, Fp::from(s.a).v}; break;
      case PX_GET_GLOBAL: fp[fi++] = Val{K_CONST, s.a == 0 ? mix_g[s.b] : out_g[s.b]}; break;
      case PX_GET: {
        const TapDef& t = c.taps[s.a];
        if (t.column >= (1u << 20) || t.back >= (1u << 10)) CircuitDef::fail("tap does not fit the VM encoding");
        fp[fi++] = Val{K_TAP, t.column | (t.group << 20) | (t.back << 22)};
        break;
      }
      case PX_ADD: case PX_SUB: case PX_MUL: {
        Val a = fp[s.a], b = fp[s.b];
        if (fp_last[fi] < 0) { fp[fi++] = Val{K_CONST, 0}; break; }     // dead value
        if (a.kind == K_CONST && b.kind == K_CONST) {                    // constant folding
          Fp x = Fp::raw(a.payload), y = Fp::raw(b.payload);
          Fp r = s.op == PX_ADD ? x + y : s.op == PX_SUB ? x - y : x * y;
          fp[fi++] = Val{K_CONST, r.v};
          break;
        }
        release_fp(s.a, i); if (s.b != s.a) release_fp(s.b, i);
        uint32_t dst = alloc(free_slots, prog.n_slots);
        uint32_t op = s.op == PX_ADD ? VM_ADD : s.op == PX_SUB ? VM_SUB : VM_MUL;
        prog.code.push_back(make_uint4(op | (a.kind << 8) | (b.kind << 12) | (dst << 16), a.payload, b.payload, 0));
        fp[fi++] = Val{K_TEMP, dst};
        break;
      }
      case PX_TRUE: mx_slot[mi] = MIX_ZERO; mx_pow[mi] = 0; ++mi; break;
      case PX_AND_EQZ: {
        Val v = fp[s.b];
        uint32_t src = mx_slot[s.a], pw = mx_pow[s.a];
        mx_pow[mi] = pw + 1;
        if (mx_last[mi] < 0) { mx_slot[mi] = MIX_ZERO; ++mi; break; }   // dead chain
        release_fp(s.b, i); release_mx(s.a, i);
        uint32_t dst = alloc(free_mix, prog.n_mix_slots);
        prog.code.push_back(make_uint4(VM_EQZ | (v.kind << 8) | (dst << 16), v.payload, src, pw));
        prog.n_powers = std::max(prog.n_powers, pw + 1);
        mx_slot[mi++] = dst;
        break;
      }
      case PX_AND_COND: {
        Val v = fp[s.b];
        uint32_t xs = mx_slot[s.a], ys = mx_slot[s.c], pw = mx_pow[s.a];
        mx_pow[mi] = pw + mx_pow[s.c];
        if (mx_last[mi] < 0) { mx_slot[mi] = MIX_ZERO; ++mi; break; }
        release_fp(s.b, i); release_mx(s.a, i); if (s.c != s.a) release_mx(s.c, i);
        uint32_t dst = alloc(free_mix, prog.n_mix_slots);
        prog.code.push_back(make_uint4(VM_COND | (v.kind << 8) | (dst << 16), v.payload, xs | (ys << 8), pw));
        prog.n_powers = std::max(prog.n_powers, pw + 1);
        mx_slot[mi++] = dst;
        break;
      }
    }
  }
  prog.ret_slot = mx_slot[c.ret];
  if (prog.n_mix_slots >= MIX_ZERO) CircuitDef::fail("too many live mix states");
  return prog;
}

struct EvalArgs {
  const uint32_t* groups[3];
  uint32_t inv_den[4];       // ((3 w_4n^r)^n - 1)^-1 for r = c mod 4
  uint32_t n_instr, n_slots, ret_slot;
  uint32_t domain_mask;
};

__global__ void __launch_bounds__(EC_BLOCK) k_eval_check(uint32_t* __restrict__ check, const uint4* __restrict__ code, const uint4* __restrict__ mixpow, EvalArgs args) {
  extern __shared__ uint32_t smem[];
  uint32_t* slots = smem;                                     // [n_slots][EC_BLOCK]
  uint32_t* mslots = smem + (size_t)args.n_slots * EC_BLOCK;  // [n_mix_slots][4][EC_BLOCK]
  const uint32_t tid = threadIdx.x;
  const uint32_t c = blockIdx.x * EC_BLOCK + tid;
  const size_t domain = (size_t)args.domain_mask + 1;

  auto fetch = [&](uint32_t kind, uint32_t payload) -> uint32_t {
    if (kind == K_CONST) return payload;
    if (kind == K_TEMP) return slots[payload * EC_BLOCK + tid];
    uint32_t col = payload & 0xfffffu, g = (payload >> 20) & 3u, back = payload >> 22;
    const uint32_t* base = g == 0 ? args.groups[0] : g == 1 ? args.groups[1] : args.groups[2];
    return __ldg(base + (size_t)col * domain + ((c - 4u * back) & args.domain_mask));
  };
  auto mix_load = [&](uint32_t s) -> Fp4 {
    if (s == MIX_ZERO) return Fp4::zero();
    const uint32_t* p = mslots + (size_t)s * 4 * EC_BLOCK + tid;
    return Fp4::raw(p[0], p[EC_BLOCK], p[2 * EC_BLOCK], p[3 * EC_BLOCK]);
  };
  auto mix_store = [&](uint32_t s, const Fp4& v) {
    uint32_t* p = mslots + (size_t)s * 4 * EC_BLOCK + tid;
    p[0] = v.c[0].v; p[EC_BLOCK] = v.c[1].v; p[2 * EC_BLOCK] = v.c[2].v; p[3 * EC_BLOCK] = v.c[3].v;
  };

  for (uint32_t pc = 0; pc < args.n_instr; ++pc) {
    const uint4 ins = __ldg(code + pc);
    const uint32_t op = ins.x & 0xffu, ka = (ins.x >> 8) & 0xfu, kb = (ins.x >> 12) & 0xfu, dst = ins.x >> 16;
    if (op <= VM_MUL) {
      uint32_t a = fetch(ka, ins.y), b = fetch(kb, ins.z);
      uint32_t r = op == VM_ADD ? add_mod(a, b) : op == VM_SUB ? sub_mod(a, b) : mont_mul(a, b);
      slots[dst * EC_BLOCK + tid] = r;
    } else if (op == VM_EQZ) {
      Fp v = Fp::raw(fetch(ka, ins.y));
      Fp4 pw = ld4(__ldg(mixpow + ins.w));
      mix_store(dst, mix_load(ins.z) + pw * v);
    } else {   // VM_COND
      Fp v = Fp::raw(fetch(ka, ins.y));
      Fp4 pw = ld4(__ldg(mixpow + ins.w));
      Fp4 inner = mix_load((ins.z >> 8) & 0xffu);
      mix_store(dst, mix_load(ins.z & 0xffu) + (inner * pw) * v);
    }
  }
  Fp4 tot = mix_load(args.ret_slot) * Fp::raw(args.inv_den[c & 3u]);
  check[c] = tot.c[0].v; check[domain + c] = tot.c[1].v; check[2 * domain + c] = tot.c[2].v; check[3 * domain + c] = tot.c[3].v;
}

bool eval_check_jit(zkb_ctx* ctx, uint32_t* d_check, const CircuitDef& c, const uint32_t* const d_groups[3], const uint32_t* mix_g, const uint32_t* out_g,
                    const Fp4& poly_mix, int po2, std::string& why);   // k_eval_jit.cu

// ZKB_EVAL_CHECK = jit (default: NVRTC-specialised kernel, interpreter if NVRTC is unavailable) | vm (interpreter) |
// jit-only (fail instead of falling back; used by the tests to make sure the JIT path is the one exercised)
void eval_check(zkb_ctx* ctx, uint32_t* d_check, const CircuitDef& c, const uint32_t* const d_groups[3], const uint32_t* mix_g, const uint32_t* out_g,
                const Fp4& poly_mix, int po2) {
  ZKB_REQUIRE(po2 >= 0 && po2 + 2 <= MAX_PO2, "eval_check: po2 out of range");
  {
    const char* mode = getenv("ZKB_EVAL_CHECK");
    const bool vm_only = mode && !strcmp(mode, "vm"), jit_only = mode && !strcmp(mode, "jit-only");
    if (!vm_only) {
      std::string why;
      if (eval_check_jit(ctx, d_check, c, d_groups, mix_g, out_g, poly_mix, po2, why)) return;
      if (jit_only) throw Error("zkb200: eval_check JIT unavailable: " + why);
      static bool warned = false;
      if (!warned) { warned = true; fprintf(stderr, "zkb200: eval_check JIT unavailable (%s); using the device interpreter\n", why.c_str()); }
    }
  }
  const size_t n = (size_t)1 << po2, domain = n * INV_RATE;
  ZKB_REQUIRE(domain % EC_BLOCK == 0 || domain < EC_BLOCK, "eval_check: domain too small");
  EvalProgram prog = lower(c, mix_g, out_g);
  size_t smem = ((size_t)prog.n_slots + 4 * (size_t)prog.n_mix_slots) * EC_BLOCK * 4;
  ZKB_REQUIRE(smem <= 200 * 1024, "eval_check: circuit needs too many live temporaries for the shared-memory VM");
  // powers of poly_mix
  std::vector<uint32_t> pw(4 * (size_t)std::max<uint32_t>(prog.n_powers, 1));
  Fp4 cur = Fp4::one();
  for (uint32_t i = 0; i < std::max<uint32_t>(prog.n_powers, 1); ++i) { cur.store(&pw[4 * i]); cur *= poly_mix; }
  uint4 *d_code = nullptr, *d_pw = nullptr;
  size_t code_bytes = std::max<size_t>(prog.code.size(), 1) * 16;
  pool_alloc(ctx, &d_code, code_bytes);
  pool_alloc(ctx, &d_pw, pw.size() * 4);
  if (!prog.code.empty()) ZKB_CUDA(cudaMemcpyAsync(d_code, prog.code.data(), prog.code.size() * 16, cudaMemcpyHostToDevice, ctx->stream));
  ZKB_CUDA(cudaMemcpyAsync(d_pw, pw.data(), pw.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
  EvalArgs args;
  for (int g = 0; g < 3; ++g) args.groups[g] = d_groups[g];
  Fp w4 = pow(Fp::from(137), (uint64_t)1 << (MAX_ROU_PO2 - 2));   // w_4
  Fp three_n = pow(Fp::from(3), n);
  // (3 w_4n^c)^n = 3^n * w_4^c
  Fp wr = Fp::one();
  for (int r = 0; r < 4; ++r) { args.inv_den[r] = inv(three_n * wr - Fp::one()).v; wr *= w4; }
  args.n_instr = (uint32_t)prog.code.size();
  args.n_slots = prog.n_slots;
  args.ret_slot = prog.ret_slot;
  args.domain_mask = (uint32_t)(domain - 1);
  unsigned block = EC_BLOCK;
  ZKB_REQUIRE(domain >= EC_BLOCK, "eval_check: domain smaller than one block (po2 >= 6 required)");
  if (smem > 48 * 1024) ZKB_CUDA(cudaFuncSetAttribute(k_eval_check, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_eval_check<<<(unsigned)(domain / block), block, smem, ctx->stream>>>(d_check, d_code, d_pw, args);
  launched(ctx);
  pool_free(ctx, d_code);
  pool_free(ctx, d_pw);
}

}  // namespace zkb

using namespace zkb;

namespace zkb { std::string eval_jit_source(const CircuitDef& c); bool eval_jit_compile_only(const CircuitDef& c, std::string& why); }

// Generated CUDA source of the specialised eval_check kernel (host-only; no device needed).  Writes up to `cap` bytes
// (NUL-terminated) and the full length (without NUL) to *needed.
extern "C" zkb_err zkb_eval_check_source(const uint32_t* h_circuit, size_t circuit_words, char* out, size_t cap, size_t* needed) {
  ZKB_API_BEGIN
  ZKB_REQUIRE(h_circuit && needed, "null argument");
  std::string src = eval_jit_source(CircuitDef::parse(h_circuit, circuit_words));
  *needed = src.size();
  if (out && cap) { size_t m = std::min(cap - 1, src.size()); memcpy(out, src.data(), m); out[m] = 0; }
  ZKB_API_END
}
// Compiles that source for sm_100a with NVRTC (host-only) and stores the cubin in the on-disk cache.
extern "C" zkb_err zkb_eval_check_precompile(const uint32_t* h_circuit, size_t circuit_words) {
  ZKB_API_BEGIN
  ZKB_REQUIRE(h_circuit, "null argument");
  std::string why;
  if (!eval_jit_compile_only(CircuitDef::parse(h_circuit, circuit_words), why)) throw Error("zkb200: eval_check precompile failed: " + why);
  ZKB_API_END
}

extern "C" zkb_err zkb_eval_check(zkb_ctx* ctx, void* d_check, const uint32_t* h_circuit, size_t circuit_words, const void* d_accum, const void* d_code,
                                  const void* d_data, const uint32_t* h_mix_g, const uint32_t* h_out_g, const uint32_t* h_poly_mix, int po2) {
  ZKB_API_BEGIN
  use(ctx);
  ZKB_REQUIRE(d_check && h_circuit && h_poly_mix, "null argument");
  CircuitDef c = CircuitDef::parse(h_circuit, circuit_words);
  ZKB_REQUIRE((h_mix_g || c.mix_size == 0) && (h_out_g || c.out_size == 0), "null globals");
  const uint32_t* groups[3] = {(const uint32_t*)d_accum, (const uint32_t*)d_code, (const uint32_t*)d_data};
  for (int g = 0; g < 3; ++g) ZKB_REQUIRE(groups[g] || c.group_size[g] == 0, "null group buffer");
  eval_check(ctx, (uint32_t*)d_check, c, groups, h_mix_g, h_out_g, Fp4::load(h_poly_mix), po2);
  ZKB_API_END
}
