// eval_check, specialised per circuit: the PolyExtStep program is turned into straight-line CUDA C++ and compiled for
// sm_100a with NVRTC the first time a circuit is seen.
//
// The reference gets its `poly_fp` the same way, only ahead of time: risc0-circuit-rv32im-sys compiles a generated
// C++/CUDA function of ~10^5 lines with nvcc at crate build time (un-vendored; call site
// /root/reference/crates/guest-prover-r0/src/prover.rs:90).  Here the constraint system arrives as data (circuit.hpp), so
// the specialisation happens at `zkb_prover_new` / first `zkb_eval_check`:
//   * every Fp temporary and every mix accumulator is an SSA value in registers (the interpreter in k_eval_check.cu
//     spends ~65 instructions of decode / shared-memory traffic per PolyExtStep; the compiled form spends 1-5);
//   * taps are read straight from the column-major LDE matrices, one coalesced 128-byte line per warp and tap;
//   * per-proof values (globals, powers of poly_mix, the four (3x)^n - 1 inverses) are kernel DATA, so one cubin serves
//     every segment of the circuit; cubins are cached in memory per ctx and on disk (ZKB_CACHE_DIR, default
//     /tmp/zkb200-cache) keyed by a hash of the generated source.
// libnvrtc / libcuda are dlopen'ed, not linked: when they are absent the caller falls back to the device interpreter
// (still CUDA -- there is no CPU path).
#include "common.cuh"
#include "circuit.hpp"
#include <cuda.h>
#include <nvrtc.h>
#include <dlfcn.h>
#include <sys/stat.h>
#include <unistd.h>
#include <fcntl.h>
#include <cerrno>
#include <fstream>
#include <sstream>
#include <mutex>
#include <set>

namespace zkb {

namespace {

struct Api {
  bool ok = false;        // NVRTC usable (compile)
  bool cu_ok = false;     // driver API usable (load + launch)
  std::string why, cu_why;
  decltype(&nvrtcCreateProgram) createProgram;
  decltype(&nvrtcCompileProgram) compileProgram;
  decltype(&nvrtcGetCUBINSize) getCUBINSize;
  decltype(&nvrtcGetCUBIN) getCUBIN;
  decltype(&nvrtcGetProgramLogSize) getLogSize;
  decltype(&nvrtcGetProgramLog) getLog;
  decltype(&nvrtcDestroyProgram) destroyProgram;
  decltype(&nvrtcVersion) version = nullptr;
  CUresult (*moduleLoadData)(CUmodule*, const void*);
  CUresult (*moduleGetFunction)(CUfunction*, CUmodule, const char*);
  CUresult (*moduleUnload)(CUmodule);
  CUresult (*launchKernel)(CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, CUstream, void**, void**);
  CUresult (*getErrorString)(CUresult, const char**);
  CUresult (*moduleGetGlobal)(CUdeviceptr*, size_t*, CUmodule, const char*);
  CUresult (*memcpyHtoDAsync)(CUdeviceptr, const void*, size_t, CUstream);
  CUresult (*funcSetAttribute)(CUfunction, CUfunction_attribute, int);
};

Api& api() {
  static Api a;
  static std::once_flag once;
  std::call_once(once, [] {
    void* rt = nullptr;
    for (const char* name : {"libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so.12"}) { rt = dlopen(name, RTLD_NOW | RTLD_GLOBAL); if (rt) break; }
    if (!rt) { a.why = "libnvrtc.so.12 not loadable"; return; }
    bool all = true;
    auto sym = [&](void* lib, const char* n) { void* p = dlsym(lib, n); if (!p) { all = false; a.why = std::string("missing symbol ") + n; } return p; };
    a.createProgram = (decltype(a.createProgram))sym(rt, "nvrtcCreateProgram");
    a.compileProgram = (decltype(a.compileProgram))sym(rt, "nvrtcCompileProgram");
    a.getCUBINSize = (decltype(a.getCUBINSize))sym(rt, "nvrtcGetCUBINSize");
    a.getCUBIN = (decltype(a.getCUBIN))sym(rt, "nvrtcGetCUBIN");
    a.getLogSize = (decltype(a.getLogSize))sym(rt, "nvrtcGetProgramLogSize");
    a.getLog = (decltype(a.getLog))sym(rt, "nvrtcGetProgramLog");
    a.destroyProgram = (decltype(a.destroyProgram))sym(rt, "nvrtcDestroyProgram");
    a.version = (decltype(a.version))dlsym(rt, "nvrtcVersion");
    a.ok = all;
    void* cu = dlopen("libcuda.so.1", RTLD_NOW | RTLD_GLOBAL);
    if (!cu) { a.cu_why = "libcuda.so.1 not loadable"; return; }
    all = true;
    a.moduleLoadData = (decltype(a.moduleLoadData))sym(cu, "cuModuleLoadData");
    a.moduleGetFunction = (decltype(a.moduleGetFunction))sym(cu, "cuModuleGetFunction");
    a.moduleUnload = (decltype(a.moduleUnload))sym(cu, "cuModuleUnload");
    a.launchKernel = (decltype(a.launchKernel))sym(cu, "cuLaunchKernel");
    a.getErrorString = (decltype(a.getErrorString))sym(cu, "cuGetErrorString");
    a.moduleGetGlobal = (decltype(a.moduleGetGlobal))sym(cu, "cuModuleGetGlobal_v2");
    a.memcpyHtoDAsync = (decltype(a.memcpyHtoDAsync))sym(cu, "cuMemcpyHtoDAsync_v2");
    a.funcSetAttribute = (decltype(a.funcSetAttribute))sym(cu, "cuFuncSetAttribute");
    a.cu_ok = all;
    if (!all) a.cu_why = a.why;
  });
  return a;
}

constexpr int JIT_BLOCK = 128;

const char* PREAMBLE = R"(
typedef unsigned int u32; typedef unsigned long long u64;
#define P 2013265921u
#define NB 1073741848u   /* Montgomery form of -11 (Fp4 = Fp[x]/(x^4+11)) */
// Two-input additions / subtractions are written min(a +- b, ONES) with ONES = 0xffffffff read from the constant bank (opaque
// to the compiler): ptxas emits ONE VIADDMNMX (ALU pipe) for it and cannot choose IMAD.IADD, which would land on the
// multiplier pipe this kernel is bound by (see poseidon2.cuh).  ZKB_EC_ALU_ADDS=0 restores plain additions.
__constant__ u32 zkb_ones = 0xffffffffu;
#if ZKB_ALU_ADDS
#define ADD2(a, b) min((a) + (b), zkb_ones)
#define SUB2(a, b) min((a) - (b), zkb_ones)
#else
#define ADD2(a, b) ((a) + (b))
#define SUB2(a, b) ((a) - (b))
#endif
__device__ __forceinline__ u32 red(u32 x) { return min(x, x - P); }
__device__ __forceinline__ u32 mul(u32 a, u32 b) { u64 t = (u64)a * b; u32 m = (u32)t * 0x88000001u; u32 r = SUB2((u32)(t >> 32), __umulhi(m, P)); return min(r, r + P); }
__device__ __forceinline__ u32 add(u32 a, u32 b) { return red(ADD2(a, b)); }
__device__ __forceinline__ u32 sub(u32 a, u32 b) { u32 d = SUB2(a, b); return min(d, d + P); }
// Lazy accumulation of sum_k pw_k * f_k (pw_k, f_k canonical): a 64-bit accumulator per Fp4 component takes one
// IMAD.WIDE per term; after every second term its high word is brought back below P (one VIADDMNMX), which keeps the
// accumulator below P * 2^32 + 2 P^2 < 2^64; a single Montgomery reduction at the end of the chain gives the canonical
// word.  An accumulator that continues from a canonical value m starts as m << 32 (= m * R).
__device__ __forceinline__ u64 wide(u32 f, u32 w) { return (u64)f * w; }
__device__ __forceinline__ u64 widem(u32 m, u32 f, u32 w) { return ((u64)m << 32) + (u64)f * w; }
__device__ __forceinline__ void wacc(u64& a, u32 f, u32 w) { a += (u64)f * w; }
__device__ __forceinline__ u64 fixhi(u64 a) { u32 hi = (u32)(a >> 32); hi = min(hi, hi - P); return ((u64)hi << 32) | (u32)a; }
__device__ __forceinline__ u32 fin(u64 a) { a = fixhi(a); u32 m = (u32)a * 0x88000001u; u32 r = SUB2((u32)(a >> 32), __umulhi(m, P)); return min(r, r + P); }
__device__ __forceinline__ void st(u32* p, u32 v) { *p = v; }
)";

// Fp4 arithmetic over a component type T (u32 here; the helpers are generic); the second operand of mul4 is a power of poly_mix
const char* PREAMBLE_F4 = R"(
template <class T> struct F4T { T a, b, c, d; };
typedef F4T<u32> F4;
template <class T> __device__ __forceinline__ F4T<T> mul4(F4T<T> x, F4 y) {
  F4T<T> r;
  r.a = add(mul(x.a, y.a), mul(add(add(mul(x.b, y.d), mul(x.c, y.c)), mul(x.d, y.b)), NB));
  r.b = add(add(mul(x.a, y.b), mul(x.b, y.a)), mul(add(mul(x.c, y.d), mul(x.d, y.c)), NB));
  r.c = add(add(add(mul(x.a, y.c), mul(x.b, y.b)), mul(x.c, y.a)), mul(mul(x.d, y.d), NB));
  r.d = add(add(add(mul(x.a, y.d), mul(x.b, y.c)), mul(x.c, y.b)), mul(x.d, y.a));
  return r;
}
__device__ __forceinline__ F4 tof4(uint4 w) { F4 r; r.a = w.x; r.b = w.y; r.c = w.z; r.d = w.w; return r; }
template <class T, class S> __device__ __forceinline__ F4T<T> scale4(F4T<T> x, S s) { F4T<T> r; r.a = mul(x.a, s); r.b = mul(x.b, s); r.c = mul(x.c, s); r.d = mul(x.d, s); return r; }
template <class T> __device__ __forceinline__ F4T<T> add4(F4T<T> x, F4T<T> y) { F4T<T> r; r.a = add(x.a, y.a); r.b = add(x.b, y.b); r.c = add(x.c, y.c); r.d = add(x.d, y.d); return r; }
)";


uint64_t fnv1a(const std::string& s) {
  uint64_t h = 1469598103934665603ull;
  for (unsigned char c : s) { h ^= c; h *= 1099511628211ull; }
  return h;
}

}  // namespace

struct EvalJitKernel {
  CUmodule mod = nullptr;
  CUfunction fn = nullptr;
  uint32_t n_powers = 1;
  int block = 128;                // threads per CTA (one domain point per thread)
  size_t smem = 0;                // dynamic shared memory per CTA
  CUdeviceptr cdata = 0;        // __constant__ zkb_cd (per-proof powers + globals) when the circuit's data fits in 64 KB
  size_t cdata_bytes = 0;
};
struct EvalJitCache {
  std::map<uint64_t, EvalJitKernel> kernels;     // keyed by hash of the circuit blob content
  std::map<uint64_t, bool> failed;
};

static int min_blocks();
// Variants of the register form that were measured on B200 and removed again (SYN-280, 2^22 points; profiles/r1_l_ec_variants.txt):
// 2 / 4 rows per thread with 64 / 128-bit tap loads and the `back` tap taken from the neighbouring lane by shuffle: 2.6-5.0 ms (the
// 64-bit accumulators of 4 rows spill); per-thread row pointers + uniform column offsets: 2.7 ms; this form: 2.4-2.5 ms.
static uint32_t ec_batch() { const char* e = getenv("ZKB_EC_BATCH"); int v = e ? atoi(e) : 4; return (uint32_t)(v < 1 ? 1 : v > 4096 ? 4096 : v); }
static uint32_t ec_prefetch() { const char* e = getenv("ZKB_EC_PREFETCH"); int v = e ? atoi(e) : 1; return (uint32_t)(v < 0 ? 0 : v > 8 ? 8 : v); }
// Per-proof kernel data: [powers of poly_mix, 4 words each][mix globals][out globals].  It lives in the module's
// __constant__ bank when it fits (operands then come straight from the constant cache), else behind a pointer.
constexpr size_t CONST_WORDS_MAX = 15 * 1024;
static bool const_mode(const CircuitDef& c, uint32_t n_powers) { return 4 * (size_t)n_powers + c.mix_size + c.out_size <= CONST_WORDS_MAX; }

// Straight-line source for the circuit (one domain point per thread); gi.n_powers = number of poly_mix powers the kernel reads.
// Liveness and mix-power bookkeeping shared by the generators.
struct Analysis {
  std::vector<char> fp_used, mx_used;
  std::vector<uint32_t> fp_of, mx_of;             // step index -> fp / mix value id
  std::vector<uint32_t> eqz_uses, other_uses;     // how each live mix value is consumed: as the base of a following AndEqz (chainable), or otherwise
  std::vector<uint32_t> mx_pow;                   // power of poly_mix a mix value has reached
  uint32_t n_powers = 1;
};
static Analysis analyse(const CircuitDef& c) {
  Analysis A;
  const size_t n = c.steps.size();
  A.fp_used.assign(c.n_fp_vars, 0); A.mx_used.assign(c.n_mix_vars, 0);
  A.fp_of.assign(n, 0); A.mx_of.assign(n, 0);
  { uint32_t fi = 0, mi = 0; for (size_t i = 0; i < n; ++i) { if (c.steps[i].op <= PX_MUL) A.fp_of[i] = fi++; else A.mx_of[i] = mi++; } }
  A.mx_used[c.ret] = 1;
  for (size_t i = n; i-- > 0;) {      // liveness from the returned mix value backwards
    const StepDef& s = c.steps[i];
    switch (s.op) {
      case PX_ADD: case PX_SUB: case PX_MUL: if (A.fp_used[A.fp_of[i]]) A.fp_used[s.a] = A.fp_used[s.b] = 1; break;
      case PX_AND_EQZ: if (A.mx_used[A.mx_of[i]]) { A.mx_used[s.a] = 1; A.fp_used[s.b] = 1; } break;
      case PX_AND_COND: if (A.mx_used[A.mx_of[i]]) { A.mx_used[s.a] = A.mx_used[s.c] = 1; A.fp_used[s.b] = 1; } break;
      default: break;
    }
  }
  A.eqz_uses.assign(c.n_mix_vars, 0); A.other_uses.assign(c.n_mix_vars, 0); A.mx_pow.assign(c.n_mix_vars, 0);
  uint32_t mi = 0;
  for (size_t i = 0; i < n; ++i) {
    const StepDef& s = c.steps[i];
    if (s.op <= PX_MUL) continue;
    uint32_t id = mi++;
    if (s.op == PX_AND_EQZ) { A.mx_pow[id] = A.mx_pow[s.a] + 1; if (A.mx_used[id]) { ++A.eqz_uses[s.a]; A.n_powers = std::max(A.n_powers, A.mx_pow[s.a] + 1); } }
    else if (s.op == PX_AND_COND) { A.mx_pow[id] = A.mx_pow[s.a] + A.mx_pow[s.c]; if (A.mx_used[id]) { ++A.other_uses[s.a]; ++A.other_uses[s.c]; A.n_powers = std::max(A.n_powers, A.mx_pow[s.a] + 1); } }
  }
  ++A.other_uses[c.ret];
  return A;
}

// Shared-memory staged form (mode "staged"): a CTA owns SG_BLOCK (512) consecutive domain points.  The columns its constraints
// read are brought into shared memory by bulk asynchronous copies (cp.async.bulk, completion on an mbarrier) issued by one
// elected thread: columns read many times (the code group's constants) stay resident, the others stream through a ring of
// `stages` slots of `cps` columns, refilled `stages` blocks ahead of their use.  Taps then are LDS with an immediate offset
// (row - 4 back is the same column slot, a few words earlier: the slot holds `halo` rows in front of the tile), the loads
// in flight per SM are decoupled from the register file (ring bytes instead of registers), and no address arithmetic runs
// on the multiplier pipe.
// threads (= domain points) per staged CTA; ZKB_EC_BLOCK in {128, 256, 512, 1024}.  Measured (SYN-280, 2^22 points): 128: 2.25 ms, 256: 1.79 ms,
// 512: 1.60 ms, 1024: 1.70 ms -- longer contiguous copies per column (2 KB) use the HBM better, at the same 2048 threads per SM.
static int sg_block() { const char* e = getenv("ZKB_EC_BLOCK"); int v = e ? atoi(e) : 512; return v == 128 || v == 256 || v == 1024 ? v : 512; }
#define SG_BLOCK sg_block()
struct GenInfo {
  uint32_t n_powers = 1;
  int block = JIT_BLOCK;     // threads per CTA
  size_t smem = 0;           // dynamic shared memory (staged form)
  bool staged = false;
};
static bool ec_staged() { const char* e = getenv("ZKB_EC_STAGED"); return !e || atoi(e) != 0; }
static uint32_t env_u32(const char* name, uint32_t def, uint32_t lo, uint32_t hi) { const char* e = getenv(name); long v = e ? atol(e) : (long)def; return (uint32_t)(v < (long)lo ? lo : v > (long)hi ? hi : v); }

const char* PREAMBLE_STAGED = R"(
__device__ __forceinline__ u32 saddr(const void* p) { return (u32)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(u64* bar, u32 count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(saddr(bar)), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(u64* bar, u32 bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(saddr(bar)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(u64* bar, u32 parity) {
  u32 ok;
  do { asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(saddr(bar)), "r"(parity) : "memory"); } while (!ok);
}
__device__ __forceinline__ void bulk_g2s(u32* dst, const u32* src, u32 bytes, u64* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" :: "r"(saddr(dst)), "l"(src), "r"(bytes), "r"(saddr(bar)) : "memory");
}
// rows [c0 - HALO, c0 + BLOCK) of one column into a slot; the first tile wraps around the end of the domain
__device__ __forceinline__ void copy_col(u32* dst, const u32* col, u32 c0, u32 mask, u64* bar) {
  if (HALO == 0) bulk_g2s(dst, col + c0, BLOCK * 4u, bar);
  else if (c0 >= HALO) bulk_g2s(dst, col + (c0 - HALO), ROWP * 4u, bar);
  else { bulk_g2s(dst, col + ((c0 - HALO) & mask), HALO * 4u, bar); bulk_g2s(dst + HALO, col + c0, BLOCK * 4u, bar); }
}
)";

static std::string generate(const CircuitDef& c, GenInfo& gi, bool staged) {
  const size_t n = c.steps.size();
  const Analysis A = analyse(c);
  uint32_t& n_powers = gi.n_powers;
  gi.block = JIT_BLOCK; gi.smem = 0; gi.staged = false;
  const std::vector<char>&fp_used = A.fp_used, &mx_used = A.mx_used;
  const std::vector<uint32_t>&fp_of = A.fp_of, &eqz_uses = A.eqz_uses, &other_uses = A.other_uses, &mx_pow = A.mx_pow;
  n_powers = A.n_powers;
  const bool cm = const_mode(c, n_powers);
  const size_t gl_off = 4 * (size_t)n_powers;
  // ---- staged form: which columns are resident, which stream, and in which block / slot every tap is found ----------
  struct ColUse { uint32_t group, column, uses; };
  uint32_t halo = 0, n_res = 0, cps = 0, stages = 0, rowp = 0;
  std::map<std::pair<uint32_t, uint32_t>, uint32_t> res_slot;               // resident (group, column) -> slot
  std::vector<std::vector<std::pair<uint32_t, uint32_t>>> blocks;           // streamed columns of every block, slot order
  std::vector<uint32_t> get_block(n, 0), get_slot(n, 0);
  if (staged) {
    std::map<std::pair<uint32_t, uint32_t>, uint32_t> uses;
    for (size_t i = 0; i < n; ++i) if (c.steps[i].op == PX_GET && fp_used[fp_of[i]]) {
      const TapDef& t = c.taps[c.steps[i].a];
      ++uses[{t.group, t.column}];
      halo = std::max(halo, 4 * t.back);
    }
    const uint32_t res_max = env_u32("ZKB_EC_RES", 32, 0, 64), res_min_uses = env_u32("ZKB_EC_RES_USES", 4, 2, 1u << 30);
    // measured on B200, SYN-280 (profiles/r1_o_ec_staged.txt), 256-point CTAs: 2 stages x 6 columns = 1.82 ms (28 KB per CTA: eight CTAs = 2048
    // threads per SM), 3 x 6: 1.89-1.96 ms; deeper or wider rings cost occupancy (6 x 8: 3.25 ms, 10 x 8: 4.35 ms); the register form takes
    // 2.4-2.5 ms.  With 512-point CTAs (the default): 2 x 6 = 1.60 ms, 2 x 4 = 1.65 ms, 2 x 8 = 1.86 ms, 3 x 6 = 1.99 ms.
    cps = env_u32("ZKB_EC_CPS", 6, 1, 32); stages = env_u32("ZKB_EC_STAGES", 2, 2, 16);
    std::vector<ColUse> cand;
    for (auto& kv : uses) if (kv.second >= res_min_uses) cand.push_back({kv.first.first, kv.first.second, kv.second});
    std::stable_sort(cand.begin(), cand.end(), [](const ColUse& x, const ColUse& y) { return x.uses > y.uses; });
    for (auto& cu : cand) { if (n_res >= res_max) break; res_slot[{cu.group, cu.column}] = n_res++; }
    rowp = SG_BLOCK + halo;
    std::map<std::pair<uint32_t, uint32_t>, uint32_t> cur;                  // streamed column -> slot in the block being formed
    blocks.emplace_back();
    for (size_t i = 0; i < n; ++i) if (c.steps[i].op == PX_GET && fp_used[fp_of[i]]) {
      const TapDef& t = c.taps[c.steps[i].a];
      std::pair<uint32_t, uint32_t> key{t.group, t.column};
      if (res_slot.count(key)) { get_block[i] = (uint32_t)blocks.size() - 1; continue; }
      auto it = cur.find(key);
      if (it == cur.end()) {
        if (cur.size() == cps) { blocks.emplace_back(); cur.clear(); }
        it = cur.emplace(key, (uint32_t)cur.size()).first;
        blocks.back().push_back(key);
      }
      get_block[i] = (uint32_t)blocks.size() - 1; get_slot[i] = it->second;
    }
    if (blocks.back().empty() && blocks.size() > 1) blocks.pop_back();
    stages = std::min<uint32_t>(stages, (uint32_t)std::max<size_t>(blocks.size(), 1));
    gi.smem = ((size_t)n_res + (size_t)stages * cps) * rowp * 4 + (stages + 1) * 8;
    if (halo > SG_BLOCK || gi.smem > 200 * 1024) return std::string();      // does not fit this form: the caller uses the register form
    gi.block = SG_BLOCK; gi.staged = true;
  }
  std::ostringstream o;
  o << "#define ZKB_ALU_ADDS " << (env_u32("ZKB_EC_ALU_ADDS", 1, 0, 1) ? 1 : 0) << "\n" << PREAMBLE;
  if (staged) o << "#define HALO " << halo << "u\n#define BLOCK " << SG_BLOCK << "u\n#define ROWP " << rowp << "u\n" << PREAMBLE_STAGED;
  o << "typedef u32 RV; typedef u64 ACC;\n#define W0(f, w) wide(f, w)\n#define WM(m, f, w) widem(m, f, w)\n#define LD(p) __ldg(p)\n";
  o << PREAMBLE_F4 << "typedef F4T<RV> MV;\n";
  if (cm) {
    o << "__constant__ u32 zkb_cd[" << std::max<size_t>(gl_off + c.mix_size + c.out_size, 4) << "];\n"
         "#define PW(k) make_uint4(zkb_cd[4 * (k)], zkb_cd[4 * (k) + 1], zkb_cd[4 * (k) + 2], zkb_cd[4 * (k) + 3])\n"
         "#define GL(i) zkb_cd[" << gl_off << " + (i)]\n";
  } else {
    o << "#define PW(k) __ldg(pw + (k))\n#define GL(i) __ldg(gl + (i))\n";
  }
  const int ctas_per_sm = staged ? (int)std::max<size_t>(1, std::min<size_t>(env_u32("ZKB_EC_MINBLOCKS", 8, 1, 16), (227 * 1024) / (gi.smem + 1024))) : min_blocks();
  o << "extern \"C\" __global__ void __launch_bounds__(" << gi.block << ", " << ctas_per_sm << ") zkb_ec(u32* __restrict__ check, const u32* __restrict__ g0, const u32* __restrict__ g1, "
       "const u32* __restrict__ g2, const uint4* __restrict__ pw, const u32* __restrict__ gl, uint4 invden, u32 mask) {\n"
       << (staged ? "  size_t dom; asm(\"add.u64 %0, %1, 1;\" : \"=l\"(dom) : \"l\"((u64)mask));     // opaque: see the note on uniform address arithmetic below\n"
                  : "  const size_t dom = (size_t)mask + 1;\n")
       << "  const u32 c = blockIdx.x * " << gi.block << "u + threadIdx.x;\n";
  // producer code of one block: expect the bytes, then one bulk copy per column (issued by thread 0 only)
  auto issue_block = [&](uint32_t b) {
    const uint32_t st = b % stages;
    o << "    mbar_expect_tx(full + " << st << ", " << blocks[b].size() * rowp * 4 << "u);\n";
    for (size_t k = 0; k < blocks[b].size(); ++k)
      o << "    copy_col(ring + " << ((size_t)st * cps + k) * rowp << "u, g" << blocks[b][k].first << " + (size_t)" << blocks[b][k].second << " * dom, c0, mask, full + " << st << ");\n";
  };
  uint32_t cur_block = 0;
  if (staged) {
    o << "  extern __shared__ __align__(128) u32 zkb_sm[];\n"
         "  u32* const ring = zkb_sm + " << (size_t)n_res * rowp << "u;\n"
         "  u64* const full = reinterpret_cast<u64*>(zkb_sm + " << ((size_t)n_res + (size_t)stages * cps) * rowp << "u);     // [stages] ring barriers + 1 for the resident columns\n"
         "  const u32 c0 = blockIdx.x * BLOCK;\n"
         "  const u32* const sp = zkb_sm + threadIdx.x + HALO;          // this thread's row inside every column slot\n"
         "  if (threadIdx.x == 0) {\n    for (u32 i = 0; i <= " << stages << "u; ++i) mbar_init(full + i, 1u);\n"
         "    asm volatile(\"fence.mbarrier_init.release.cluster;\" ::: \"memory\");\n  }\n  __syncthreads();\n"
         "  if (threadIdx.x == 0) {\n";
    if (n_res) {
      o << "    mbar_expect_tx(full + " << stages << ", " << (size_t)n_res * rowp * 4 << "u);\n";
      for (auto& kv : res_slot) o << "    copy_col(zkb_sm + " << (size_t)kv.second * rowp << "u, g" << kv.first.first << " + (size_t)" << kv.first.second << " * dom, c0, mask, full + " << stages << ");\n";
    }
    for (uint32_t b = 0; b < stages && b < blocks.size(); ++b) if (!blocks[b].empty()) issue_block(b);
    o << "  }\n";
    if (n_res) o << "  mbar_wait(full + " << stages << ", 0u);\n";
    if (!blocks[0].empty()) o << "  mbar_wait(full + 0, 0u);\n";
  }
  // moves the kernel from block `cur_block` to block b: everybody has finished reading the old slot before it is refilled
  auto advance_to = [&](uint32_t b) {
    while (cur_block < b) {
      if (cur_block + stages < blocks.size()) {
        o << "  __syncthreads();\n  if (threadIdx.x == 0) {\n    asm volatile(\"fence.proxy.async.shared::cta;\" ::: \"memory\");\n";
        issue_block(cur_block + stages);
        o << "  }\n";
      }
      ++cur_block;
      o << "  mbar_wait(full + " << cur_block % stages << ", " << ((cur_block / stages) & 1) << "u);\n";
    }
  };
  auto staged_get = [&](size_t i) {
    const TapDef& t = c.taps[c.steps[i].a];
    advance_to(get_block[i]);
    auto r = res_slot.find({t.group, t.column});
    std::ostringstream a;
    const long slot_words = r != res_slot.end() ? (long)r->second * rowp : (long)((size_t)n_res + (size_t)(get_block[i] % stages) * cps + get_slot[i]) * rowp;
    a << "sp[" << slot_words - 4 * (long)t.back << "]";
    return a.str();
  };
  auto tap_addr = [&](const TapDef& t) {
    std::ostringstream a;
    a << "g" << t.group << " + (size_t)" << t.column << " * dom + ((c - " << 4 * t.back << "u) & mask)";
    return a.str();
  };
  enum { ST_ZERO = 0, ST_CANON = 1, ST_ACC = 2 };
  std::vector<char> state(c.n_mix_vars, ST_ZERO);
  std::vector<uint32_t> acc_set(c.n_mix_vars, 0), acc_terms(c.n_mix_vars, 0);
  auto materialize = [&](uint32_t id) {      // accumulators of `id` -> canonical Fp4 m<id>
    uint32_t a = acc_set[id];
    o << "  MV m" << id << "; m" << id << ".a = fin(A" << a << "_0); m" << id << ".b = fin(A" << a << "_1); m" << id << ".c = fin(A" << a << "_2); m" << id << ".d = fin(A" << a << "_3);\n";
    state[id] = ST_CANON;
  };
  // Load scheduling.  The kernel is bound by global-load latency (ncu: 87 % long-scoreboard stalls when every tap is loaded
  // right before its use), so the tap loads of the next `batch` constraints are hoisted in front of the arithmetic of the
  // current ones (software pipelining, distance `prefetch` batches); a tap needed twice within the hoisted group is loaded once.
  const uint32_t batch = ec_batch(), prefetch = ec_prefetch();
  std::vector<uint32_t> get_batch(n, 0);
  uint32_t n_batches = 1;
  {
    uint32_t live_mix = 0, mj = 0;
    for (size_t i = 0; i < n; ++i) {
      const StepDef& s = c.steps[i];
      if (s.op <= PX_MUL) { get_batch[i] = live_mix / batch; continue; }
      uint32_t id = mj++;
      if (s.op != PX_TRUE && mx_used[id]) ++live_mix;
    }
    n_batches = live_mix / batch + 1;
  }
  std::vector<std::vector<size_t>> gets_of(n_batches + 1);
  if (!staged) for (size_t i = 0; i < n; ++i) if (c.steps[i].op == PX_GET && fp_used[fp_of[i]]) gets_of[std::min<uint32_t>(get_batch[i], n_batches)].push_back(i);
  std::vector<char> get_done(n, 0);
  auto hoist = [&](uint32_t b) {
    if (b >= gets_of.size()) return;
    std::map<uint32_t, uint32_t> seen;       // tap -> fp id, within this group
    for (size_t i : gets_of[b]) {
      const StepDef& s = c.steps[i];
      auto it = seen.find(s.a);
      if (it != seen.end()) { o << "  const RV f" << fp_of[i] << " = f" << it->second << ";\n"; }
      else { o << "  const RV f" << fp_of[i] << " = LD(" << tap_addr(c.taps[s.a]) << ");\n"; seen[s.a] = fp_of[i]; }
      get_done[i] = 1;
    }
  };
  uint32_t hoisted_upto = 0;                  // batches [0, hoisted_upto) have had their loads emitted
  auto hoist_until = [&](uint32_t b_end) { while (hoisted_upto < b_end && hoisted_upto < gets_of.size()) hoist(hoisted_upto++); };
  hoist_until(1 + prefetch);
  uint32_t live_seen = 0;
  uint32_t fi = 0, mi = 0;
  for (size_t i = 0; i < n; ++i) {
    const StepDef& s = c.steps[i];
    if (s.op <= PX_MUL) {
      uint32_t id = fi++;
      if (!fp_used[id]) continue;
      if (s.op == PX_GET && get_done[i]) continue;
      std::string get_expr;
      if (s.op == PX_GET) get_expr = staged ? staged_get(i) : "LD(" + tap_addr(c.taps[s.a]) + ")";      // (staged_get may first emit a block change)
      o << "  const auto f" << id << " = ";
      switch (s.op) {
        case PX_CONST: o << Fp::from(s.a).v << "u"; break;
        case PX_GET: o << get_expr; break;
        case PX_GET_GLOBAL: o << "GL(" << (s.a == 0 ? s.b : c.mix_size + s.b) << ")"; break;
        case PX_ADD: o << "add(f" << s.a << ", f" << s.b << ")"; break;
        case PX_SUB: o << "sub(f" << s.a << ", f" << s.b << ")"; break;
        case PX_MUL: o << "mul(f" << s.a << ", f" << s.b << ")"; break;
      }
      o << ";\n";
      continue;
    }
    uint32_t id = mi++;
    if (s.op == PX_TRUE) { state[id] = ST_ZERO; continue; }
    if (!mx_used[id]) continue;
    ++live_seen;
    if (live_seen % batch == 0) hoist_until(live_seen / batch + 1 + prefetch);      // entering the next batch: start the loads `prefetch` batches ahead
    if (s.op == PX_AND_EQZ) {
      const uint32_t base = s.a, k = mx_pow[base];
      uint32_t set, terms;
      if (state[base] == ST_ACC && eqz_uses[base] == 1 && other_uses[base] == 0) {          // continue the chain in place
        set = acc_set[base]; terms = acc_terms[base];
        o << "  { const uint4 w = PW(" << k << "); wacc(A" << set << "_0, f" << s.b << ", w.x); wacc(A" << set << "_1, f" << s.b << ", w.y); wacc(A" << set
          << "_2, f" << s.b << ", w.z); wacc(A" << set << "_3, f" << s.b << ", w.w); }\n";
        ++terms;
      } else {
        set = id;
        if (state[base] == ST_ACC) materialize(base);       // (cannot happen: multi-use values are materialised at definition)
        if (state[base] == ST_ZERO) {
          o << "  ACC A" << set << "_0, A" << set << "_1, A" << set << "_2, A" << set << "_3; { const uint4 w = PW(" << k << "); A" << set << "_0 = W0(f" << s.b
            << ", w.x); A" << set << "_1 = W0(f" << s.b << ", w.y); A" << set << "_2 = W0(f" << s.b << ", w.z); A" << set << "_3 = W0(f" << s.b << ", w.w); }\n";
        } else {
          o << "  ACC A" << set << "_0, A" << set << "_1, A" << set << "_2, A" << set << "_3; { const uint4 w = PW(" << k << "); A" << set << "_0 = WM(m" << base
            << ".a, f" << s.b << ", w.x); A" << set << "_1 = WM(m" << base << ".b, f" << s.b << ", w.y); A" << set << "_2 = WM(m" << base
            << ".c, f" << s.b << ", w.z); A" << set << "_3 = WM(m" << base << ".d, f" << s.b << ", w.w); }\n";
        }
        terms = 1;
      }
      if (terms == 2) {
        o << "  A" << set << "_0 = fixhi(A" << set << "_0); A" << set << "_1 = fixhi(A" << set << "_1); A" << set << "_2 = fixhi(A" << set << "_2); A" << set << "_3 = fixhi(A" << set << "_3);\n";
        terms = 0;
      }
      state[id] = ST_ACC; acc_set[id] = set; acc_terms[id] = terms;
      if (!(eqz_uses[id] == 1 && other_uses[id] == 0)) materialize(id);
      continue;
    }
    // PX_AND_COND: m_id = m_a + f_b * (m_c (x) poly_mix^pow(a)); operands are canonical (materialised at definition)
    if (state[s.c] == ST_ZERO) {           // inner chain is empty: nothing is added
      if (state[s.a] == ST_ZERO) state[id] = ST_ZERO;
      else { o << "  const MV m" << id << " = m" << s.a << ";\n"; state[id] = ST_CANON; }
      continue;
    }
    o << "  const MV m" << id << " = ";
    if (state[s.a] == ST_ZERO) o << "scale4(mul4(m" << s.c << ", tof4(PW(" << mx_pow[s.a] << "))), f" << s.b << ")";
    else o << "add4(m" << s.a << ", scale4(mul4(m" << s.c << ", tof4(PW(" << mx_pow[s.a] << "))), f" << s.b << "))";
    o << ";\n";
    state[id] = ST_CANON;
  }
  o << "  const u32 den = (c & 3u) == 0 ? invden.x : (c & 3u) == 1 ? invden.y : (c & 3u) == 2 ? invden.z : invden.w;\n";
  if (state[c.ret] == ST_ZERO) o << "  MV r; r.a = r.b = r.c = r.d = sub(den, den);\n";
  else o << "  const MV r = scale4(m" << c.ret << ", den);\n";
  o << "  st(check + c, r.a); st(check + dom + c, r.b); st(check + 2 * dom + c, r.c); st(check + 3 * dom + c, r.d);\n}\n";
  return o.str();
}

// ---- on-disk cubin cache -------------------------------------------------------------------------------------------------
// A cubin found on disk is code that will run inside the prover's CUDA context, next to the private witness, so the cache
// follows the rules of any per-user code cache:
//   * directory: ZKB_CACHE_DIR, else `_jitcache/` next to libzkb200.so when this user can write it (cubins compiled at build time
//     ship with the library), else $XDG_CACHE_HOME/zkb200 or ~/.cache/zkb200 created 0700.  Never a shared /tmp path; with no
//     usable directory the cache is memory-only.
//   * a directory or file is trusted only if it is owned by this uid (or root) and not writable by group / others;
//   * every file carries a header (magic, key = fnv1a64(source), a second hash over source + options + NVRTC version, payload size,
//     payload checksum), so stale or torn files are cache misses, not load errors;
//   * files are written to a mkstemp() name and renamed; compilation is serialised on a process-wide mutex (bench.py runs three
//     prover threads per process); a cubin that fails to LOAD is deleted and recompiled once.
static const char* const NVRTC_OPTS[] = {"--gpu-architecture=sm_100a", "--std=c++17", "-lineinfo", "--restrict"};
static bool trusted(const struct stat& st) { return (st.st_uid == geteuid() || st.st_uid == 0) && (st.st_mode & (S_IWGRP | S_IWOTH)) == 0; }
static bool usable_dir(const std::string& d, bool create) {
  if (create && mkdir(d.c_str(), 0700) != 0 && errno != EEXIST) return false;
  struct stat st;
  if (lstat(d.c_str(), &st) != 0 || !S_ISDIR(st.st_mode)) return false;
  return trusted(st);
}
static std::string cache_dir() {
  const char* e = getenv("ZKB_CACHE_DIR");
  if (e && *e) return usable_dir(e, true) ? std::string(e) : std::string();
  Dl_info info;
  if (dladdr((void*)&cache_dir, &info) && info.dli_fname) {
    std::string so = info.dli_fname;
    size_t slash = so.rfind('/');
    if (slash != std::string::npos) {
      std::string d = so.substr(0, slash) + "/_jitcache";
      if (usable_dir(d, access(so.substr(0, slash).c_str(), W_OK) == 0)) return d;     // read-only installs still use the shipped cubins
    }
  }
  std::string base;
  const char* x = getenv("XDG_CACHE_HOME"); const char* h = getenv("HOME");
  if (x && *x == '/') base = x; else if (h && *h == '/') { base = std::string(h) + "/.cache"; mkdir(base.c_str(), 0700); }
  if (base.empty()) return std::string();
  std::string d = base + "/zkb200";
  return usable_dir(d, true) ? d : std::string();
}
static int min_blocks() { const char* e = getenv("ZKB_EC_MINBLOCKS"); int v = e ? atoi(e) : 8; return v < 1 ? 1 : v > 16 ? 16 : v; }

struct CubinHeader { uint32_t magic; uint32_t version; uint64_t key; uint64_t key2; uint64_t size; uint64_t sum; };
static constexpr uint32_t CUBIN_MAGIC = 0x4a424b5au;      // "ZKBJ"
static uint64_t fnv1a_bytes(const char* p, size_t n, uint64_t h = 1469598103934665603ull) { for (size_t i = 0; i < n; ++i) { h ^= (unsigned char)p[i]; h *= 1099511628211ull; } return h; }
static uint64_t second_key(const std::string& src) {
  int maj = 0, min = 0; if (api().version) api().version(&maj, &min);
  std::string salt = "zkb200-jit-v2|nvrtc " + std::to_string(maj) + "." + std::to_string(min);
  for (const char* o : NVRTC_OPTS) { salt += '|'; salt += o; }
  uint64_t h = fnv1a_bytes(salt.data(), salt.size(), 0x9e3779b97f4a7c15ull);
  return fnv1a_bytes(src.data(), src.size(), h) ^ (uint64_t)src.size();
}
static std::string cubin_path(const std::string& dir, uint64_t key) {
  char name[64]; snprintf(name, sizeof name, "/ec_%016llx_sm100a.zkbj", (unsigned long long)key);
  return dir + name;
}
static bool cache_read(const std::string& path, uint64_t key, uint64_t key2, std::vector<char>& cubin) {
  int fd = open(path.c_str(), O_RDONLY | O_NOFOLLOW | O_CLOEXEC);
  if (fd < 0) return false;
  struct stat st; CubinHeader h;
  bool ok = fstat(fd, &st) == 0 && S_ISREG(st.st_mode) && trusted(st) && (size_t)st.st_size > sizeof h && read(fd, &h, sizeof h) == (ssize_t)sizeof h &&
            h.magic == CUBIN_MAGIC && h.version == 2 && h.key == key && h.key2 == key2 && h.size == (uint64_t)st.st_size - sizeof h;
  if (ok) {
    cubin.resize((size_t)h.size);
    size_t got = 0;
    while (got < cubin.size()) { ssize_t r = read(fd, cubin.data() + got, cubin.size() - got); if (r <= 0) break; got += (size_t)r; }
    ok = got == cubin.size() && fnv1a_bytes(cubin.data(), cubin.size()) == h.sum;
  }
  close(fd);
  if (!ok) cubin.clear();
  return ok;
}
static void cache_write(const std::string& dir, const std::string& path, uint64_t key, uint64_t key2, const std::vector<char>& cubin) {
  std::string tmpl = dir + "/.ec_tmp_XXXXXX";
  std::vector<char> tmp(tmpl.begin(), tmpl.end()); tmp.push_back('\0');
  int fd = mkstemp(tmp.data());          // unique per call (threads, processes, ranks), created 0600
  if (fd < 0) return;
  CubinHeader h{CUBIN_MAGIC, 2, key, key2, (uint64_t)cubin.size(), fnv1a_bytes(cubin.data(), cubin.size())};
  bool ok = write(fd, &h, sizeof h) == (ssize_t)sizeof h && write(fd, cubin.data(), cubin.size()) == (ssize_t)cubin.size();
  ok = (fchmod(fd, 0644) == 0) && ok;
  close(fd);
  if (!ok || rename(tmp.data(), path.c_str()) != 0) unlink(tmp.data());
}

static std::mutex g_compile_mutex;
static bool compile(const std::string& src, std::vector<char>& cubin, std::string& why, bool ignore_disk = false) {
  std::lock_guard<std::mutex> lock(g_compile_mutex);
  Api& a = api();
  const uint64_t key = fnv1a(src), key2 = second_key(src);
  const std::string dir = cache_dir();
  const std::string path = dir.empty() ? std::string() : cubin_path(dir, key);
  if (!path.empty()) {
    if (ignore_disk) unlink(path.c_str());
    else if (cache_read(path, key, key2, cubin)) return true;
  }
  nvrtcProgram prog;
  if (a.createProgram(&prog, src.c_str(), "zkb_eval_check.cu", 0, nullptr, nullptr) != NVRTC_SUCCESS) { why = "nvrtcCreateProgram failed"; return false; }
  nvrtcResult r = a.compileProgram(prog, (int)(sizeof NVRTC_OPTS / sizeof NVRTC_OPTS[0]), NVRTC_OPTS);
  if (r != NVRTC_SUCCESS) {
    size_t ls = 0; a.getLogSize(prog, &ls);
    std::string log(ls, '\0'); if (ls) a.getLog(prog, &log[0]);
    why = "nvrtc compile failed: " + log.substr(0, 600);
    a.destroyProgram(&prog);
    return false;
  }
  size_t sz = 0; a.getCUBINSize(prog, &sz);
  cubin.resize(sz); a.getCUBIN(prog, cubin.data());
  a.destroyProgram(&prog);
  if (!path.empty()) cache_write(dir, path, key, key2, cubin);
  return true;
}

void eval_jit_free(zkb_ctx* ctx) {
  if (!ctx->jit) return;
  EvalJitCache* cache = (EvalJitCache*)ctx->jit;
  for (auto& kv : cache->kernels) if (kv.second.mod) api().moduleUnload(kv.second.mod);
  delete cache;
  ctx->jit = nullptr;
}

// The generated source (for tests / inspection) -- no device needed.
std::string eval_jit_source(const CircuitDef& c) {
  GenInfo gi;
  if (ec_staged()) { std::string src = generate(c, gi, true); if (!src.empty()) return src; }
  return generate(c, gi, false);
}
// Compiles the source with NVRTC without loading it (CPU-only check that the generator emits valid CUDA).
bool eval_jit_compile_only(const CircuitDef& c, std::string& why) {
  Api& a = api();
  if (!a.ok) { why = a.why; return false; }
  std::vector<char> cubin; GenInfo gi;
  // every form a proof may use: the staged kernel, and the register form used for tiny domains / unaligned sub-buffers
  if (ec_staged()) { std::string src = generate(c, gi, true); if (!src.empty() && !compile(src, cubin, why)) return false; }
  return compile(generate(c, gi, false), cubin, why);
}

static const EvalJitKernel* get_kernel(zkb_ctx* ctx, const CircuitDef& c, bool staged, std::string& why) {
  Api& a = api();
  if (!a.ok) { why = a.why; return nullptr; }
  if (!a.cu_ok) { why = a.cu_why; return nullptr; }
  if (!ctx->jit) ctx->jit = new EvalJitCache();
  EvalJitCache* cache = (EvalJitCache*)ctx->jit;
  GenInfo gi;
  std::string src = generate(c, gi, staged);
  if (src.empty()) { why = "circuit does not fit the staged form"; return nullptr; }
  const uint32_t np = gi.n_powers;
  uint64_t key = fnv1a(src);
  auto it = cache->kernels.find(key);
  if (it != cache->kernels.end()) return &it->second;
  if (cache->failed.count(key)) { why = "previous JIT attempt failed"; return nullptr; }
  std::vector<char> cubin;
  EvalJitKernel k; k.n_powers = np; k.block = gi.block; k.smem = gi.smem;
  if (!compile(src, cubin, why)) { cache->failed[key] = true; return nullptr; }
  ZKB_CUDA(cudaFree(0));     // make sure the primary context is current for the driver API
  CUresult r = a.moduleLoadData(&k.mod, cubin.data());
  if (r != CUDA_SUCCESS) {      // a cached cubin that does not load (other driver, damaged file) is a cache miss: drop it, recompile once
    cubin.clear();
    if (!compile(src, cubin, why, /*ignore_disk=*/true)) { cache->failed[key] = true; return nullptr; }
    r = a.moduleLoadData(&k.mod, cubin.data());
  }
  if (r == CUDA_SUCCESS) r = a.moduleGetFunction(&k.fn, k.mod, "zkb_ec");
  if (r == CUDA_SUCCESS && k.smem > 48 * 1024) r = a.funcSetAttribute(k.fn, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, (int)k.smem);
  if (r != CUDA_SUCCESS) { const char* s = nullptr; a.getErrorString(r, &s); why = std::string("cuModuleLoadData: ") + (s ? s : "?"); cache->failed[key] = true; return nullptr; }
  if (const_mode(c, np)) {
    r = a.moduleGetGlobal(&k.cdata, &k.cdata_bytes, k.mod, "zkb_cd");
    if (r != CUDA_SUCCESS) { const char* s = nullptr; a.getErrorString(r, &s); why = std::string("cuModuleGetGlobal(zkb_cd): ") + (s ? s : "?"); cache->failed[key] = true; return nullptr; }
  }
  return &(cache->kernels[key] = k);
}

// Returns false (with `why`) when the JIT path is unavailable; the caller then uses the interpreter.
bool eval_check_jit(zkb_ctx* ctx, uint32_t* d_check, const CircuitDef& c, const uint32_t* const d_groups[3], const uint32_t* mix_g, const uint32_t* out_g,
                    const Fp4& poly_mix, int po2, std::string& why) {
  const size_t n = (size_t)1 << po2, domain = n * INV_RATE;
  if (domain < (size_t)JIT_BLOCK) { why = "domain smaller than one block"; return false; }
  bool aligned = ((uintptr_t)d_check & 15) == 0;
  for (int g = 0; g < 3; ++g) if ((uintptr_t)d_groups[g] & 15) aligned = false;
  const EvalJitKernel* k = nullptr;
  if (ec_staged() && aligned && domain >= (size_t)SG_BLOCK) {       // bulk copies need 16-byte aligned columns (pool allocations are; a caller's sub-buffer view may not be)
    std::string why_staged;
    k = get_kernel(ctx, c, true, why_staged);
  }
  if (!k) k = get_kernel(ctx, c, false, why);
  if (!k) return false;
  // per-proof data: [powers of poly_mix (4 words each)] [mix globals] [out globals]
  std::vector<uint32_t> h(4 * (size_t)k->n_powers + c.mix_size + c.out_size + 4);
  Fp4 cur = Fp4::one();
  for (uint32_t i = 0; i < k->n_powers; ++i) { cur.store(&h[4 * i]); cur *= poly_mix; }
  uint32_t* gl = h.data() + 4 * (size_t)k->n_powers;
  for (uint32_t i = 0; i < c.mix_size; ++i) gl[i] = mix_g[i];
  for (uint32_t i = 0; i < c.out_size; ++i) gl[c.mix_size + i] = out_g[i];
  uint32_t* d_data = nullptr;
  const size_t data_words = 4 * (size_t)k->n_powers + c.mix_size + c.out_size;
  if (k->cdata) {
    ZKB_REQUIRE(data_words * 4 <= k->cdata_bytes, "eval_check JIT: constant bank smaller than the per-proof data");
    CUresult cr = api().memcpyHtoDAsync(k->cdata, h.data(), data_words * 4, (CUstream)ctx->stream);    // pageable source: staged before the call returns
    if (cr != CUDA_SUCCESS) { const char* s = nullptr; api().getErrorString(cr, &s); throw Error(std::string("zkb200: eval_check JIT constant upload failed: ") + (s ? s : "?")); }
  } else {
    pool_alloc(ctx, &d_data, h.size() * 4);
    ZKB_CUDA(cudaMemcpyAsync(d_data, h.data(), h.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
  }
  Fp w4 = pow(Fp::from(137), (uint64_t)1 << (MAX_ROU_PO2 - 2));
  Fp three_n = pow(Fp::from(3), n);
  uint4 invden; uint32_t* idp = &invden.x;
  Fp wr = Fp::one();
  for (int r = 0; r < 4; ++r) { idp[r] = inv(three_n * wr - Fp::one()).v; wr *= w4; }
  const uint32_t* g0 = d_groups[0]; const uint32_t* g1 = d_groups[1]; const uint32_t* g2 = d_groups[2];
  const uint4* pw = (const uint4*)d_data; const uint32_t* d_gl = d_data + 4 * (size_t)k->n_powers;
  uint32_t mask = (uint32_t)(domain - 1);
  void* args[] = {&d_check, &g0, &g1, &g2, &pw, &d_gl, &invden, &mask};
  CUresult r = api().launchKernel(k->fn, (unsigned)(domain / (size_t)k->block), 1, 1, (unsigned)k->block, 1, 1, (unsigned)k->smem, (CUstream)ctx->stream, args, nullptr);
  if (r != CUDA_SUCCESS) { const char* s = nullptr; api().getErrorString(r, &s); throw Error(std::string("zkb200: eval_check JIT launch failed: ") + (s ? s : "?")); }
  launched(ctx);
  if (d_data) pool_free(ctx, d_data);
  return true;
}

}  // namespace zkb
