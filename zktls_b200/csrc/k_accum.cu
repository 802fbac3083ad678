// CircuitHal::accumulate: the circuit's witness program for the accum group, delivered as DATA in the circuit blob.
//
// Stand-in for risc0-circuit-rv32im 1.2.5 `CircuitHal::accumulate(ctrl, io, data, mix, accum, steps)` (un-vendored; the circuit
// crate's `prove/hal/{cpu,cuda}.rs`, called by `prove_segment` between the data commit and the accum commit -- in-tree call site
// /root/reference/crates/guest-prover-r0/src/prover.rs:90).  In the reference the per-cycle accumulation step is generated code
// of the circuit (`step_compute_accum` / `step_verify_accum`, run over every cycle with the Fiat-Shamir `mix` values) followed by
// `Hal::prefix_products` for the grand-product columns.  Here -- exactly like eval_check's `poly_ext` -- it arrives as a step
// program in the blob (DESIGN.md "circuit blob": Const / Get(group, column, back) / GetGlobal / Add / Sub / Mul, Set(accum column,
// value, condition), Barrier, PrefixProduct(first of 4 accum columns)), is turned into one straight-line kernel per phase with NVRTC
// (csrc/k_eval_jit.cu: accumulate_jit; cubins share eval_check's cache) and run one thread per row: consecutive threads on
// consecutive rows, every access a full 128-byte line, HBM-bound.  A PrefixProduct phase gathers its four planar columns into Fp4
// elements, runs the chunked device scan of `prefix_products` and scatters them back.  The SYN family is simply the first user
// (its program is written next to its constraints in zktls_b200/circuit.py); nothing here knows a circuit by name.
#include "common.cuh"
#include "circuit.hpp"
#include "ops.cuh"

namespace zkb {

bool accumulate_jit(zkb_ctx* ctx, const CircuitDef& c, size_t phase, uint32_t* d_accum, const uint32_t* d_code, const uint32_t* d_data, const uint32_t* d_gl, uint32_t n, std::string& why);

__global__ void __launch_bounds__(256) k_planar_to_fp4(uint4* __restrict__ out, const uint32_t* __restrict__ col, uint32_t n) {
  const uint32_t i = blockIdx.x * 256u + threadIdx.x;
  if (i < n) out[i] = make_uint4(col[i], col[(size_t)n + i], col[2 * (size_t)n + i], col[3 * (size_t)n + i]);
}
__global__ void __launch_bounds__(256) k_fp4_to_planar(uint32_t* __restrict__ col, const uint4* __restrict__ in, uint32_t n) {
  const uint32_t i = blockIdx.x * 256u + threadIdx.x;
  if (i < n) { const uint4 v = in[i]; col[i] = v.x; col[(size_t)n + i] = v.y; col[2 * (size_t)n + i] = v.z; col[3 * (size_t)n + i] = v.w; }
}

void accumulate(zkb_ctx* ctx, const CircuitDef& c, uint32_t* d_accum, const uint32_t* d_code, const uint32_t* d_data, const uint32_t* h_mix, const uint32_t* h_io, int po2) {
  ZKB_REQUIRE(po2 >= 1 && po2 <= 24, "accumulate: po2 out of range");
  if (c.wsteps.empty()) throw Error("zkb200: accumulate: the circuit blob of '" + std::string((const char*)c.info, 16) + "' carries no witness program");
  if (c.group_size[GROUP_ACCUM] == 0) return;
  const uint32_t n = 1u << po2;
  uint32_t* d_gl = nullptr;
  pool_alloc(ctx, &d_gl, (size_t)std::max<uint32_t>(c.mix_size + c.out_size, 4) * 4);
  if (c.mix_size) ZKB_CUDA(cudaMemcpyAsync(d_gl, h_mix, c.mix_size * 4, cudaMemcpyHostToDevice, ctx->stream));
  if (c.out_size) ZKB_CUDA(cudaMemcpyAsync(d_gl + c.mix_size, h_io, c.out_size * 4, cudaMemcpyHostToDevice, ctx->stream));
  size_t phase = 0, lo = 0;
  for (size_t i = 0; i <= c.wsteps.size(); ++i) {
    const bool end = i == c.wsteps.size();
    if (!end && c.wsteps[i].op != WX_BARRIER && c.wsteps[i].op != WX_PREFIX_PRODUCT) continue;
    if (i > lo) {      // a non-empty phase
      std::string why;
      if (!accumulate_jit(ctx, c, phase, d_accum, d_code, d_data, d_gl, n, why)) { pool_free(ctx, d_gl); throw Error("zkb200: accumulate needs the NVRTC JIT: " + why); }
    }
    ++phase; lo = i + 1;
    if (!end && c.wsteps[i].op == WX_PREFIX_PRODUCT) {
      uint32_t* col = d_accum + (size_t)c.wsteps[i].a * n;
      uint32_t* tmp = nullptr;
      pool_alloc(ctx, &tmp, (size_t)n * 16);
      k_planar_to_fp4<<<grid_for(n, 256), 256, 0, ctx->stream>>>((uint4*)tmp, col, n); launched(ctx);
      prefix_products(ctx, tmp, n);
      k_fp4_to_planar<<<grid_for(n, 256), 256, 0, ctx->stream>>>(col, (const uint4*)tmp, n); launched(ctx);
      pool_free(ctx, tmp);
    }
  }
  pool_free(ctx, d_gl);
}

}  // namespace zkb

using namespace zkb;

extern "C" zkb_err zkb_accumulate(zkb_ctx* ctx, const uint32_t* h_circuit, size_t circuit_words, void* d_accum, const void* d_code, const void* d_data,
                                  const uint32_t* h_mix, const uint32_t* h_io, int po2) {
  ZKB_API_BEGIN
  use(ctx);
  ZKB_REQUIRE(h_circuit, "null argument");
  CircuitDef c = CircuitDef::parse(h_circuit, circuit_words);
  ZKB_REQUIRE((h_mix || c.mix_size == 0) && (h_io || c.out_size == 0), "null globals");
  ZKB_REQUIRE((d_accum || c.group_size[GROUP_ACCUM] == 0) && (d_code || c.group_size[GROUP_CODE] == 0) && (d_data || c.group_size[GROUP_DATA] == 0), "null group buffer");
  accumulate(ctx, c, (uint32_t*)d_accum, (const uint32_t*)d_code, (const uint32_t*)d_data, h_mix, h_io, po2);
  ZKB_API_END
}
