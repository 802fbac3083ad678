// CircuitHal::accumulate for the circuits this library carries a witness program for.
//
// Stand-in for risc0-circuit-rv32im 1.2.5 `CircuitHal::accumulate(ctrl, io, data, mix, accum, steps)` (un-vendored; the circuit
// crate's `prove/hal/{cpu,cuda}.rs`, called by `prove_segment` between the data commit and the accum commit -- in-tree call site
// /root/reference/crates/guest-prover-r0/src/prover.rs:90).  In the reference the per-cycle accumulation step is generated code
// of the circuit (`step_compute_accum` / `step_verify_accum`, run over every cycle with the Fiat-Shamir `mix` values, plus
// `Hal::prefix_products` for the grand products).  The generated rv32im step functions are not obtainable offline (SURVEY.md
// 8c), so -- exactly like eval_check -- the op is implemented for the synthetic SYN family ("SYN<W>:v1" circuit info, defined in
// zktls_b200/circuit.py next to its constraints):
//     even accum column j, live row i:  a_j[i] = m_{j mod M} * d_{j mod D}[i] + d_{(j+1) mod D}[i-1]
//     odd  accum column j, live row i:  a_j[i] = a_{j-1}[i-1] * a_{j-1}[i] + m_{j mod M} + o_{j mod O}
// where a row is live when the selector code[0][i] is one; the other rows keep what the caller put there (the reference fills
// its trailing ZK rows with noise the same way).  One thread per row, consecutive threads on consecutive rows: every access is
// a full 128-byte line; 4 bytes read per (row, source column), 4 written: HBM-bound.
#include "common.cuh"
#include "circuit.hpp"

namespace zkb {

struct AccumArgs { uint32_t accum_cols, data_cols, mix_size, out_size; };

// parity = 0: even columns (functions of data only); parity = 1: odd columns (functions of the finished even columns)
__global__ void __launch_bounds__(256) k_syn_accumulate(uint32_t* __restrict__ accum, const uint32_t* __restrict__ code, const uint32_t* __restrict__ data,
                                                         const uint32_t* __restrict__ mix, const uint32_t* __restrict__ out_g, AccumArgs a, uint32_t n, uint32_t parity) {
  const uint32_t i = blockIdx.x * 256u + threadIdx.x;
  if (i >= n) return;
  if (__ldg(code + i) != R_MOD_P) return;                       // selector column: Montgomery one on live rows
  const uint32_t ip = (i + n - 1u) & (n - 1u);                  // previous row, cyclic
  for (uint32_t j = parity; j < a.accum_cols; j += 2) {
    const uint32_t m = __ldg(mix + j % a.mix_size);
    uint32_t v;
    if (parity == 0) {
      const uint32_t d0 = __ldg(data + (size_t)(j % a.data_cols) * n + i), d1 = __ldg(data + (size_t)((j + 1) % a.data_cols) * n + ip);
      v = add_mod(mont_mul(m, d0), d1);
    } else {
      const uint32_t* prev = accum + (size_t)(j - 1) * n;
      v = add_mod(add_mod(mont_mul(prev[ip], prev[i]), m), __ldg(out_g + j % a.out_size));
    }
    accum[(size_t)j * n + i] = v;
  }
}

static bool is_syn(const CircuitDef& c) { return c.info[0] == 'S' && c.info[1] == 'Y' && c.info[2] == 'N' && memchr(c.info, ':', 16) && !memcmp((const char*)memchr(c.info, ':', 16), ":v1", 3); }

void accumulate(zkb_ctx* ctx, const CircuitDef& c, uint32_t* d_accum, const uint32_t* d_code, const uint32_t* d_data, const uint32_t* h_mix, const uint32_t* h_io, int po2) {
  ZKB_REQUIRE(po2 >= 1 && po2 <= 24, "accumulate: po2 out of range");
  if (!is_syn(c)) throw Error("zkb200: accumulate: no witness program for circuit '" + std::string((const char*)c.info, 16) + "' (built in: the SYN family)");
  ZKB_REQUIRE(c.group_size[GROUP_CODE] >= 1 && c.group_size[GROUP_DATA] >= 1 && c.mix_size >= 1 && c.out_size >= 1, "accumulate: malformed SYN circuit");
  if (c.group_size[GROUP_ACCUM] == 0) return;
  const uint32_t n = 1u << po2;
  uint32_t* d_gl = nullptr;
  pool_alloc(ctx, &d_gl, (size_t)(c.mix_size + c.out_size) * 4);
  ZKB_CUDA(cudaMemcpyAsync(d_gl, h_mix, c.mix_size * 4, cudaMemcpyHostToDevice, ctx->stream));
  ZKB_CUDA(cudaMemcpyAsync(d_gl + c.mix_size, h_io, c.out_size * 4, cudaMemcpyHostToDevice, ctx->stream));
  AccumArgs a{c.group_size[GROUP_ACCUM], c.group_size[GROUP_DATA], c.mix_size, c.out_size};
  for (uint32_t parity = 0; parity < 2 && parity < a.accum_cols; ++parity) {
    k_syn_accumulate<<<grid_for(n, 256), 256, 0, ctx->stream>>>(d_accum, d_code, d_data, d_gl, d_gl + c.mix_size, a, n, parity);
    launched(ctx);
  }
  pool_free(ctx, d_gl);
}

}  // namespace zkb

using namespace zkb;

extern "C" zkb_err zkb_accumulate(zkb_ctx* ctx, const uint32_t* h_circuit, size_t circuit_words, void* d_accum, const void* d_code, const void* d_data,
                                  const uint32_t* h_mix, const uint32_t* h_io, int po2) {
  ZKB_API_BEGIN
  use(ctx);
  ZKB_REQUIRE(h_circuit && h_mix && h_io, "null argument");
  CircuitDef c = CircuitDef::parse(h_circuit, circuit_words);
  ZKB_REQUIRE((d_accum || c.group_size[GROUP_ACCUM] == 0) && d_code && d_data, "null group buffer");
  accumulate(ctx, c, (uint32_t*)d_accum, (const uint32_t*)d_code, (const uint32_t*)d_data, h_mix, h_io, po2);
  ZKB_API_END
}
