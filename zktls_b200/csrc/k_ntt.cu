// Batched BabyBear NTT / iNTT / coset LDE (sm_100a).
//
// Stand-ins for risc0-zkp `Hal::{batch_interpolate_ntt, batch_evaluate_ntt, batch_expand_into_evaluate_ntt}`
// (CudaHal binds sppark_batch_iNTT / sppark_batch_NTT / sppark_batch_expand) with the exact index
// conventions of risc0-zkp `core/ntt.rs` (App. C.1-C.3): the inverse is a DIF transform, natural-order evaluations
// in -> bit-reversed coefficients out, scaled by 1/n; the forward is a DIT transform, bit-reversed coefficients in
// -> natural-order evaluations out, with the first `expand_bits` levels replaced by replication.
//
// This file holds the twiddle tables and the level-at-a-time reference path (`*_levels`), which is used for
// sizes the tiled path does not cover and as the on-device cross-check of the tiled kernels in k_ntt_tiled.cu.
#include "common.cuh"
#include "ntt.cuh"

namespace zkb {

// ---- twiddle tables ----------------------------------------------------------------------------------------
// Universal two-level table over W = w_{2^26}:  W^e = hi[e >> 13] * lo[e & 8191].
NttTables* ntt_tables(zkb_ctx* ctx) {
  if (ctx->ntt) return ctx->ntt;
  NttTables* t = new NttTables();
  std::vector<uint32_t> hi(TW_SPLIT), lo(TW_SPLIT);
  Fp w = pow(Fp::from(137), 2);                 // w_{2^27} = 137  ->  w_{2^26} = 137^2
  Fp w_hi = pow(w, TW_SPLIT);
  Fp a = Fp::one(), b = Fp::one();
  for (uint32_t i = 0; i < TW_SPLIT; ++i) { lo[i] = a.v; hi[i] = b.v; a *= w; b *= w_hi; }
  ZKB_CUDA(cudaMalloc((void**)&t->d_hi, TW_SPLIT * 4));
  ZKB_CUDA(cudaMalloc((void**)&t->d_lo, TW_SPLIT * 4));
  ZKB_CUDA(cudaMemcpy(t->d_hi, hi.data(), TW_SPLIT * 4, cudaMemcpyHostToDevice));
  ZKB_CUDA(cudaMemcpy(t->d_lo, lo.data(), TW_SPLIT * 4, cudaMemcpyHostToDevice));
  for (int i = 0; i <= MAX_ROU_PO2; ++i) {
    t->rou_fwd[i] = pow(Fp::from(137), (uint64_t)1 << (MAX_ROU_PO2 - i));
    t->rou_rev[i] = inv(t->rou_fwd[i]);
  }
  ctx->ntt = t;
  return t;
}
void ntt_tables_free(zkb_ctx* ctx) {
  if (!ctx->ntt) return;
  cudaFree(ctx->ntt->d_hi);
  cudaFree(ctx->ntt->d_lo);
  for (auto& kv : ctx->ntt->level_tables) cudaFree(kv.second);
  for (auto& kv : ctx->ntt->shift_tables) cudaFree(kv.second);
  for (auto& kv : ctx->ntt->s_tables) cudaFree(kv.second);
  delete ctx->ntt;
  ctx->ntt = nullptr;
}

// ---- level-at-a-time reference path --------------------------------------------------------------------
constexpr int LV_BLOCK = 256;
// DIF level q (half = 2^q): (a, b) -> (a + b, (a - b) * w_{2^(q+1)}^(-j));  `scale` multiplies both outputs.
__global__ void k_dif_level(uint32_t* __restrict__ io, int k, int q, size_t total_bfly, TwiddleRef tw, uint32_t scale) {
  size_t g = (size_t)blockIdx.x * LV_BLOCK + threadIdx.x;
  if (g >= total_bfly) return;
  size_t half_n = (size_t)1 << (k - 1);
  size_t col = g >> (k - 1), b = g & (half_n - 1);
  size_t j = b & (((size_t)1 << q) - 1);
  size_t i0 = ((b >> q) << (q + 1)) + j;
  uint32_t* p = io + (col << k);
  uint32_t x = p[i0], y = p[i0 + ((size_t)1 << q)];
  uint32_t s = add_mod(x, y);
  uint32_t d = mont_mul(sub_mod(x, y), tw.inv((uint32_t)j, q + 1));
  if (scale != R_MOD_P) { s = mont_mul(s, scale); d = mont_mul(d, scale); }
  p[i0] = s; p[i0 + ((size_t)1 << q)] = d;
}
// DIT level q: (a, b) -> (a + b * w^j, a - b * w^j)
__global__ void k_dit_level(uint32_t* __restrict__ io, int k, int q, size_t total_bfly, TwiddleRef tw) {
  size_t g = (size_t)blockIdx.x * LV_BLOCK + threadIdx.x;
  if (g >= total_bfly) return;
  size_t half_n = (size_t)1 << (k - 1);
  size_t col = g >> (k - 1), b = g & (half_n - 1);
  size_t j = b & (((size_t)1 << q) - 1);
  size_t i0 = ((b >> q) << (q + 1)) + j;
  uint32_t* p = io + (col << k);
  uint32_t x = p[i0], y = mont_mul(p[i0 + ((size_t)1 << q)], tw.fwd((uint32_t)j, q + 1));
  p[i0] = add_mod(x, y); p[i0 + ((size_t)1 << q)] = sub_mod(x, y);
}
__global__ void k_scale(uint32_t* __restrict__ io, size_t n, uint32_t scale) {
  size_t i = (size_t)blockIdx.x * LV_BLOCK + threadIdx.x;
  if (i < n) io[i] = mont_mul(io[i], scale);
}

void zk_shift(zkb_ctx* ctx, uint32_t* io, size_t count, int po2);                                  // k_poly.cu
void batch_expand(zkb_ctx* ctx, uint32_t* out, const uint32_t* in, size_t count, int in_po2, int expand_bits);

void ntt_inverse_levels(zkb_ctx* ctx, uint32_t* io, size_t count, int k, bool shift) {
  NttTables* t = ntt_tables(ctx);
  if (count == 0) return;
  TwiddleRef tw{t->d_hi, t->d_lo};
  uint32_t scale = inv(Fp::from((uint32_t)1 << k)).v;
  size_t total = count << k;
  if (k == 0) {
    // 1/1 == 1: nothing to do
  } else {
    size_t bfly = total >> 1;
    for (int q = k - 1; q >= 0; --q) {
      k_dif_level<<<grid_for(bfly, LV_BLOCK), LV_BLOCK, 0, ctx->stream>>>(io, k, q, bfly, tw, q == 0 ? scale : R_MOD_P);
      launched(ctx);
    }
  }
  if (shift) zk_shift(ctx, io, count, k);
}
void ntt_forward_levels(zkb_ctx* ctx, uint32_t* io, size_t count, int k, int expand_bits) {
  NttTables* t = ntt_tables(ctx);
  if (count == 0 || k == 0) return;
  TwiddleRef tw{t->d_hi, t->d_lo};
  size_t bfly = (count << k) >> 1;
  for (int q = expand_bits; q < k; ++q) {
    k_dit_level<<<grid_for(bfly, LV_BLOCK), LV_BLOCK, 0, ctx->stream>>>(io, k, q, bfly, tw);
    launched(ctx);
  }
}

// ---- dispatch ---------------------------------------------------------------------------------------------
bool ntt_inverse_tiled(zkb_ctx* ctx, uint32_t* io, size_t count, int k, bool shift, const uint32_t* src);     // k_ntt_tiled.cu
bool ntt_forward_tiled(zkb_ctx* ctx, uint32_t* out, const uint32_t* in, size_t count, int k_out, int expand_bits);

void ntt_inverse(zkb_ctx* ctx, uint32_t* io, size_t count, int k, bool shift, const uint32_t* src) {
  if (ntt_inverse_tiled(ctx, io, count, k, shift, src)) return;
  if (src && src != io && count) ZKB_CUDA(cudaMemcpyAsync(io, src, (count << k) * 4, cudaMemcpyDeviceToDevice, ctx->stream));
  ntt_inverse_levels(ctx, io, count, k, shift);
}
// out: count x 2^k_out; in: count x 2^(k_out - expand_bits) (may alias out only when expand_bits == 0)
void ntt_forward(zkb_ctx* ctx, uint32_t* out, const uint32_t* in, size_t count, int k_out, int expand_bits) {
  if (ntt_forward_tiled(ctx, out, in, count, k_out, expand_bits)) return;
  if (expand_bits > 0 || out != in) {
    if (expand_bits == 0) ZKB_CUDA(cudaMemcpyAsync(out, in, (count << k_out) * 4, cudaMemcpyDeviceToDevice, ctx->stream));
    else batch_expand(ctx, out, in, count, k_out - expand_bits, expand_bits);
  }
  ntt_forward_levels(ctx, out, count, k_out, expand_bits);
}

}  // namespace zkb

using namespace zkb;

extern "C" {

zkb_err zkb_batch_interpolate_ntt(zkb_ctx* ctx, void* d_io, size_t count, int po2) {
  ZKB_API_BEGIN use(ctx);
  ZKB_REQUIRE(po2 >= 0 && po2 <= MAX_PO2, "po2 out of range [0, 26]");
  ZKB_REQUIRE((d_io && aligned16(d_io)) || !count, "null or misaligned buffer");
  ntt_inverse(ctx, (uint32_t*)d_io, count, po2, false, nullptr);
  ZKB_API_END
}
zkb_err zkb_batch_interpolate_ntt_zk_shift(zkb_ctx* ctx, void* d_io, size_t count, int po2) {
  ZKB_API_BEGIN use(ctx);
  ZKB_REQUIRE(po2 >= 0 && po2 <= MAX_PO2, "po2 out of range [0, 26]");
  ZKB_REQUIRE((d_io && aligned16(d_io)) || !count, "null or misaligned buffer");
  ntt_inverse(ctx, (uint32_t*)d_io, count, po2, true, nullptr);
  ZKB_API_END
}
zkb_err zkb_batch_evaluate_ntt(zkb_ctx* ctx, void* d_io, size_t count, int po2, int expand_bits) {
  ZKB_API_BEGIN use(ctx);
  ZKB_REQUIRE(po2 >= 0 && po2 <= MAX_PO2, "po2 out of range [0, 26]");
  ZKB_REQUIRE(expand_bits >= 0 && expand_bits <= po2, "expand_bits out of range");
  ZKB_REQUIRE((d_io && aligned16(d_io)) || !count, "null or misaligned buffer");
  // in-place: the input already holds the replicated values, so only the butterfly levels above expand_bits run
  ntt_forward_levels(ctx, (uint32_t*)d_io, count, po2, expand_bits);
  ZKB_API_END
}
zkb_err zkb_batch_expand_into_evaluate_ntt(zkb_ctx* ctx, void* d_out, const void* d_in, size_t count, int in_po2, int expand_bits) {
  ZKB_API_BEGIN use(ctx);
  ZKB_REQUIRE(in_po2 >= 0 && expand_bits >= 0 && in_po2 + expand_bits <= MAX_PO2, "po2 out of range [0, 26]");
  ZKB_REQUIRE((d_out && d_in && aligned16(d_out) && aligned16(d_in)) || !count, "null or misaligned buffer");
  ZKB_REQUIRE(d_out != d_in || expand_bits == 0, "out must not alias in when expanding");
  ntt_forward(ctx, (uint32_t*)d_out, (const uint32_t*)d_in, count, in_po2 + expand_bits, expand_bits);
  ZKB_API_END
}

}  // extern "C"
