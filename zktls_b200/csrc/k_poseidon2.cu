// Poseidon2 row hashing and Merkle folding kernels (sm_100a).
//
// Stand-ins for risc0-zkp `Hal::hash_rows` / `Hal::hash_fold` (CudaHal binds sppark_poseidon2_rows /
// sppark_poseidon2_fold for them).  INT32-pipe bound: one thread owns one
// sponge, the 24-cell state stays in registers, consecutive threads read consecutive rows of each column, so
// every global load is a fully coalesced 128-byte line.
#include "common.cuh"
#include "poseidon2.cuh"
#include "poseidon2_warp.cuh"

namespace zkb {

constexpr int HASH_BLOCK = 128;
// the all-ones word p2::permute uses to keep additions on the ALU pipe (poseidon2.cuh): a launch argument, so ptxas cannot fold it
static inline uint32_t p2_ones() { return 0xffffffffu; }

// out[r] = unpadded_hash(matrix[0*rows + r], matrix[1*rows + r], ...)   (rate 16, overwrite mode, zero pad)
template <int BLOCK>
__global__ void __launch_bounds__(BLOCK) k_hash_rows(uint32_t* __restrict__ out, const uint32_t* __restrict__ matrix, size_t rows, uint32_t cols, uint32_t ones) {
  size_t r = (size_t)blockIdx.x * BLOCK + threadIdx.x;
  if (r >= rows) return;
  uint32_t s[24];
#pragma unroll
  for (int i = 0; i < 24; ++i) s[i] = 0;
  const uint32_t* p = matrix + r;
  uint32_t c = 0;
  for (; c + 16 <= cols; c += 16) {
#pragma unroll
    for (int i = 0; i < 16; ++i) s[i] = __ldg(p + (size_t)(c + i) * rows);
    p2::permute(s, ZKB_P2_TABLES, ones);
  }
  if (c < cols || cols == 0) {
#pragma unroll
    for (int i = 0; i < 16; ++i) s[i] = (c + i < cols) ? __ldg(p + (size_t)(c + i) * rows) : 0u;
    p2::permute(s, ZKB_P2_TABLES, ones);
  }
  uint4* o = reinterpret_cast<uint4*>(out + r * 8);
  o[0] = make_uint4(s[0], s[1], s[2], s[3]);
  o[1] = make_uint4(s[4], s[5], s[6], s[7]);
}

// nodes[out_base + i] = hash_pair(nodes[in_base + 2i], nodes[in_base + 2i + 1]); digests are 8 words.
__global__ void __launch_bounds__(HASH_BLOCK) k_hash_fold(uint32_t* __restrict__ nodes, size_t in_base, size_t out_base, size_t count, uint32_t ones) {
  size_t i = (size_t)blockIdx.x * HASH_BLOCK + threadIdx.x;
  if (i >= count) return;
  const uint4* in = reinterpret_cast<const uint4*>(nodes + (in_base + 2 * i) * 8);
  uint4 a0 = in[0], a1 = in[1], b0 = in[2], b1 = in[3];
  uint32_t s[24] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w, b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w, 0, 0, 0, 0, 0, 0, 0, 0};
  p2::permute(s, ZKB_P2_TABLES, ones);
  uint4* o = reinterpret_cast<uint4*>(nodes + (out_base + i) * 8);
  o[0] = make_uint4(s[0], s[1], s[2], s[3]);
  o[1] = make_uint4(s[4], s[5], s[6], s[7]);
}

// ---- top of the tree: warp-cooperative permutations (poseidon2_warp.cuh) -----------------------------------------------------
// One warp per parent digest: lanes 0..15 load the two children (64 contiguous bytes), lanes 0..7 store the parent.
constexpr int COOP_WARPS = 8;                       // warps per CTA
constexpr size_t COOP_MAX_OUTPUTS = 4096;           // levels with at most this many parents use the cooperative kernels
constexpr uint32_t COOP_TAIL_INPUTS = 64;           // the last levels (<= 32 parents) run inside one 32-warp CTA
__global__ void __launch_bounds__(COOP_WARPS * 32) k_hash_fold_coop(uint32_t* __restrict__ nodes, size_t in_base, size_t out_base, uint32_t count) {
  p2w::Lane L; L.init();
  const uint32_t i = blockIdx.x * COOP_WARPS + (threadIdx.x >> 5);
  if (i >= count) return;                           // whole warps leave together
  uint32_t s = L.lane < 16u ? nodes[(in_base + 2 * (size_t)i) * 8 + L.lane] : 0u;
  s = p2w::permute(L, s);
  if (L.lane < 8u) nodes[(out_base + i) * 8 + L.lane] = s;
}
// levels top_inputs -> top_inputs / 2 -> ... -> 1 in one CTA of 32 warps (top_inputs <= 64)
__global__ void __launch_bounds__(1024) k_merkle_tail_coop(uint32_t* __restrict__ nodes, uint32_t top_inputs) {
  p2w::Lane L; L.init();
  const uint32_t w = threadIdx.x >> 5;
  for (uint32_t outs = top_inputs >> 1; outs >= 1; outs >>= 1) {
    if (w < outs) {
      uint32_t s = L.lane < 16u ? nodes[((size_t)2 * outs + 2 * w) * 8 + L.lane] : 0u;
      s = p2w::permute(L, s);
      if (L.lane < 8u) nodes[((size_t)outs + w) * 8 + L.lane] = s;
    }
    __syncthreads();
  }
}

void hash_rows(zkb_ctx* ctx, uint32_t* out, const uint32_t* matrix, size_t rows, size_t cols) {
  if (rows == 0) return;
  static int block = [] { const char* e = getenv("ZKB_HASH_BLOCK"); return e ? atoi(e) : 256; }();
  if (block == 256) k_hash_rows<256><<<grid_for(rows, 256), 256, 0, ctx->stream>>>(out, matrix, rows, (uint32_t)cols, p2_ones());
  else if (block == 64) k_hash_rows<64><<<grid_for(rows, 64), 64, 0, ctx->stream>>>(out, matrix, rows, (uint32_t)cols, p2_ones());
  else if (block == 128) k_hash_rows<128><<<grid_for(rows, 128), 128, 0, ctx->stream>>>(out, matrix, rows, (uint32_t)cols, p2_ones());
  else k_hash_rows<256><<<grid_for(rows, 256), 256, 0, ctx->stream>>>(out, matrix, rows, (uint32_t)cols, p2_ones());
  launched(ctx);
}
// Levels with more than COOP_MAX_OUTPUTS parents: one thread per permutation (throughput-bound, k_hash_fold).  Below that the
// tree is latency-bound -- fewer permutations than warps -- and every level is one launch of the warp-cooperative kernel
// (~2.3 us of dependent work per level instead of ~19 us); the last six levels share one CTA.  Round 1 ran the levels below 2048
// inputs in a single 256-thread CTA: eleven dependent levels at ~19 us each, 0.2 ms per tree that only other in-flight segments hid.
void hash_fold(zkb_ctx* ctx, uint32_t* nodes, size_t input_size, size_t output_size) {
  if (output_size == 0) return;
  static const bool coop = [] { const char* e = getenv("ZKB_MERKLE_COOP"); return !e || atoi(e) != 0; }();
  if (coop && output_size <= COOP_MAX_OUTPUTS) {
    k_hash_fold_coop<<<(unsigned)((output_size + COOP_WARPS - 1) / COOP_WARPS), COOP_WARPS * 32, 0, ctx->stream>>>(nodes, input_size, output_size, (uint32_t)output_size);
  } else {
    k_hash_fold<<<grid_for(output_size, HASH_BLOCK), HASH_BLOCK, 0, ctx->stream>>>(nodes, input_size, output_size, output_size, p2_ones());
  }
  launched(ctx);
}
void merkle_build(zkb_ctx* ctx, uint32_t* nodes, size_t rows) {
  size_t in = rows;
  while (in > COOP_TAIL_INPUTS) {
    hash_fold(ctx, nodes, in, in / 2);
    in /= 2;
  }
  if (in >= 2) {
    k_merkle_tail_coop<<<1, 1024, 0, ctx->stream>>>(nodes, (uint32_t)in);
    launched(ctx);
  }
}

}  // namespace zkb

using namespace zkb;

extern "C" {

zkb_err zkb_poseidon2_hash_rows(zkb_ctx* ctx, void* d_out_digests, const void* d_matrix, size_t rows, size_t cols) {
  ZKB_API_BEGIN
  use(ctx);
  ZKB_REQUIRE(d_out_digests && (d_matrix || cols == 0), "null buffer");
  ZKB_REQUIRE(aligned16(d_out_digests), "digest buffer must be 16-byte aligned");
  ZKB_REQUIRE(cols < (1u << 24), "too many columns");
  hash_rows(ctx, (uint32_t*)d_out_digests, (const uint32_t*)d_matrix, rows, cols);
  ZKB_API_END
}
zkb_err zkb_poseidon2_hash_fold(zkb_ctx* ctx, void* d_nodes, size_t input_size, size_t output_size) {
  ZKB_API_BEGIN
  use(ctx);
  ZKB_REQUIRE(d_nodes && aligned16(d_nodes), "null or misaligned node buffer");
  ZKB_REQUIRE(input_size == 2 * output_size, "hash_fold: input_size must be 2 * output_size");
  hash_fold(ctx, (uint32_t*)d_nodes, input_size, output_size);
  ZKB_API_END
}
zkb_err zkb_poseidon2_merkle_build(zkb_ctx* ctx, void* d_nodes, size_t rows) {
  ZKB_API_BEGIN
  use(ctx);
  ZKB_REQUIRE(d_nodes && aligned16(d_nodes), "null or misaligned node buffer");
  ZKB_REQUIRE(rows >= 1 && (rows & (rows - 1)) == 0, "merkle_build: rows must be a power of two");
  merkle_build(ctx, (uint32_t*)d_nodes, rows);
  ZKB_API_END
}

}  // extern "C"
