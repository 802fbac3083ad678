// Poseidon2 row hashing and Merkle folding kernels (sm_100a).
//
// Stand-ins for risc0-zkp `Hal::hash_rows` / `Hal::hash_fold` (CudaHal binds sppark_poseidon2_rows /
// sppark_poseidon2_fold for them).  INT32-pipe bound: one thread owns one
// sponge, the 24-cell state stays in registers, consecutive threads read consecutive rows of each column, so
// every global load is a fully coalesced 128-byte line.
#include "common.cuh"
#include "poseidon2.cuh"

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

// Top of the tree in one CTA: levels with <= TAIL_LEAVES/2 outputs, synchronised with __syncthreads.
constexpr int TAIL_THREADS = 256;
__global__ void __launch_bounds__(TAIL_THREADS) k_merkle_tail(uint32_t* __restrict__ nodes, uint32_t top_inputs, uint32_t ones) {
  for (uint32_t outs = top_inputs >> 1; outs >= 1; outs >>= 1) {
    for (uint32_t i = threadIdx.x; i < outs; i += TAIL_THREADS) {
      const uint4* in = reinterpret_cast<const uint4*>(nodes + ((size_t)2 * outs + 2 * i) * 8);
      uint4 a0 = in[0], a1 = in[1], b0 = in[2], b1 = in[3];
      uint32_t s[24] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w, b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w, 0, 0, 0, 0, 0, 0, 0, 0};
      p2::permute(s, ZKB_P2_TABLES, ones);
      uint4* o = reinterpret_cast<uint4*>(nodes + ((size_t)outs + i) * 8);
      o[0] = make_uint4(s[0], s[1], s[2], s[3]);
      o[1] = make_uint4(s[4], s[5], s[6], s[7]);
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
void hash_fold(zkb_ctx* ctx, uint32_t* nodes, size_t input_size, size_t output_size) {
  if (output_size == 0) return;
  k_hash_fold<<<grid_for(output_size, HASH_BLOCK), HASH_BLOCK, 0, ctx->stream>>>(nodes, input_size, output_size, output_size, p2_ones());
  launched(ctx);
}
void merkle_build(zkb_ctx* ctx, uint32_t* nodes, size_t rows) {
  const size_t TAIL_INPUTS = 2048;     // levels whose input has <= 2048 digests run inside one CTA
  size_t in = rows;
  while (in > TAIL_INPUTS) {
    hash_fold(ctx, nodes, in, in / 2);
    in /= 2;
  }
  if (in >= 2) {
    k_merkle_tail<<<1, TAIL_THREADS, 0, ctx->stream>>>(nodes, (uint32_t)in, p2_ones());
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
