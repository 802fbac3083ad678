// Batched query-phase gathers (sm_100a).
//
// In the reference, `MerkleTreeProver::prove(idx)` issues one `gather_sample` launch plus ~17 single-digest
// `Buffer::get_at` device->host reads per (query, tree): 50 x 7 x 18 tiny synchronous transfers (SURVEY.md 3.2.3,
// 7 hard-part 5).  Fiat-Shamir does not depend on what is *written* during the query phase, so all positions are
// known up front; one launch per tree writes every row and sibling path directly at its final offset in the seal,
// and the whole query section comes back in a single copy.  Word order per (query, tree) follows App. D.5: the row
// (`cols` Fp words, matrix[idx + c*rows]) then nodes[i ^ 1] for i = idx + rows, i >>= 1 while i >= 2*top_size.
#include "ops.cuh"

namespace zkb {

__global__ void k_gather_queries(uint32_t* __restrict__ out, uint32_t words_per_query, QueryTree t, const uint32_t* __restrict__ idx_list) {
  const uint32_t q = blockIdx.x;
  const uint32_t idx = idx_list[q];
  uint32_t* o = out + (size_t)q * words_per_query + t.out_offset;
  for (uint32_t c = threadIdx.x; c < t.cols; c += blockDim.x) o[c] = t.matrix[(size_t)c * t.rows + idx];
  o += t.cols;
  for (uint32_t w = threadIdx.x; w < t.path_len * 8; w += blockDim.x) {
    uint32_t level = w >> 3;
    uint32_t i = (idx + t.rows) >> level;
    o[w] = t.nodes[(size_t)(i ^ 1u) * 8 + (w & 7u)];
  }
}
void gather_queries(zkb_ctx* ctx, uint32_t* d_out, uint32_t words_per_query, const QueryTree& tree, const uint32_t* d_idx, uint32_t n_queries) {
  if (!n_queries) return;
  k_gather_queries<<<n_queries, 128, 0, ctx->stream>>>(d_out, words_per_query, tree, d_idx);
  launched(ctx);
}

// dst[r * stride + i] -= deltas[r * per_row + i]   (Fp4), for the few low coefficients touched by Prover::finalize's
// "subtract the interpolants" step.
__global__ void k_sub_small(uint4* __restrict__ dst, size_t stride, const uint4* __restrict__ deltas, uint32_t rows, uint32_t per_row) {
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= rows * per_row) return;
  uint32_t r = t / per_row, i = t % per_row;
  uint4* p = dst + (size_t)r * stride + i;
  *p = st4(ld4(*p) - ld4(deltas[t]));
}
void sub_small(zkb_ctx* ctx, uint32_t* d_fp4_dst, size_t stride_fp4, const uint32_t* d_deltas, uint32_t rows, uint32_t per_row) {
  if (!rows || !per_row) return;
  k_sub_small<<<grid_for((size_t)rows * per_row, 64), 64, 0, ctx->stream>>>((uint4*)d_fp4_dst, stride_fp4, (const uint4*)d_deltas, rows, per_row);
  launched(ctx);
}

}  // namespace zkb
