// Shared plumbing for libzkb200: context, error convention, launch helpers.
// Error convention follows risc0-sys `ffi_wrap`: C entry points return NULL or a malloc'd string.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>
#include <map>
#include "field.cuh"
#include "../../include/zkb200.h"

namespace zkb {

struct Error : std::runtime_error { using std::runtime_error::runtime_error; };

#define ZKB_CUDA(expr)                                                                                   \
  do {                                                                                                   \
    cudaError_t _e = (expr);                                                                             \
    if (_e != cudaSuccess)                                                                               \
      throw zkb::Error(std::string(#expr) + ": " + cudaGetErrorString(_e) + " (" __FILE__ ":" + std::to_string(__LINE__) + ")"); \
  } while (0)

#define ZKB_REQUIRE(cond, msg)                                    \
  do { if (!(cond)) throw zkb::Error(std::string("zkb200: ") + (msg)); } while (0)

#define ZKB_API_BEGIN try {
#define ZKB_API_END                                                              \
  }                                                                              \
  catch (const std::exception& e) { return strdup(e.what()); }                   \
  catch (...) { return strdup("zkb200: unknown error"); }                        \
  return nullptr;

struct NttTables;   // k_ntt.cu

}  // namespace zkb

struct zkb_ctx {
  int device = 0;
  int sm_count = 0;
  cudaStream_t stream = nullptr;
  bool owns_stream = false;
  uint64_t launches = 0;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  zkb::NttTables* ntt = nullptr;          // lazily built twiddle tables (device)
  void* scratch = nullptr;                // small device scratch (reductions, carries)
  size_t scratch_bytes = 0;
  void* staging = nullptr;                // pinned host staging for small parameter uploads
  size_t staging_bytes = 0;
  void* jit = nullptr;                    // EvalJitCache* (k_eval_jit.cu): per-circuit NVRTC-compiled eval_check kernels
  cudaMemPool_t pool = nullptr;           // this ctx's own stream-ordered pool: several ctxs proving concurrently on one GPU never
                                          // wait on (or steal) each other's freed blocks; release threshold = never
};

namespace zkb {

inline void use(zkb_ctx* ctx) {
  ZKB_REQUIRE(ctx != nullptr, "null ctx");
  ZKB_CUDA(cudaSetDevice(ctx->device));
}
inline void launched(zkb_ctx* ctx, int n = 1) {
  ctx->launches += (uint64_t)n;
  ZKB_CUDA(cudaGetLastError());
}
inline void* scratch(zkb_ctx* ctx, size_t bytes) {
  if (ctx->scratch_bytes < bytes) {
    if (ctx->scratch) { ZKB_CUDA(cudaStreamSynchronize(ctx->stream)); ZKB_CUDA(cudaFree(ctx->scratch)); }
    size_t cap = bytes < (1u << 20) ? (1u << 20) : bytes;
    ZKB_CUDA(cudaMalloc(&ctx->scratch, cap));
    ctx->scratch_bytes = cap;
  }
  return ctx->scratch;
}
// stream-ordered temporary from the ctx's own pool (freed with pool_free on the same stream)
template <typename T> inline void pool_alloc(zkb_ctx* ctx, T** out, size_t bytes) {
  ZKB_CUDA(cudaMallocFromPoolAsync((void**)out, bytes < 16 ? 16 : bytes, ctx->pool, ctx->stream));
}
inline void pool_free(zkb_ctx* ctx, void* p) { if (p) cudaFreeAsync(p, ctx->stream); }
inline bool aligned16(const void* p) { return ((uintptr_t)p & 15u) == 0; }
inline unsigned grid_for(size_t work, unsigned block) { return (unsigned)((work + block - 1) / block); }

void ntt_tables_free(zkb_ctx* ctx);   // k_ntt.cu
void eval_jit_free(zkb_ctx* ctx);     // k_eval_jit.cu

}  // namespace zkb
