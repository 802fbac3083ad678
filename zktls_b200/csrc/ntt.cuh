// Twiddle tables shared by the NTT kernels.
#pragma once
#include "common.cuh"

namespace zkb {

constexpr uint32_t TW_SPLIT_BITS = 13, TW_SPLIT = 1u << TW_SPLIT_BITS;   // 2 x 8192 words = 64 KB, L1/L2 resident
constexpr int TW_ORDER_PO2 = 26;

struct NttTables {
  uint32_t* d_hi = nullptr;   // W^(i * 8192), W = w_{2^26}
  uint32_t* d_lo = nullptr;   // W^i
  Fp rou_fwd[MAX_ROU_PO2 + 1], rou_rev[MAX_ROU_PO2 + 1];
  std::map<int, uint32_t*> level_tables;   // keyed by (log2 size << 1 | inverse): per-level twiddles for tiled kernels
  std::map<uint64_t, uint32_t*> shift_tables;
  std::map<uint64_t, uint32_t*> s_tables;      // strided-pass inter-pass factor tables (Shoup pairs), see k_ntt_s TAB
};
NttTables* ntt_tables(zkb_ctx* ctx);

// w_{2^lg}^e (fwd) or w_{2^lg}^(-e) (inv) for 0 <= e < 2^lg, lg <= 26: one modmul + two cached loads.
struct TwiddleRef {
  const uint32_t* hi;
  const uint32_t* lo;
  __device__ __forceinline__ uint32_t at(uint32_t e26) const {
    return mont_mul(__ldg(hi + (e26 >> TW_SPLIT_BITS)), __ldg(lo + (e26 & (TW_SPLIT - 1))));
  }
  __device__ __forceinline__ uint32_t fwd(uint32_t e, int lg) const { return at(e << (TW_ORDER_PO2 - lg)); }
  __device__ __forceinline__ uint32_t inv(uint32_t e, int lg) const {
    uint32_t e26 = e << (TW_ORDER_PO2 - lg);
    return at((0u - e26) & ((1u << TW_ORDER_PO2) - 1));
  }
};

}  // namespace zkb
