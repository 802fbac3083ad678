// Tiled BabyBear NTT / iNTT / LDE kernels (sm_100a): register radix-16 butterflies, shared-memory exchanges,
// four-step passes (one strided, one contiguous for n <= 2^22).
//
// Index conventions are those of risc0-zkp `core/ntt.rs`, decomposed as follows.  Write a
// position of the 2^k array as p = p1 * S + p2 (S = 2^s) and the coefficient index it holds in bit-reversed layout as
// j = rev_k(p) = rev_s(p2) * 2^a + rev_a(p1) (a = k - s).
//   inverse (DIF, natural evaluations -> bit-reversed coefficients, x 1/n):
//     pass S: for every p2, a size-2^a DIF over p1 (stride S) leaves frequency j1 = rev_a(p1') at p1'; multiply by
//             w_n^(-j1 * p2) (and by 3^j1 / n when the zk-shift is fused);
//     pass C: for every p1', a size-2^s DIF over the contiguous p2 (then x 3^(2^a * rev_s(p2')) for the zk-shift).
//   forward (DIT, bit-reversed coefficients -> natural evaluations) is the mirror image: pass C first (its lowest
//     `expand_bits` levels replaced by replication, reading the n/4-element input directly), then the twiddle w_n^(+j1 p2)
//     and the strided DIT.
// Sub-transforms larger than 2^12 are split again the same way (three passes for k > 22).
//
// Each thread keeps 16 elements in registers and runs up to four butterfly levels on them; between such rounds the
// CTA's tile goes through shared memory (padded by one word per 16 so that every round's access pattern is
// bank-conflict free).  Pass C works on 4096 contiguous elements per CTA (128-bit global accesses on the
// consecutive-16 side, 128-byte lines on the strided side); pass S works on a (2^a x T) tile whose rows are T >= 8
// consecutive elements (whole 32-byte sectors; T = 8 for a >= 9 so that two 512-thread CTAs share an SM and overlap each
// other's load / compute / store phases -- 14 % faster than one 1024-thread CTA with T = 16).  Per-level twiddles w_{2^(q+1)}^x come from an 8192-word table that
// stays in L1; the inter-pass twiddles are generated per thread as a running product G^m from two table look-ups.
#include "common.cuh"
#include "ntt.cuh"

namespace zkb {

constexpr int TILE_LOG = 12, TILE = 1 << TILE_LOG;      // elements per CTA in pass C
constexpr int C_THREADS = TILE / 16;                    // 256
constexpr int MAX_LEVEL_LOG = 12;                       // level tables cover q < 12
// the all-ones launch argument of the ALU-pinned additions in radix_round
static inline uint32_t ntt_ones() { return 0xffffffffu; }

__device__ __forceinline__ uint32_t phys(uint32_t x) { return x + (x >> 4); }
constexpr uint32_t phys_size(uint32_t n) { return n + (n >> 4) + 1; }

// bit reversal of the 4-bit register index
__host__ __device__ constexpr int rev4(int r) { return ((r & 1) << 3) | ((r & 2) << 1) | ((r & 4) >> 1) | ((r & 8) >> 3); }

// One round of butterfly levels on the 16 registers.  The registers are the values of local-index bits [p, p+4);
// levels JLO <= j < JHI (absolute bit q = p + j) are processed: descending for the DIF (inverse), ascending for the
// DIT (forward).  `lo` = the thread's local-index bits below p.  twl = per-level twiddle table for this direction.
// x * w mod P in [0, 2P) for ANY 32-bit x, w a plain (non-Montgomery) constant with its Shoup quotient wq = floor(w 2^32 / P):
// one IMAD.HI + two IMAD (the Montgomery form needs IMAD.WIDE + IMAD + IMAD.HI and a canonical-range operand).
// (written with an explicit mad.lo: left to itself ptxas computed q * P and x * w separately and joined them with an IMAD.IADD in two
// thirds of the butterflies -- a third multiplier-pipe slot per product)
__device__ __forceinline__ uint32_t shoup_lazy(uint32_t x, uint2 w) {
  uint32_t r = x * w.x, q = __umulhi(x, w.y);
  asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(r) : "r"(q), "r"(0u - P));
  return r;
}

// Values stay LAZY, in [0, 2P), between butterfly levels: each butterfly brings its two inputs below P (one VIADDMNMX
// each), then a' = a + t and b' = a - t + P are again below 2P with no further correction -- 4 ALU + 3 multiplier-pipe
// instructions per butterfly instead of 6 + 3.  Callers reduce once before storing (or multiply by a canonical
// Montgomery word, which accepts a lazy operand).
// `ones` = 0xffffffff from a launch argument: min(a + b, ones) is ONE VIADDMNMX, i.e. the butterfly's two-input addition stays
// on the ALU pipe instead of becoming IMAD.IADD on the multiplier pipe the Shoup products use (see poseidon2.cuh).
template <bool INV, int JLO, int JHI>
__device__ __forceinline__ void radix_round(uint32_t (&v)[16], const int p, const uint32_t lo, const uint2* __restrict__ twl, const uint32_t ones) {
  if (INV) {
#pragma unroll
    for (int j = JHI - 1; j >= JLO; --j) {
      const int q = p + j;
#pragma unroll
      for (int r = 0; r < 16; ++r) {
        if (r & (1 << j)) continue;
        uint32_t a = reduce_2p(v[r]), b = reduce_2p(v[r | (1 << j)]);
        v[r] = min(a + b, ones);
        uint32_t d = a - b + P;
        // w^0 = 1: level 0 always; in the round at bit 0 (lo == 0 there) also every butterfly whose in-register offset is 0
        if (q == 0 || (p == 0 && (r & ((1 << j) - 1)) == 0)) v[r | (1 << j)] = d;
        else v[r | (1 << j)] = shoup_lazy(d, __ldg(twl + (1u << q) + ((uint32_t)(r & ((1 << j) - 1)) << p) + lo));
      }
    }
  } else {
#pragma unroll
    for (int j = JLO; j < JHI; ++j) {
      const int q = p + j;
#pragma unroll
      for (int r = 0; r < 16; ++r) {
        if (r & (1 << j)) continue;
        uint32_t a = reduce_2p(v[r]), b = v[r | (1 << j)];
        if (q != 0 && !(p == 0 && (r & ((1 << j) - 1)) == 0)) b = shoup_lazy(b, __ldg(twl + (1u << q) + ((uint32_t)(r & ((1 << j) - 1)) << p) + lo));
        b = reduce_2p(b);
        v[r] = min(a + b, ones);
        v[r | (1 << j)] = a - b + P;
      }
    }
  }
}
template <int N> __device__ __forceinline__ void canonicalize(uint32_t (&v)[N]) {
#pragma unroll
  for (int r = 0; r < N; ++r) v[r] = reduce_2p(v[r]);
}

// Shared-memory slot of register r in a round whose register field starts at bit SH of the (unpadded) tile index, given
// pb = phys(x0) for the index x0 with r = 0: phys(x0 + (r << SH)) = pb + K + (K >> 4), K = r << SH, because the r-field of x0
// is empty (no carry into or out of it).  K is a compile-time constant after unrolling, so every access is one LDS / STS with
// an immediate offset instead of ~4 index instructions.
__device__ __forceinline__ uint32_t phys_r(uint32_t pb, int r, int sh) { const uint32_t K = (uint32_t)r << sh; return pb + K + (K >> 4); }
// Strided pass: the tile is (rows x 2^TL); padding adds 2^TL words per 16 rows (same footprint as phys).  A warp holds
// 2^(5-TL) consecutive threads-in-column x 2^TL columns: when their rows are 16 apart (round p = 0) the pad moves each onto its
// own group of banks, and when they are consecutive rows the pad is constant across the warp, so every round is conflict-free
// (phys(), whose pad changes every 16 WORDS, made neighbouring rows collide once the tile became 8 wide: 88 M conflict cycles).
template <int TL> __device__ __forceinline__ uint32_t phys_t(uint32_t x) { return x + ((x >> (4 + TL)) << TL); }
template <int TL> __device__ __forceinline__ uint32_t phys_tr(uint32_t pb, int r, int sh) { const uint32_t K = (uint32_t)r << sh; return pb + K + ((K >> (4 + TL)) << TL); }
// the same with the width as an argument (a compile-time constant after unrolling): used by the contiguous pass, whose exchange
// between the rounds at bits p_prev and pc is conflict-free for BOTH access patterns with TL = min(p_prev, pc) -- with the plain
// phys() padding the 32 consecutive words a warp touches in the round at bit 8 straddle a pad step (2-way conflicts, 30 M cycles).
__device__ __forceinline__ uint32_t phys_v(uint32_t x, int tl) { return x + ((x >> (4 + tl)) << tl); }
__device__ __forceinline__ uint32_t phys_vr(uint32_t pb, int r, int sh, int tl) { const uint32_t K = (uint32_t)r << sh; return pb + K + ((K >> (4 + tl)) << tl); }
// local index of register r for thread t in a round whose register field starts at bit p
__device__ __forceinline__ uint32_t local_index(uint32_t t, int p, int r) {
  uint32_t lo = t & ((1u << p) - 1u), hi = t >> p;
  return (hi << (p + 4)) | ((uint32_t)r << p) | lo;
}

// ---- pass C: contiguous sub-transforms of size 2^A (4 <= A <= 12), 4096 elements per CTA ---------------------------
enum : int { EPI_NONE = 0, EPI_SCALE = 1, EPI_TABLE = 2 };

// Inverse: in place.  epilogue multiplies position x of every 2^A block by table[x] (EPI_TABLE) or by `scale`.
template <int A, int EPI>
__global__ void __launch_bounds__(C_THREADS) k_ntt_c_inv(uint32_t* __restrict__ io, const uint2* __restrict__ twl, const uint32_t* __restrict__ table, uint32_t scale, uint32_t ones) {
  __shared__ uint32_t sm[phys_size(TILE)];
  constexpr int L = 1 << A, TPB = L / 16;           // threads per sub-transform
  const uint32_t tid = threadIdx.x, sub = tid / TPB, t = tid % TPB;
  uint32_t* base = io + (size_t)blockIdx.x * TILE + (size_t)sub * L;
  uint32_t* s = sm;   // sub-transform `sub` lives at [sub*L, (sub+1)*L) of the tile
  const uint32_t soff = sub * L;
  uint32_t v[16];
  constexpr int REM = A % 4;
  constexpr int P0 = A - 4;
  // first round: field [A-4, A): element r * TPB + t -> coalesced loads
#pragma unroll
  for (int r = 0; r < 16; ++r) v[r] = base[(uint32_t)r * TPB + t];
  radix_round<true, 0, 4>(v, P0, t, twl, ones);
  int p_prev = P0;
#pragma unroll
  for (int p = P0 - 4; p >= 0 || (p > -4 && REM != 0); p -= 4) {
    const int pc = p < 0 ? 0 : p;
    const int tl = pc < p_prev ? pc : p_prev;
    const uint32_t pb_w = phys_v(soff + local_index(t, p_prev, 0), tl), pb_r = phys_v(soff + local_index(t, pc, 0), tl);
#pragma unroll
    for (int r = 0; r < 16; ++r) s[phys_vr(pb_w, r, p_prev, tl)] = v[r];
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 16; ++r) v[r] = s[phys_vr(pb_r, r, pc, tl)];
    __syncthreads();
    if (p >= 0) radix_round<true, 0, 4>(v, pc, t & ((1u << pc) - 1u), twl, ones);
    else radix_round<true, 0, (REM == 0 ? 4 : REM)>(v, 0, 0u, twl, ones);
    p_prev = pc;
  }
  // here p_prev == 0 (A >= 4): registers are 16 consecutive elements
  uint32_t* o = base + 16 * t;
  if (EPI == EPI_TABLE) {
    const uint4* tb = reinterpret_cast<const uint4*>(table + 16 * t);
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      uint4 w = __ldg(tb + g);
      v[4 * g] = mont_mul_alu(v[4 * g], w.x, ones); v[4 * g + 1] = mont_mul_alu(v[4 * g + 1], w.y, ones);
      v[4 * g + 2] = mont_mul_alu(v[4 * g + 2], w.z, ones); v[4 * g + 3] = mont_mul_alu(v[4 * g + 3], w.w, ones);
    }
  } else if (EPI == EPI_SCALE) {
#pragma unroll
    for (int r = 0; r < 16; ++r) v[r] = mont_mul_alu(v[r], scale, ones);
  } else {
    canonicalize(v);
  }
#pragma unroll
  for (int g = 0; g < 4; ++g) reinterpret_cast<uint4*>(o)[g] = make_uint4(v[4 * g], v[4 * g + 1], v[4 * g + 2], v[4 * g + 3]);
}

// Forward: out-of-place capable; EB = expand bits (0 or 2): out block element x <- in[(block_base + x) >> EB], and the
// lowest EB levels are skipped.
template <int A, int EB>
__global__ void __launch_bounds__(C_THREADS) k_ntt_c_fwd(uint32_t* __restrict__ out, const uint32_t* __restrict__ in, const uint2* __restrict__ twl, uint32_t ones) {
  __shared__ uint32_t sm[phys_size(TILE)];
  constexpr int L = 1 << A, TPB = L / 16;
  const uint32_t tid = threadIdx.x, sub = tid / TPB, t = tid % TPB;
  const size_t gbase = (size_t)blockIdx.x * TILE + (size_t)sub * L;
  const uint32_t soff = sub * L;
  uint32_t* s = sm;
  uint32_t v[16];
  if (EB == 0) {
    const uint4* src = reinterpret_cast<const uint4*>(in + gbase + 16 * t);
#pragma unroll
    for (int g = 0; g < 4; ++g) { uint4 w = src[g]; v[4 * g] = w.x; v[4 * g + 1] = w.y; v[4 * g + 2] = w.z; v[4 * g + 3] = w.w; }
  } else {   // EB == 2: 16 consecutive outputs come from 4 consecutive inputs
    uint4 w = *reinterpret_cast<const uint4*>(in + ((gbase + 16 * t) >> 2));
#pragma unroll
    for (int r = 0; r < 4; ++r) { v[r] = w.x; v[4 + r] = w.y; v[8 + r] = w.z; v[12 + r] = w.w; }
  }
  constexpr int REM = A % 4;
  constexpr int PLAST = A - 4;
  radix_round<false, EB, 4>(v, 0, 0u, twl, ones);
  int p_prev = 0;
#pragma unroll
  for (int p = 4; p <= PLAST || (p < PLAST + 4 && REM != 0); p += 4) {
    const int pc = p > PLAST ? PLAST : p;
    const int tl = pc < p_prev ? pc : p_prev;
    const uint32_t pb_w = phys_v(soff + local_index(t, p_prev, 0), tl), pb_r = phys_v(soff + local_index(t, pc, 0), tl);
#pragma unroll
    for (int r = 0; r < 16; ++r) s[phys_vr(pb_w, r, p_prev, tl)] = v[r];
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 16; ++r) v[r] = s[phys_vr(pb_r, r, pc, tl)];
    __syncthreads();
    if (p <= PLAST) radix_round<false, 0, 4>(v, pc, t & ((1u << pc) - 1u), twl, ones);
    else radix_round<false, (REM == 0 ? 0 : 4 - REM), 4>(v, pc, t & ((1u << pc) - 1u), twl, ones);
    p_prev = pc;
  }
  // p_prev == A - 4: register r is element r * TPB + t
  canonicalize(v);
  uint32_t* o = out + gbase;
#pragma unroll
  for (int r = 0; r < 16; ++r) o[(uint32_t)r * TPB + t] = v[r];
}

// ---- pass S: strided sub-transforms of size 2^A over a (2^A x T) tile, T consecutive elements per row -------------
// The array is a sequence of blocks of M = 2^A * S elements (S = 2^s_log); within a block, position = p1 * S + p2.
struct SArgs {
  uint32_t s_log;          // log2 S
  uint32_t tiles_per_block_log;   // log2 (S / T)
  uint32_t shift_g;        // inverse only: B^(2^(A-4)) (Montgomery), B = shift base of this pass; R_MOD_P if none
  uint32_t use_table;      // inverse only: multiply V by table[rev_{A-4}(t')] (includes the 1/n scale where needed)
};

// TL = log2 of the tile width (consecutive elements per row); 0 = default: 16 columns for A >= 8, else a 4096-element tile.
template <int A, int TL = 0> struct STile {
  static constexpr int T_LOG = TL ? TL : ((A >= 8) ? 4 : (12 - A));
  static constexpr int T = 1 << T_LOG; static constexpr int THREADS = (1 << A) * T / 16;
};

// TAB: the inter-pass factors come from a precomputed (2^A x S) table of Shoup pairs laid out like the data block
// (ttab[l * S + p2] = mult * B^(rev_A(l)) * w_M^(+-rev_A(l) * p2)): one LDG.64 + 3 multiplier-pipe instructions per element,
// against ~2.2 Montgomery multiplications per element for the running-product form.  The table (8 M bytes, <= 32 MB) is
// shared by all columns and stays in L2.
// SL: log2 S as a compile-time constant (0 = take it from args): every global access of the tile is then base + immediate
// (the 32 loads / stores of a thread cost ~80 address instructions otherwise -- 7 % of the kernel's issue slots).
// TAB = 2 (forward): FACTORED tables.  rev_A(16 t + r) = rev4(r) 2^(A-4) + rev_{A-4}(t), so the factor of register r is
// gtab[r][p2] * vtab[t][p2] with gtab[r][p2] = w_M^(rev4(r) 2^(A-4) p2) (16 x S Shoup pairs, 512 KB: the 8 columns of a tile are
// 1 KB, L1-resident) and vtab[t][p2] = w_M^(rev_{A-4}(t) p2) (2^(A-4) x S pairs, 2 MB, one load per thread): two Shoup products per
// element (8 multiplier-pipe slots, no ALU work) instead of the running product's two Montgomery products (10 slots + 6 ALU
// instructions) and without the full table's 8 bytes of L2 traffic per element.
template <int A, bool INV, int TL = 0, int TAB = 0, int SL = 0>
__global__ void __launch_bounds__(STile<A, TL>::THREADS) k_ntt_s(uint32_t* __restrict__ io, const uint2* __restrict__ twl, TwiddleRef tw, const uint32_t* __restrict__ table, SArgs args,
                                                                 const uint2* __restrict__ ttab, const uint32_t* __restrict__ src, uint32_t ones) {
  constexpr int T_LOG = STile<A, TL>::T_LOG, T = STile<A, TL>::T, TPB = (1 << A) / 16;
  extern __shared__ uint32_t sm[];
  const uint32_t tid = threadIdx.x;
  const uint32_t c2 = tid & (T - 1), t = tid >> T_LOG;                      // column inside the tile, thread inside the column
  const uint32_t tile = blockIdx.x & ((1u << args.tiles_per_block_log) - 1u);
  const size_t block = blockIdx.x >> args.tiles_per_block_log;
  const uint32_t p2 = (tile << T_LOG) + c2;                                 // position inside the row of S
  const uint32_t s_log = SL ? (uint32_t)SL : args.s_log;
  const int m_log = A + (int)s_log;
  uint32_t* const cb = io + (block << m_log);            // CTA-uniform; everything below is a 32-bit offset (< 2^26 elements)
  const uint32_t* const cin = src ? src + (block << m_log) : cb;      // inverse pass: optional separate input (the caller's trace), saving a copy
  const uint32_t off_r = (t << s_log) + p2;                // element (r * TPB + t, p2): off_r + r * (TPB << s_log)
  const uint32_t off_c = ((16u * t) << s_log) + p2;        // element (16 t + r, p2):   off_c + (r << s_log)
  const uint32_t S = 1u << s_log;
  uint32_t v[16];
  constexpr int REM = A % 4;

  // inter-pass twiddles: register r (holding local position 16 t + r) needs G^(rev_A(16 t + r)) = V * g^(rev4(r)),
  // G = w_M^(+-p2) [* shift base], V = G^(rev_{A-4}(t)), g = G^(2^(A-4)).
  auto twiddle_all = [&](bool inverse) {
    if (TAB == 2) {
      const uint2 V = __ldg(ttab + ((size_t)16 << s_log) + ((size_t)t << s_log) + p2);
#pragma unroll
      for (int r = 0; r < 16; ++r) v[r] = shoup_lazy(shoup_lazy(v[r], __ldg(ttab + ((size_t)r << s_log) + p2)), V);
      return;
    }
    if (TAB == 1) {
#pragma unroll
      for (int r = 0; r < 16; ++r) {
        uint32_t x = shoup_lazy(v[r], __ldg(ttab + off_c + ((size_t)r << s_log)));
        v[r] = inverse ? reduce_2p(x) : x;          // forward: the butterflies take lazy values; inverse: stored next
      }
      return;
    }
    uint32_t rt = A > 4 ? bit_rev32(t, A - 4) : 0u;
    uint32_t e_v = rt * p2, e_g = p2 << (A - 4);
    uint32_t V = inverse ? tw.inv(e_v, m_log) : tw.fwd(e_v, m_log);
    uint32_t g = inverse ? tw.inv(e_g, m_log) : tw.fwd(e_g, m_log);
    if (inverse) {
      if (args.use_table) V = mont_mul(V, __ldg(table + t));     // table[t] = mult * B^(rev_{A-4}(t))
      if (args.shift_g != R_MOD_P) g = mont_mul(g, args.shift_g);
    }
    uint32_t cur = V;
#pragma unroll
    for (int m = 0; m < 16; ++m) {
      v[rev4(m)] = mont_mul_alu(v[rev4(m)], cur, ones);
      if (m != 15) cur = mont_mul_alu(cur, g, ones);
    }
  };

  if (INV) {
    constexpr int P0 = A - 4;
#pragma unroll
    { const uint32_t* const pr = cin + off_r;
#pragma unroll
      for (int r = 0; r < 16; ++r) v[r] = pr[(size_t)((uint32_t)r * TPB) << s_log]; }
    radix_round<true, 0, 4>(v, P0, t, twl, ones);
    int p_prev = P0;
#pragma unroll
    for (int p = P0 - 4; p >= 0 || (p > -4 && REM != 0); p -= 4) {
      const int pc = p < 0 ? 0 : p;
      const uint32_t pb_w = phys_t<T_LOG>((local_index(t, p_prev, 0) << T_LOG) + c2), pb_r = phys_t<T_LOG>((local_index(t, pc, 0) << T_LOG) + c2);
#pragma unroll
      for (int r = 0; r < 16; ++r) sm[phys_tr<T_LOG>(pb_w, r, p_prev + T_LOG)] = v[r];
      __syncthreads();
#pragma unroll
      for (int r = 0; r < 16; ++r) v[r] = sm[phys_tr<T_LOG>(pb_r, r, pc + T_LOG)];
      __syncthreads();
      if (p >= 0) radix_round<true, 0, 4>(v, pc, t & ((1u << pc) - 1u), twl, ones);
      else radix_round<true, 0, (REM == 0 ? 4 : REM)>(v, 0, 0u, twl, ones);
      p_prev = pc;
    }
    twiddle_all(true);
    { uint32_t* const pc = cb + off_c;
#pragma unroll
      for (int r = 0; r < 16; ++r) pc[(size_t)r << s_log] = v[r]; }
  } else {
    constexpr int PLAST = A - 4;
    { const uint32_t* const pc = cb + off_c;
#pragma unroll
      for (int r = 0; r < 16; ++r) v[r] = pc[(size_t)r << s_log]; }
    twiddle_all(false);
    radix_round<false, 0, 4>(v, 0, 0u, twl, ones);
    int p_prev = 0;
#pragma unroll
    for (int p = 4; p <= PLAST || (p < PLAST + 4 && REM != 0); p += 4) {
      const int pc = p > PLAST ? PLAST : p;
      const uint32_t pb_w = phys_t<T_LOG>((local_index(t, p_prev, 0) << T_LOG) + c2), pb_r = phys_t<T_LOG>((local_index(t, pc, 0) << T_LOG) + c2);
#pragma unroll
      for (int r = 0; r < 16; ++r) sm[phys_tr<T_LOG>(pb_w, r, p_prev + T_LOG)] = v[r];
      __syncthreads();
#pragma unroll
      for (int r = 0; r < 16; ++r) v[r] = sm[phys_tr<T_LOG>(pb_r, r, pc + T_LOG)];
      __syncthreads();
      if (p <= PLAST) radix_round<false, 0, 4>(v, pc, t & ((1u << pc) - 1u), twl, ones);
      else radix_round<false, (REM == 0 ? 0 : 4 - REM), 4>(v, pc, t & ((1u << pc) - 1u), twl, ones);
      p_prev = pc;
    }
    canonicalize(v);
    { uint32_t* const pr = cb + off_r;
#pragma unroll
      for (int r = 0; r < 16; ++r) pr[(size_t)((uint32_t)r * TPB) << s_log] = v[r]; }
  }
}

// factored forward tables (k_ntt_s TAB = 2): [16 x S] g-part, then [2^(a-4) x S] v-part
__global__ void k_ntt_s_table2(uint2* __restrict__ out, TwiddleRef tw, int a, int s_log) {
  const uint32_t m_log = (uint32_t)(a + s_log), S = 1u << s_log;
  const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x, total = (16u + (1u << (a - 4))) << s_log;
  if (idx >= total) return;
  const uint32_t row = idx >> s_log, p2 = idx & (S - 1u);
  const uint32_t j = row < 16u ? (bit_rev32(row, 4) << (a - 4)) : bit_rev32(row - 16u, a - 4);
  const uint32_t e = (j * p2) & ((1u << m_log) - 1u);
  const uint32_t plain = mont_mul(tw.fwd(e, (int)m_log), 1u);
  out[idx] = make_uint2(plain, (uint32_t)(((uint64_t)plain << 32) / P));
}

// ttab[l * S + p2] for one strided pass (see k_ntt_s TAB)
__global__ void k_ntt_s_table(uint2* __restrict__ out, TwiddleRef tw, int a, int s_log, int inverse, uint32_t shift_base, uint32_t mult) {
  const uint32_t m_log = (uint32_t)(a + s_log);
  uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (1u << m_log)) return;
  uint32_t l = idx >> s_log, p2 = idx & ((1u << s_log) - 1u);
  uint32_t j1 = bit_rev32(l, a);
  uint32_t e = (j1 * p2) & ((1u << m_log) - 1u);
  uint32_t w = inverse ? tw.inv(e, (int)m_log) : tw.fwd(e, (int)m_log);
  if (inverse) w = mont_mul(w, mont_mul(mult, pow(Fp::raw(shift_base), j1).v));
  uint32_t plain = mont_mul(w, 1u);
  out[idx] = make_uint2(plain, (uint32_t)(((uint64_t)plain << 32) / P));
}

// ---- host side: tables, plans, launches ---------------------------------------------------------------------------
constexpr int S_TABLE_MAX_LOG = 22;      // tables up to 2^22 entries (32 MB)
// Measured (B200, 224 x 2^20): the table form is 4 % faster for the inverse pass (which also folds the zk-shift / scale
// factors) and neutral for the forward pass, where the extra 32 MB of L2 traffic cancels the saved multiplications --
// so it is used for the inverse only.  ZKB_NTT_S_TABLE = 0: never, 2: both directions.
static int s_tables_mode() { static int v = [] { const char* e = getenv("ZKB_NTT_S_TABLE"); return e ? atoi(e) : 1; }(); return v; }
static const uint2* s_pass_table(zkb_ctx* ctx, int a, int s_log, bool inverse, Fp shift_base, Fp mult) {
  if (s_tables_mode() == 0 || (!inverse && s_tables_mode() < 2) || a + s_log > S_TABLE_MAX_LOG) return nullptr;
  NttTables* t = ntt_tables(ctx);
  uint64_t key = 0x5000000000000000ull ^ ((uint64_t)a << 52) ^ ((uint64_t)s_log << 44) ^ ((uint64_t)inverse << 43) ^ ((uint64_t)shift_base.v * 0x9E3779B97F4A7C15ull) ^ ((uint64_t)mult.v << 7);
  auto it = t->s_tables.find(key);
  if (it != t->s_tables.end()) return (const uint2*)it->second;
  TwiddleRef tw{t->d_hi, t->d_lo};
  size_t m = (size_t)1 << (a + s_log);
  uint32_t* d = nullptr;
  ZKB_CUDA(cudaMalloc((void**)&d, m * 8));
  k_ntt_s_table<<<(unsigned)((m + 255) / 256), 256, 0, ctx->stream>>>((uint2*)d, tw, a, s_log, inverse ? 1 : 0, shift_base.v, mult.v);
  launched(ctx);
  t->s_tables[key] = d;
  return (const uint2*)d;
}

// 0 = running product, 1 = factored tables (default; a >= 5 only), 2 = full table
static int s_fwd_mode() { static int v = [] { const char* e = getenv("ZKB_NTT_S_FWD"); return e ? atoi(e) : 1; }(); return v; }
static const uint2* s_pass_table2(zkb_ctx* ctx, int a, int s_log) {
  if (s_fwd_mode() != 1 || a < 5 || a + s_log > 26) return nullptr;
  NttTables* t = ntt_tables(ctx);
  uint64_t key = 0x6000000000000000ull ^ ((uint64_t)a << 52) ^ ((uint64_t)s_log << 44);
  auto it = t->s_tables.find(key);
  if (it != t->s_tables.end()) return (const uint2*)it->second;
  TwiddleRef tw{t->d_hi, t->d_lo};
  size_t m = ((size_t)16 + ((size_t)1 << (a - 4))) << s_log;
  uint32_t* d = nullptr;
  ZKB_CUDA(cudaMalloc((void**)&d, m * 8));
  k_ntt_s_table2<<<(unsigned)((m + 255) / 256), 256, 0, ctx->stream>>>((uint2*)d, tw, a, s_log);
  launched(ctx);
  t->s_tables[key] = d;
  return (const uint2*)d;
}

static const uint2* level_table(zkb_ctx* ctx, bool inverse) {
  NttTables* t = ntt_tables(ctx);
  int key = inverse ? 1 : 0;
  auto it = t->level_tables.find(key);
  if (it != t->level_tables.end()) return (const uint2*)it->second;
  // entry (1 << q) + x = w_{2^(q+1)}^(+-x) as a PLAIN integer with its Shoup quotient floor(w 2^32 / P)
  std::vector<uint32_t> h((size_t)2 << MAX_LEVEL_LOG, 0);
  h[0] = 1; h[1] = (uint32_t)(((uint64_t)1 << 32) / P);
  for (int q = 0; q < MAX_LEVEL_LOG; ++q) {
    Fp w = inverse ? t->rou_rev[q + 1] : t->rou_fwd[q + 1];
    Fp cur = Fp::one();
    for (uint32_t x = 0; x < (1u << q); ++x) {
      uint32_t plain = cur.as_u32();
      h[2 * ((1u << q) + x)] = plain;
      h[2 * ((1u << q) + x) + 1] = (uint32_t)(((uint64_t)plain << 32) / P);
      cur *= w;
    }
  }
  uint32_t* d = nullptr;
  ZKB_CUDA(cudaMalloc((void**)&d, h.size() * 4));
  ZKB_CUDA(cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
  t->level_tables[key] = d;
  return (const uint2*)d;
}
// table[x] = mult * base^(rev_bits(x)) for x < 2^bits
static const uint32_t* shift_table(zkb_ctx* ctx, int bits, Fp base, Fp mult) {
  NttTables* t = ntt_tables(ctx);
  uint64_t key = ((uint64_t)bits << 58) ^ ((uint64_t)base.v << 26) ^ (uint64_t)mult.v * 0x9E3779B97F4A7C15ull;
  auto it = t->shift_tables.find(key);
  if (it != t->shift_tables.end()) return it->second;
  size_t n = (size_t)1 << bits;
  std::vector<uint32_t> h(n);
  std::vector<Fp> pw(n);
  Fp cur = mult;
  for (size_t j = 0; j < n; ++j) { pw[j] = cur; cur *= base; }
  for (size_t x = 0; x < n; ++x) h[x] = pw[bit_rev32((uint32_t)x, bits)].v;
  uint32_t* d = nullptr;
  ZKB_CUDA(cudaMalloc((void**)&d, std::max<size_t>(n, 4) * 4));
  ZKB_CUDA(cudaMemcpy(d, h.data(), n * 4, cudaMemcpyHostToDevice));
  t->shift_tables[key] = d;
  return d;
}

struct Plan { int n_s; int a_s[2]; int a_c; };
// k = a_s[0] + a_s[1] + a_c; prefers multiples of 4 (no partial rounds) and a 2^12 contiguous pass
static bool make_plan(int k, Plan& pl) {
  if (k < 4) return false;
  pl.n_s = 0; pl.a_s[0] = pl.a_s[1] = 0;
  if (k <= 12) { pl.a_c = k; return true; }
  if (k <= 15) { pl.a_c = 8; pl.a_s[0] = k - 8; pl.n_s = 1; return true; }     // 5..7
  if (k <= 22) { pl.a_c = 12; pl.a_s[0] = k - 12; pl.n_s = 1; return true; }   // 4..10
  if (k <= 32) { pl.a_c = 12; int rem = k - 12; pl.a_s[1] = (rem + 1) / 2; pl.a_s[0] = rem - pl.a_s[1]; pl.n_s = 2; return pl.a_s[1] <= 10; }
  return false;
}

template <int A, bool INV, int TL = 0>
static void launch_s(zkb_ctx* ctx, uint32_t* io, size_t total_elems, int s_log, const uint2* twl, TwiddleRef tw, const uint32_t* table, uint32_t shift_g, const uint2* ttab, const uint32_t* src, int tab_mode) {
  using ST = STile<A, TL>;
  SArgs args{(uint32_t)s_log, (uint32_t)(s_log - ST::T_LOG), shift_g, table ? 1u : 0u};
  size_t tile_elems = (size_t)(1 << A) * ST::T;
  size_t smem = (size_t)phys_size((uint32_t)tile_elems) * 4;
  if (!ttab) tab_mode = 0;
  auto kern = s_log == 12 ? (tab_mode == 2 ? k_ntt_s<A, INV, TL, 2, 12> : tab_mode == 1 ? k_ntt_s<A, INV, TL, 1, 12> : k_ntt_s<A, INV, TL, 0, 12>)
                          : (tab_mode == 2 ? k_ntt_s<A, INV, TL, 2> : tab_mode == 1 ? k_ntt_s<A, INV, TL, 1> : k_ntt_s<A, INV, TL, 0>);
  if (smem > 48 * 1024) ZKB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<(unsigned)(total_elems / tile_elems), ST::THREADS, smem, ctx->stream>>>(io, twl, tw, table, args, ttab, src, ntt_ones());
  launched(ctx);
}
static int s_tile_log() { static int v = [] { const char* e = getenv("ZKB_NTT_S_TLOG"); return e ? atoi(e) : 3; }(); return v; }
template <bool INV>
static void dispatch_s(zkb_ctx* ctx, int a, uint32_t* io, size_t total, int s_log, const uint2* twl, TwiddleRef tw, const uint32_t* table, uint32_t shift_g, const uint2* ttab, const uint32_t* src = nullptr, int tab_mode = 1) {
  switch (a) {
    case 4: launch_s<4, INV>(ctx, io, total, s_log, twl, tw, table, shift_g, ttab, src, tab_mode); break;
    case 5: launch_s<5, INV>(ctx, io, total, s_log, twl, tw, table, shift_g, ttab, src, tab_mode); break;
    case 6: launch_s<6, INV>(ctx, io, total, s_log, twl, tw, table, shift_g, ttab, src, tab_mode); break;
    case 7: launch_s<7, INV>(ctx, io, total, s_log, twl, tw, table, shift_g, ttab, src, tab_mode); break;
    case 8: launch_s<8, INV>(ctx, io, total, s_log, twl, tw, table, shift_g, ttab, src, tab_mode); break;
    case 9:
      if (s_tile_log() == 3) launch_s<9, INV, 3>(ctx, io, total, s_log, twl, tw, table, shift_g, ttab, src, tab_mode);
      else if (s_tile_log() == 2) launch_s<9, INV, 2>(ctx, io, total, s_log, twl, tw, table, shift_g, ttab, src, tab_mode);
      else launch_s<9, INV>(ctx, io, total, s_log, twl, tw, table, shift_g, ttab, src, tab_mode);
      break;
    case 10:
      if (s_tile_log() == 3) launch_s<10, INV, 3>(ctx, io, total, s_log, twl, tw, table, shift_g, ttab, src, tab_mode);
      else if (s_tile_log() == 2) launch_s<10, INV, 2>(ctx, io, total, s_log, twl, tw, table, shift_g, ttab, src, tab_mode);
      else launch_s<10, INV>(ctx, io, total, s_log, twl, tw, table, shift_g, ttab, src, tab_mode);
      break;
    default: throw Error("zkb200: unsupported strided NTT size");
  }
}
template <int EPI>
static void dispatch_c_inv(zkb_ctx* ctx, int a, uint32_t* io, size_t total, const uint2* twl, const uint32_t* table, uint32_t scale) {
  unsigned grid = (unsigned)(total / TILE);
#define ZKB_CI(AA) case AA: k_ntt_c_inv<AA, EPI><<<grid, C_THREADS, 0, ctx->stream>>>(io, twl, table, scale, ntt_ones()); break;
  switch (a) { ZKB_CI(4) ZKB_CI(5) ZKB_CI(6) ZKB_CI(7) ZKB_CI(8) ZKB_CI(9) ZKB_CI(10) ZKB_CI(11) ZKB_CI(12) default: throw Error("zkb200: unsupported contiguous NTT size"); }
#undef ZKB_CI
  launched(ctx);
}
template <int EB>
static void dispatch_c_fwd(zkb_ctx* ctx, int a, uint32_t* out, const uint32_t* in, size_t total, const uint2* twl) {
  unsigned grid = (unsigned)(total / TILE);
#define ZKB_CF(AA) case AA: k_ntt_c_fwd<AA, EB><<<grid, C_THREADS, 0, ctx->stream>>>(out, in, twl, ntt_ones()); break;
  switch (a) { ZKB_CF(4) ZKB_CF(5) ZKB_CF(6) ZKB_CF(7) ZKB_CF(8) ZKB_CF(9) ZKB_CF(10) ZKB_CF(11) ZKB_CF(12) default: throw Error("zkb200: unsupported contiguous NTT size"); }
#undef ZKB_CF
  launched(ctx);
}

// Columns per launch.  Measured on B200 (profiles/r1_e_ntt_sweep.txt): the passes are INT32-pipe bound, not HBM bound, so
// keeping a batch L2-resident between passes buys nothing while small launches lose to tails -- the default is the whole
// group in one launch per pass (LDE of 224 x 2^20: 8.0 ms with 48 MB batches, 5.8 ms unbatched).  ZKB_NTT_L2_BYTES
// restores batching for experiments.
static size_t batch_columns(size_t count, size_t bytes_per_column) {
  static size_t budget = [] { const char* e = getenv("ZKB_NTT_L2_BYTES"); return e ? (size_t)atoll(e) : (size_t)1 << 40; }();
  size_t b = std::max<size_t>(1, budget / std::max<size_t>(bytes_per_column, 1));
  size_t gran = std::max<size_t>(1, ((size_t)TILE * 4) / std::max<size_t>(bytes_per_column, 1));   // whole tiles per batch
  if (b >= count) return count;
  return std::max(gran, b / gran * gran);
}

static bool force_levels() { const char* e = getenv("ZKB_NTT_FORCE_LEVELS"); return e && e[0] == '1'; }

// src (optional): the evaluations are read from src and the coefficients written to io (src is left untouched); when the plan
// has no strided pass the data is copied first.
bool ntt_inverse_tiled(zkb_ctx* ctx, uint32_t* io, size_t count, int k, bool shift, const uint32_t* src) {
  Plan pl;
  if (force_levels()) return false;
  if (!make_plan(k, pl)) return false;
  if (count == 0) return true;
  size_t n = (size_t)1 << k;
  if ((count * n) % TILE != 0) return false;          // tiny batches of tiny columns: level path
  if (src == io) src = nullptr;
  if (src && pl.n_s == 0) { ZKB_CUDA(cudaMemcpyAsync(io, src, count * n * 4, cudaMemcpyDeviceToDevice, ctx->stream)); src = nullptr; }
  NttTables* t = ntt_tables(ctx);
  TwiddleRef tw{t->d_hi, t->d_lo};
  const uint2* twl = level_table(ctx, true);
  const Fp scale = inv(Fp::from((uint32_t)1 << k));
  const Fp three = Fp::from(3);
  // per-pass parameters
  struct SPass { int a, s_log; const uint32_t* table; uint32_t shift_g; const uint2* ttab; } sp[2];
  int consumed = 0;      // bits of the coefficient index already assigned (weight of the next pass = 2^consumed)
  for (int i = 0; i < pl.n_s; ++i) {
    int a = pl.a_s[i];
    sp[i].a = a; sp[i].s_log = k - consumed - a;
    Fp base = shift ? pow(three, (uint64_t)1 << consumed) : Fp::one();
    Fp mult = i == 0 ? scale : Fp::one();
    bool need_table = shift || i == 0;
    sp[i].table = need_table ? shift_table(ctx, std::max(a - 4, 0), base, mult) : nullptr;
    sp[i].shift_g = shift ? pow(base, (uint64_t)1 << (a - 4)).v : R_MOD_P;
    sp[i].ttab = s_pass_table(ctx, a, sp[i].s_log, true, base, mult);
    consumed += a;
  }
  const uint32_t* c_table = nullptr;
  int epi = EPI_NONE;
  if (shift) { c_table = shift_table(ctx, pl.a_c, pow(three, (uint64_t)1 << consumed), pl.n_s == 0 ? scale : Fp::one()); epi = EPI_TABLE; }
  else if (pl.n_s == 0) epi = EPI_SCALE;
  size_t bc = batch_columns(count, n * 4);
  for (size_t c0 = 0; c0 < count; c0 += bc) {
    size_t cols = std::min(bc, count - c0);
    uint32_t* p = io + c0 * n;
    size_t total = cols * n;
    for (int i = 0; i < pl.n_s; ++i) dispatch_s<true>(ctx, sp[i].a, p, total, sp[i].s_log, twl, tw, sp[i].table, sp[i].shift_g, sp[i].ttab, (i == 0 && src) ? src + c0 * n : nullptr);
    if (epi == EPI_TABLE) dispatch_c_inv<EPI_TABLE>(ctx, pl.a_c, p, total, twl, c_table, 0);
    else if (epi == EPI_SCALE) dispatch_c_inv<EPI_SCALE>(ctx, pl.a_c, p, total, twl, nullptr, scale.v);
    else dispatch_c_inv<EPI_NONE>(ctx, pl.a_c, p, total, twl, nullptr, 0);
  }
  return true;
}

bool ntt_forward_tiled(zkb_ctx* ctx, uint32_t* out, const uint32_t* in, size_t count, int k, int expand_bits) {
  Plan pl;
  if (force_levels()) return false;
  if (!make_plan(k, pl)) return false;
  if (expand_bits != 0 && expand_bits != 2) return false;
  if (expand_bits > pl.a_c - 0 || pl.a_c < 4) return false;
  if (count == 0) return true;
  size_t n = (size_t)1 << k;
  if ((count * n) % TILE != 0) return false;
  NttTables* t = ntt_tables(ctx);
  TwiddleRef tw{t->d_hi, t->d_lo};
  const uint2* twl = level_table(ctx, false);
  size_t n_in = n >> expand_bits;
  size_t bc = batch_columns(count, (n + (out == in ? 0 : n_in)) * 4);
  for (size_t c0 = 0; c0 < count; c0 += bc) {
    size_t cols = std::min(bc, count - c0);
    uint32_t* o = out + c0 * n;
    const uint32_t* i_ = in + c0 * n_in;
    size_t total = cols * n;
    if (expand_bits == 2) dispatch_c_fwd<2>(ctx, pl.a_c, o, i_, total, twl);
    else dispatch_c_fwd<0>(ctx, pl.a_c, o, i_, total, twl);
    // strided passes, innermost first: pass i works inside blocks of 2^(a_c + a_s[n_s-1] + ... + a_s[i]) elements
    int s_log = pl.a_c;
    for (int i = pl.n_s - 1; i >= 0; --i) {
      const uint2* t2 = s_pass_table2(ctx, pl.a_s[i], s_log);
      if (t2) dispatch_s<false>(ctx, pl.a_s[i], o, total, s_log, twl, tw, nullptr, R_MOD_P, t2, nullptr, 2);
      else dispatch_s<false>(ctx, pl.a_s[i], o, total, s_log, twl, tw, nullptr, R_MOD_P, s_pass_table(ctx, pl.a_s[i], s_log, false, Fp::one(), Fp::one()));
      s_log += pl.a_s[i];
    }
  }
  return true;
}

}  // namespace zkb
