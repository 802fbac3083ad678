// poseidon_254 hash suite on the device: hash_fold / merkle_build / hash_rows (first slice).
//
// Stand-ins for risc0-zkp 1.2.5 `Hal::{hash_rows, hash_fold}` under `Poseidon254HashSuite` (CudaHal binds sppark_poseidon254_rows /
// _fold; un-vendored).  hash_pair(a, b) = poseidon([0, fr(a), fr(b)])[0] -- circomlib's 2-input Poseidon, pinned by its public known
// answers (poseidon254.cuh).  hash_rows packs the BabyBear elements of a row into field elements and absorbs them two at a time;
// the PACKING RULE below is provisional: upstream's is not recoverable offline (it is recalled as "eight elements per 254-bit word",
// not in detail), so only the permutation and hash_pair are claimed to match risc0.
//   provisional rule: canonical values e_k (decoded from Montgomery), 8 per word, word = sum_k e_k 2^(31 k) (248 bits);
//   overwrite-mode sponge of rate 2 in cells 1..2, zero padding, one permutation per two words (and one for an empty row);
//   the digest is cell 0.
// One thread per hash; ~830 Montgomery products of 256 bits per permutation.
#include "common.cuh"
#include "poseidon254.cuh"
#include <mutex>
#include <set>

namespace zkb {

__constant__ p254::Consts c_p254;

namespace {

// host: the reference Grain LFSR (the same generator as oracle/poseidon2.hpp, with n = 254, t = 3, R_F = 8, R_P = 57)
struct Grain {
  uint8_t s[80];
  Grain() {
    int k = 0;
    auto put = [&](uint32_t v, int w) { for (int i = w - 1; i >= 0; --i) s[k++] = (v >> i) & 1u; };
    put(1, 2); put(0, 4); put(254, 12); put(p254::T, 12); put(p254::RF, 10); put(p254::RP, 10);
    for (int i = 0; i < 30; ++i) s[k++] = 1;
    for (int i = 0; i < 160; ++i) clock();
  }
  uint8_t clock() {
    uint8_t n = s[62] ^ s[51] ^ s[38] ^ s[23] ^ s[13] ^ s[0];
    memmove(s, s + 1, 79); s[79] = n; return n;
  }
  uint8_t bit() { for (;;) { uint8_t a = clock(), b = clock(); if (a) return b; } }
  p254::Fr bits254() {       // 254 bits, most significant first
    p254::Fr x = {{0, 0, 0, 0, 0, 0, 0, 0}};
    for (int i = 253; i >= 0; --i) if (bit()) x.l[i / 32] |= 1u << (i % 32);
    return x;
  }
};
p254::Fr fr_pow(p254::Fr base, const uint32_t* e) {      // Montgomery in / out
  p254::Fr r = p254::one_m();
  for (int i = 255; i >= 0; --i) { r = p254::mul(r, r); if ((e[i / 32] >> (i % 32)) & 1u) r = p254::mul(r, base); }
  return r;
}
p254::Fr fr_inv(const p254::Fr& x) {
  uint32_t e[8] = ZKB_P254_LIMBS; e[0] -= 2;             // p - 2
  return fr_pow(x, e);
}
const p254::Consts& host_consts() {
  static p254::Consts k;
  static std::once_flag once;
  std::call_once(once, [] {
    Grain g;
    for (int i = 0; i < p254::N_RC;) { p254::Fr v = g.bits254(); if (!p254::geq_p(v.l)) k.rc[i++] = p254::to_mont(v); }      // rejection sampling
    // Cauchy matrix 1 / (x_i + y_j) from six further field elements (reduced, not rejected), all distinct, no zero denominators
    for (;;) {
      p254::Fr xy[6]; bool ok = true;
      for (int i = 0; i < 6; ++i) { xy[i] = g.bits254(); while (p254::geq_p(xy[i].l)) p254::sub_p(xy[i].l); }
      for (int i = 0; i < 6 && ok; ++i) for (int j = 0; j < i; ++j) if (!memcmp(xy[i].l, xy[j].l, 32)) ok = false;
      if (!ok) continue;
      for (int i = 0; i < 3 && ok; ++i) for (int j = 0; j < 3; ++j) {
        p254::Fr d = p254::add(p254::to_mont(xy[i]), p254::to_mont(xy[3 + j]));
        bool zero = true; for (int q = 0; q < 8; ++q) zero = zero && d.l[q] == 0;
        if (zero) { ok = false; break; }
        k.mds[i][j] = fr_inv(d);
      }
      if (ok) break;
    }
    // start-up self-check against circomlib's public known answer poseidon([1, 2])
    p254::Fr s[3] = {{{0, 0, 0, 0, 0, 0, 0, 0}}, p254::to_mont(p254::Fr{{1, 0, 0, 0, 0, 0, 0, 0}}), p254::to_mont(p254::Fr{{2, 0, 0, 0, 0, 0, 0, 0}})};
    p254::permute(s, k);
    uint32_t out[8]; p254::fr_to_digest(out, s[0]);
    static const uint32_t KAT[8] = {0x4417189au, 0x9e19607au, 0x74324551u, 0x2a3617f2u, 0x9662e9cfu, 0x3df64c6bu, 0xe7d69041u, 0x115cc0f5u};
    if (memcmp(out, KAT, 32)) throw Error("zkb200: poseidon_254 constants failed their known-answer check (poseidon([1,2]))");
  });
  return k;
}
void upload_consts(zkb_ctx* ctx) {
  static std::mutex mu; static std::set<int> done;
  std::lock_guard<std::mutex> lock(mu);
  if (done.count(ctx->device)) return;
  ZKB_CUDA(cudaMemcpyToSymbol(c_p254, &host_consts(), sizeof(p254::Consts)));
  done.insert(ctx->device);
}

}  // namespace

__global__ void __launch_bounds__(128) k_p254_hash_fold(uint32_t* __restrict__ nodes, size_t in_base, size_t out_base, size_t count) {
  const size_t i = (size_t)blockIdx.x * 128 + threadIdx.x;
  if (i >= count) return;
  const uint32_t* in = nodes + (in_base + 2 * i) * 8;
  p254::Fr s[3];
#pragma unroll
  for (int q = 0; q < 8; ++q) s[0].l[q] = 0;
  s[1] = p254::digest_to_fr(in); s[2] = p254::digest_to_fr(in + 8);
  p254::permute(s, c_p254);
  p254::fr_to_digest(nodes + (out_base + i) * 8, s[0]);
}
// provisional packing rule: see the header of this file
__global__ void __launch_bounds__(128) k_p254_hash_rows(uint32_t* __restrict__ out, const uint32_t* __restrict__ matrix, size_t rows, uint32_t cols) {
  const size_t r = (size_t)blockIdx.x * 128 + threadIdx.x;
  if (r >= rows) return;
  p254::Fr s[3];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int q = 0; q < 8; ++q) s[i].l[q] = 0;
  uint32_t c = 0; int cell = 0; bool mixed_any = false;
  while (c < cols) {
    p254::Fr w = {{0, 0, 0, 0, 0, 0, 0, 0}};
    for (int k = 0; k < 8 && c < cols; ++k, ++c) {
      const uint32_t e = mont_mul(1u, __ldg(matrix + (size_t)c * rows + r));      // canonical value < 2^31
      const int bit = 31 * k;
      w.l[bit / 32] |= e << (bit % 32);
      if (bit % 32 > 1) w.l[bit / 32 + 1] |= e >> (32 - bit % 32);
    }
    s[1 + cell] = p254::to_mont(w);
    if (++cell == 2) { p254::permute(s, c_p254); cell = 0; mixed_any = true; }
  }
  if (cell != 0 || !mixed_any) {
    if (cell == 1) for (int q = 0; q < 8; ++q) s[2].l[q] = 0;
    else if (!mixed_any && cell == 0) { for (int q = 0; q < 8; ++q) s[1].l[q] = s[2].l[q] = 0; }
    p254::permute(s, c_p254);
  }
  p254::fr_to_digest(out + r * 8, s[0]);
}

void p254_hash_fold(zkb_ctx* ctx, uint32_t* nodes, size_t input_size, size_t output_size) {
  if (!output_size) return;
  upload_consts(ctx);
  k_p254_hash_fold<<<grid_for(output_size, 128), 128, 0, ctx->stream>>>(nodes, input_size, output_size, output_size);
  launched(ctx);
}

}  // namespace zkb

using namespace zkb;

extern "C" {

zkb_err zkb_poseidon254_hash_fold(zkb_ctx* ctx, void* d_nodes, size_t input_size, size_t output_size) {
  ZKB_API_BEGIN
  use(ctx);
  ZKB_REQUIRE(d_nodes && aligned16(d_nodes), "null or misaligned node buffer");
  ZKB_REQUIRE(input_size == 2 * output_size, "hash_fold: input_size must be 2 * output_size");
  p254_hash_fold(ctx, (uint32_t*)d_nodes, input_size, output_size);
  ZKB_API_END
}
zkb_err zkb_poseidon254_merkle_build(zkb_ctx* ctx, void* d_nodes, size_t rows) {
  ZKB_API_BEGIN
  use(ctx);
  ZKB_REQUIRE(d_nodes && aligned16(d_nodes), "null or misaligned node buffer");
  ZKB_REQUIRE(rows >= 1 && (rows & (rows - 1)) == 0, "merkle_build: rows must be a power of two");
  for (size_t in = rows; in >= 2; in /= 2) p254_hash_fold(ctx, (uint32_t*)d_nodes, in, in / 2);
  ZKB_API_END
}
zkb_err zkb_poseidon254_hash_rows(zkb_ctx* ctx, void* d_out_digests, const void* d_matrix, size_t rows, size_t cols) {
  ZKB_API_BEGIN
  use(ctx);
  ZKB_REQUIRE(d_out_digests && (d_matrix || cols == 0), "null buffer");
  ZKB_REQUIRE(cols < (1u << 24), "too many columns");
  if (rows) {
    upload_consts(ctx);
    k_p254_hash_rows<<<grid_for(rows, 128), 128, 0, ctx->stream>>>((uint32_t*)d_out_digests, (const uint32_t*)d_matrix, rows, (uint32_t)cols);
    launched(ctx);
  }
  ZKB_API_END
}
/* host-only: one permutation of three canonical field elements (24 little-endian words in, 24 out) -- the known-answer hook */
zkb_err zkb_poseidon254_permute_host(const uint32_t* h_in, uint32_t* h_out) {
  ZKB_API_BEGIN
  ZKB_REQUIRE(h_in && h_out, "null argument");
  p254::Fr s[3];
  for (int i = 0; i < 3; ++i) s[i] = p254::digest_to_fr(h_in + 8 * i);
  p254::permute(s, host_consts());
  for (int i = 0; i < 3; ++i) p254::fr_to_digest(h_out + 8 * i, s[i]);
  ZKB_API_END
}

}  // extern "C"
