#include "common.cuh"
extern "C" {
#define NI return strdup("not implemented yet");
zkb_err zkb_eval_check(zkb_ctx*, void*, const uint32_t*, size_t, const void*, const void*, const void*, const uint32_t*, const uint32_t*, const uint32_t*, int) { NI }
zkb_err zkb_prover_new(zkb_ctx*, const uint32_t*, size_t, zkb_prover**) { NI }
zkb_err zkb_prover_free(zkb_prover*) { NI }
zkb_err zkb_prover_segment_begin(zkb_prover*, int, const uint32_t*, const void*, const void*, int, uint32_t*) { NI }
zkb_err zkb_prover_segment_finish(zkb_prover*, const void*, int) { NI }
zkb_err zkb_prover_seal_words(zkb_prover*, size_t*) { NI }
zkb_err zkb_prover_seal_copy(zkb_prover*, uint32_t*) { NI }
zkb_err zkb_prover_root_count(zkb_prover*, size_t*) { NI }
zkb_err zkb_prover_roots_copy(zkb_prover*, uint32_t*) { NI }
zkb_err zkb_prove_segment(zkb_prover*, int, const uint32_t*, const void*, const void*, const void*, int) { NI }
zkb_err zkb_verify_segment(const uint32_t*, size_t, const uint32_t*, size_t) { NI }
}
