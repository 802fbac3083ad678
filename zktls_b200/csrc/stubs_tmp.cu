#include "common.cuh"
extern "C" {
zkb_err zkb_verify_segment(const uint32_t*, size_t, const uint32_t*, size_t) { return strdup("not implemented yet"); }
}
