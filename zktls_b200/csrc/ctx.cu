// Lifecycle + memory entry points of libzkb200 (include/zkb200.h).  Stand-ins for risc0-zkp
// `Hal::{alloc_*, copy_from_*}` / `Buffer::{view, view_mut, get_at, to_vec}` and for the cust context/DeviceBuffer
// handling inside CudaHal.
#include "common.cuh"

using namespace zkb;

extern "C" {

const char* zkb_version(void) { return "zkb200 0.1 (sm_100a)"; }
void zkb_free_error(const char* err) { free((void*)err); }

static zkb_err init_common(int device, cudaStream_t borrowed, bool borrow, zkb_ctx** out) {
  ZKB_API_BEGIN
  ZKB_REQUIRE(out != nullptr, "null out pointer");
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0)
    throw Error(std::string("zkb200: no CUDA device available (") + cudaGetErrorString(e) + "); this backend has no CPU fallback");
  ZKB_REQUIRE(device >= 0 && device < count, "device ordinal out of range");
  ZKB_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  ZKB_CUDA(cudaGetDeviceProperties(&prop, device));
  ZKB_REQUIRE(prop.major >= 10, "zkb200 kernels are built for sm_100a only (device is sm_" + std::to_string(prop.major) + std::to_string(prop.minor) + ")");
  zkb_ctx* ctx = new zkb_ctx();
  ctx->device = device;
  ctx->sm_count = prop.multiProcessorCount;
  if (borrow) { ctx->stream = borrowed; ctx->owns_stream = false; }
  else { ZKB_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)); ctx->owns_stream = true; }
  {
    cudaMemPoolProps props = {};
    props.allocType = cudaMemAllocationTypePinned;
    props.handleTypes = cudaMemHandleTypeNone;
    props.location.type = cudaMemLocationTypeDevice;
    props.location.id = device;
    ZKB_CUDA(cudaMemPoolCreate(&ctx->pool, &props));
    uint64_t threshold = UINT64_MAX;      // keep freed blocks: a segment re-uses the same ~8 GB every time
    ZKB_CUDA(cudaMemPoolSetAttribute(ctx->pool, cudaMemPoolAttrReleaseThreshold, &threshold));
  }
  ZKB_CUDA(cudaEventCreate(&ctx->ev0));
  ZKB_CUDA(cudaEventCreate(&ctx->ev1));
  ctx->staging_bytes = 1u << 20;
  ZKB_CUDA(cudaHostAlloc(&ctx->staging, ctx->staging_bytes, cudaHostAllocDefault));
  *out = ctx;
  ZKB_API_END
}
zkb_err zkb_init(int device, zkb_ctx** out) { return init_common(device, nullptr, false, out); }
zkb_err zkb_init_on_stream(int device, void* cuda_stream, zkb_ctx** out) { return init_common(device, (cudaStream_t)cuda_stream, true, out); }

zkb_err zkb_destroy(zkb_ctx* ctx) {
  ZKB_API_BEGIN
  if (!ctx) return nullptr;
  use(ctx);
  cudaStreamSynchronize(ctx->stream);
  ntt_tables_free(ctx);
  eval_jit_free(ctx);
  if (ctx->scratch) cudaFree(ctx->scratch);
  if (ctx->staging) cudaFreeHost(ctx->staging);
  cudaEventDestroy(ctx->ev0);
  cudaEventDestroy(ctx->ev1);
  if (ctx->pool) cudaMemPoolDestroy(ctx->pool);
  if (ctx->owns_stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
  ZKB_API_END
}
zkb_err zkb_sync(zkb_ctx* ctx) {
  ZKB_API_BEGIN
  use(ctx);
  ZKB_CUDA(cudaStreamSynchronize(ctx->stream));
  ZKB_API_END
}
zkb_err zkb_device_info(zkb_ctx* ctx, int* sm_count, int* cc_major, int* cc_minor, size_t* total_mem) {
  ZKB_API_BEGIN
  use(ctx);
  cudaDeviceProp prop;
  ZKB_CUDA(cudaGetDeviceProperties(&prop, ctx->device));
  if (sm_count) *sm_count = prop.multiProcessorCount;
  if (cc_major) *cc_major = prop.major;
  if (cc_minor) *cc_minor = prop.minor;
  if (total_mem) *total_mem = prop.totalGlobalMem;
  ZKB_API_END
}
zkb_err zkb_device_count(int* out) {
  ZKB_API_BEGIN
  ZKB_REQUIRE(out != nullptr, "null out pointer");
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) { cudaGetLastError(); n = 0; }      // no driver / no device: zero, not an error
  *out = n;
  ZKB_API_END
}
zkb_err zkb_kernel_launches(zkb_ctx* ctx, uint64_t* out) {
  ZKB_API_BEGIN
  ZKB_REQUIRE(ctx && out, "null argument");
  *out = ctx->launches;
  ZKB_API_END
}
zkb_err zkb_timer_start(zkb_ctx* ctx) {
  ZKB_API_BEGIN
  use(ctx);
  ZKB_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
  ZKB_API_END
}
zkb_err zkb_timer_stop(zkb_ctx* ctx, float* ms) {
  ZKB_API_BEGIN
  use(ctx);
  ZKB_REQUIRE(ms != nullptr, "null ms");
  ZKB_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
  ZKB_CUDA(cudaEventSynchronize(ctx->ev1));
  ZKB_CUDA(cudaEventElapsedTime(ms, ctx->ev0, ctx->ev1));
  ZKB_API_END
}

zkb_err zkb_alloc(zkb_ctx* ctx, size_t bytes, void** d_out) {
  ZKB_API_BEGIN
  use(ctx);
  ZKB_REQUIRE(d_out != nullptr, "null out pointer");
  ZKB_CUDA(cudaMalloc(d_out, bytes ? bytes : 16));
  ZKB_API_END
}
zkb_err zkb_free(zkb_ctx* ctx, void* d_ptr) {
  ZKB_API_BEGIN
  use(ctx);
  if (d_ptr) { ZKB_CUDA(cudaStreamSynchronize(ctx->stream)); ZKB_CUDA(cudaFree(d_ptr)); }
  ZKB_API_END
}
zkb_err zkb_host_alloc(zkb_ctx* ctx, size_t bytes, void** h_out) {
  ZKB_API_BEGIN
  use(ctx);
  ZKB_REQUIRE(h_out != nullptr, "null out pointer");
  ZKB_CUDA(cudaHostAlloc(h_out, bytes ? bytes : 16, cudaHostAllocDefault));
  ZKB_API_END
}
zkb_err zkb_host_free(zkb_ctx* ctx, void* h_ptr) {
  ZKB_API_BEGIN
  use(ctx);
  if (h_ptr) ZKB_CUDA(cudaFreeHost(h_ptr));
  ZKB_API_END
}
zkb_err zkb_memset0(zkb_ctx* ctx, void* d_ptr, size_t bytes) {
  ZKB_API_BEGIN
  use(ctx);
  if (bytes) ZKB_CUDA(cudaMemsetAsync(d_ptr, 0, bytes, ctx->stream));
  ZKB_API_END
}
zkb_err zkb_h2d(zkb_ctx* ctx, void* d_dst, const void* h_src, size_t bytes) {
  ZKB_API_BEGIN
  use(ctx);
  if (bytes) ZKB_CUDA(cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, ctx->stream));
  ZKB_API_END
}
zkb_err zkb_d2h(zkb_ctx* ctx, void* h_dst, const void* d_src, size_t bytes) {
  ZKB_API_BEGIN
  use(ctx);
  if (bytes) ZKB_CUDA(cudaMemcpyAsync(h_dst, d_src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  ZKB_CUDA(cudaStreamSynchronize(ctx->stream));
  ZKB_API_END
}
zkb_err zkb_d2d(zkb_ctx* ctx, void* d_dst, const void* d_src, size_t bytes) {
  ZKB_API_BEGIN
  use(ctx);
  if (bytes) ZKB_CUDA(cudaMemcpyAsync(d_dst, d_src, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
  ZKB_API_END
}

}  // extern "C"
