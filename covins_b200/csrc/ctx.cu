// ctx.cu — context lifetime, error reporting, grow-only workspaces.
#include <stdarg.h>

#include "cvb_internal.cuh"

int cvb_fail(cvb_ctx* ctx, int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (ctx) ctx->err = buf;
  return code;
}

void* cvb_ws(cvb_ctx* ctx, int slot, size_t bytes) {
  cvb_buf& b = ctx->ws[slot];
  if (bytes <= b.cap && b.p) return b.p;
  size_t want = bytes < 256 ? 256 : bytes;
  if (b.p) {
    // previous users of the buffer may still be in flight on the ctx stream
    cudaStreamSynchronize(ctx->stream);
    cudaFree(b.p);
    b.p = nullptr;
    b.cap = 0;
    want = want + want / 4;  // amortise regrowth
  }
  cudaError_t e = cudaMalloc(&b.p, want);
  if (e != cudaSuccess) {
    cvb_fail(ctx, CVB_ERR_CUDA, "cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e));
    b.p = nullptr;
    return nullptr;
  }
  b.cap = want;
  return b.p;
}

void* cvb_pinned(cvb_ctx* ctx, size_t bytes) {
  if (bytes <= ctx->h_pin_cap && ctx->h_pin) return ctx->h_pin;
  if (ctx->h_pin) {
    cudaStreamSynchronize(ctx->stream);
    cudaFreeHost(ctx->h_pin);
    ctx->h_pin = nullptr;
    ctx->h_pin_cap = 0;
  }
  size_t want = bytes < 4096 ? 4096 : bytes + bytes / 4;
  cudaError_t e = cudaMallocHost(&ctx->h_pin, want);
  if (e != cudaSuccess) {
    cvb_fail(ctx, CVB_ERR_CUDA, "cudaMallocHost(%zu) failed: %s", want, cudaGetErrorString(e));
    ctx->h_pin = nullptr;
    return nullptr;
  }
  ctx->h_pin_cap = want;
  return ctx->h_pin;
}

extern "C" {

int cvb_version(void) { return 100; }

int cvb_ctx_create(int device, cvb_ctx** out) {
  if (!out) return CVB_ERR_INVALID;
  *out = nullptr;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n <= 0 || device < 0 || device >= n) return CVB_ERR_CUDA;  // no CPU fallback
  if (cudaSetDevice(device) != cudaSuccess) return CVB_ERR_CUDA;
  cvb_ctx* c = new cvb_ctx();
  c->device = device;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) {
    delete c;
    return CVB_ERR_CUDA;
  }
  c->sm_count = prop.multiProcessorCount;
  int prio_lo = 0, prio_hi = 0;   // the ctx stream carries the critical path (diagonal tiles, panels): greatest priority
  cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
  if (cudaStreamCreateWithPriority(&c->stream, cudaStreamNonBlocking, prio_hi) != cudaSuccess) {
    delete c;
    return CVB_ERR_CUDA;
  }
  {   // keep memory freed by the stream-ordered allocator in the pool (problem set-up re-uses it)
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
      unsigned long long thr = ~0ull;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    }
  }
  *out = c;
  return CVB_OK;
}

void cvb_ba_free(cvb_ctx* ctx);  // ba_api.cu

int cvb_ctx_destroy(cvb_ctx* ctx) {
  if (!ctx) return CVB_OK;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  cvb_ba_free(ctx);
  for (auto& b : ctx->ws)
    if (b.p) cudaFree(b.p);
  if (ctx->h_pin) cudaFreeHost(ctx->h_pin);
  cudaStreamDestroy(ctx->stream);
  delete ctx;
  return CVB_OK;
}

const char* cvb_last_error(const cvb_ctx* ctx) { return ctx ? ctx->err.c_str() : "null ctx"; }

int cvb_ctx_sync(cvb_ctx* ctx) {
  if (!ctx) return CVB_ERR_INVALID;
  CVB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return CVB_OK;
}

int64_t cvb_launch_count(const cvb_ctx* ctx) { return ctx ? ctx->launches : 0; }

}  // extern "C"
