// microbench.cu — INT-pipe (XOR+POPC+IADD) issue-rate microbenchmark.  SURVEY.md §8d: the Hamming k-NN
// kernel is bound by the POPC pipe, whose peak is not in MEASURED_PEAKS.json, so it is measured here and
// used as the second roofline denominator next to HBM bandwidth.
#include "cvb_internal.cuh"

namespace {
__global__ void __launch_bounds__(256) popc_kernel(int iters, uint32_t seed, uint32_t* out) {
  uint32_t x[8], acc[8];
#pragma unroll
  for (int i = 0; i < 8; i++) {
    x[i] = seed * (threadIdx.x + 1) + i * 0x9E3779B9u;
    acc[i] = 0;
  }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
      acc[i] += __popc(x[i] ^ (uint32_t)it);
      x[i] += acc[i];  // keeps the chain alive without adding POPC work
    }
  }
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) s += acc[i];
  if (s == 0xFFFFFFFFu) out[0] = s;  // never true in practice; defeats dead-code elimination
}
}  // namespace

extern "C" int cvb_microbench_popc(cvb_ctx* ctx, int iters, double* gpopc_per_s) {
  if (!ctx || !gpopc_per_s || iters <= 0) return CVB_ERR_INVALID;
  uint32_t* d = (uint32_t*)cvb_ws(ctx, WS_MISC, 256);
  if (!d) return CVB_ERR_CUDA;
  const int blocks = ctx->sm_count * 8;
  cudaEvent_t e0, e1;
  CVB_CUDA(ctx, cudaEventCreate(&e0));
  CVB_CUDA(ctx, cudaEventCreate(&e1));
  popc_kernel<<<blocks, 256, 0, ctx->stream>>>(iters / 4 + 1, 12345u, d);  // warm-up
  CVB_CHECK_LAUNCH(ctx);
  CVB_CUDA(ctx, cudaEventRecord(e0, ctx->stream));
  popc_kernel<<<blocks, 256, 0, ctx->stream>>>(iters, 12345u, d);
  CVB_CHECK_LAUNCH(ctx);
  CVB_CUDA(ctx, cudaEventRecord(e1, ctx->stream));
  CVB_CUDA(ctx, cudaEventSynchronize(e1));
  float ms = 0.f;
  CVB_CUDA(ctx, cudaEventElapsedTime(&ms, e0, e1));
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  *gpopc_per_s = (double)blocks * 256.0 * (double)iters * 8.0 / (ms * 1e-3) / 1e9;
  return CVB_OK;
}
