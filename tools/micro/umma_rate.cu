// Microbenchmark: issue rate of tcgen05.mma (cta_group::1) for kind::i8 / kind::f8f6f4, N = 128 / 256, A from shared memory or
// tensor memory.  One CTA per SM, one warp, garbage operands; reports cycles per instruction.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_rate umma_rate.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t a, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((a >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n.reg .pred P;\nelect.sync _|P, 0xffffffff;\nselp.b32 %0, 1, 0, P;\n}\n" : "=r"(pred));
  return pred != 0;
}
template <int KIND, bool A_TMEM>
__device__ __forceinline__ void mma(uint32_t d, uint32_t a_t, uint64_t ad, uint64_t bd, uint32_t idesc) {
  if (A_TMEM) {
    if (KIND == 0) asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, 1, 0;\ntcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n}\n" ::"r"(d), "r"(a_t), "l"(bd), "r"(idesc) : "memory");
    else asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, 1, 0;\ntcgen05.mma.cta_group::1.kind::f8f6f4 [%0], [%1], %2, %3, p;\n}\n" ::"r"(d), "r"(a_t), "l"(bd), "r"(idesc) : "memory");
  } else {
    if (KIND == 0) asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, 1, 0;\ntcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}\n" ::"r"(d), "l"(ad), "l"(bd), "r"(idesc) : "memory");
    else asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, 1, 0;\ntcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n}\n" ::"r"(d), "l"(ad), "l"(bd), "r"(idesc) : "memory");
  }
}
template <int KIND, int N, bool A_TMEM>
__global__ void __launch_bounds__(32, 1) k(int iters, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t tb;
  __shared__ uint64_t bar;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = threadIdx.x; i < (128 + N) * 256 / 4; i += 32) reinterpret_cast<uint32_t*>(smem)[i] = 0x01010101u;
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_addr(&tb)));
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncwarp();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tb;
  // idesc: D format at [4,6): s32 = 2 (i8) / f32 = 1 (f8f6f4); A/B formats 0 (u8 / e4m3); N >> 3 at [17,23); M >> 4 at [24,29)
  const uint32_t idesc = ((KIND == 0 ? 2u : 1u) << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  const uint64_t ad = make_desc(smem_addr(smem), 128, 2048), bd = make_desc(smem_addr(smem) + 128 * 256, 128, 2048);
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
    if (elect_one()) {
#pragma unroll
      for (int kk = 0; kk < 8; kk++) mma<KIND, A_TMEM>(tmem + (it & 1) * (N == 256 ? 0 : 128), tmem + 448 + kk * 8, ad + kk * 16, bd + kk * 16, idesc);
    }
    __syncwarp();
  }
  if (elect_one()) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_addr(&bar)) : "memory");
  __syncwarp();
  asm volatile("{\n.reg .pred p;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra D;\nbra W;\nD:\n}\n" ::"r"(smem_addr(&bar)) : "memory");
  long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) *out = t1 - t0;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncwarp();
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
}
template <int KIND, int N, bool A_TMEM>
void run(const char* name) {
  long long* d; cudaMalloc(&d, 8);
  const int smem = (128 + 256) * 256 + 1024, iters = 2000;
  cudaFuncSetAttribute(k<KIND, N, A_TMEM>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int grid : {1, 148}) {
    k<KIND, N, A_TMEM><<<grid, 32, smem>>>(iters, d);
    k<KIND, N, A_TMEM><<<grid, 32, smem>>>(iters, d);
    cudaError_t e = cudaDeviceSynchronize();
    long long h = 0; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    printf("%-34s grid %3d: %7.1f cycles per MMA (M=128, N=%d, K=32 B)  -> %6.0f MAC/clk/SM  %s\n", name, grid, (double)h / (iters * 8.0), N,
           128.0 * N * 32 * iters * 8 / (double)h, e == cudaSuccess ? "" : cudaGetErrorString(e));
  }
  cudaFree(d);
}
int main() {
  run<0, 128, false>("i8      A smem  N=128");
  run<0, 256, false>("i8      A smem  N=256");
  run<0, 128, true>("i8      A tmem  N=128");
  run<0, 256, true>("i8      A tmem  N=256");
  run<1, 128, false>("f8f6f4  A smem  N=128");
  run<1, 256, false>("f8f6f4  A smem  N=256");
  run<1, 256, true>("f8f6f4  A tmem  N=256");
  return 0;
}
