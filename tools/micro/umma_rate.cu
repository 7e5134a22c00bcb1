// Microbenchmark: clocks per tcgen05.mma (cta_group::1, kind::f16, M=128, K=16, SS mode, SWIZZLE_64B K-major operands) as a
// function of N, and with the A operand shared or distinct between consecutive MMAs.  One CTA per SM, one issuing thread.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t sbo_bytes) {
  return (uint64_t)((addr & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)4 << 61);
}
__device__ __forceinline__ uint32_t make_idesc(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
__global__ void __launch_bounds__(128, 1) k(int n, int iters, int a_stride_rows, long long* out, int issuers) {
  extern __shared__ unsigned char smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  __shared__ uint32_t tmem_slot; __shared__ unsigned long long bar;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem_raw)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(issuers)); asm volatile("fence.mbarrier_init.release.cluster;"); }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;"); __syncthreads(); asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = tmem_slot;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if ((threadIdx.x & 31) == 0 && (threadIdx.x >> 5) < issuers) {
    const int wid = threadIdx.x >> 5;
    const uint32_t idesc = make_idesc(n);
    const uint64_t bd = make_desc(base + 96 * 1024, 512);
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const uint64_t ad = make_desc(base + (uint32_t)(j * a_stride_rows * 64), 512);
        const uint32_t d = tmem + (uint32_t)(wid * 128 + (j & 1) * 64);
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                     ::"r"(d), "l"(ad), "l"(bd), "r"(idesc), "r"(1u) : "memory");
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    uint32_t ok = 0;
    while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
    const long long t1 = clock64();
    if (blockIdx.x == 0 && wid == 0) out[0] = t1 - t0;
  }
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
}
int main() {
  long long* d; cudaMalloc(&d, 8);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int iters = 2000;
  for (int issuers : {1, 2, 4}) for (int n : {16, 32, 64, 96, 128, 192, 256}) {
    if (issuers > 1 && n > 64) continue;   // accumulator columns: issuers * 128 must fit 512
    const int stride = 128;
    k<<<148, 128, 200 * 1024>>>(n, iters, stride, d, issuers);
    long long c = 0; cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
    cudaError_t e = cudaGetLastError();
    printf("issuers %d, N=%3d: %.1f clk/MMA overall (math floor N/2 = %d)  %s\n", issuers, n, (double)c / (iters * 8.0 * issuers), n / 2, e ? cudaGetErrorString(e) : "");
  }
  return 0;
}
