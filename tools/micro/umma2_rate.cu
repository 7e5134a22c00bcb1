// Microbenchmark (NOT yet run on hardware -- written at the end of round 1 for round 2): clocks per tcgen05.mma with
// cta_group::2 (a CTA pair, M = 256 = 128 rows per CTA, kind::f16, K = 16, SS mode, SWIZZLE_64B K-major operands) as a
// function of N.  Question it answers: does the small-N cost floor of the single-CTA MMA (45 clk for N <= 32, 48 @ 64,
// profiles/r1_micro.md) apply per PAIR instruction?  If an M = 256, N = 64 MMA still costs ~48 clk, the Cout = 32 trunk
// convs (wide mode, N = 64 + N = 32 per K step) get twice the pixels per floor-bound instruction.
// In cta_group::2 each CTA holds its own 128 rows of A and HALF of B (N/2 rows) at the same shared-memory offsets; the
// leader CTA (cluster rank 0) issues, the accumulator lives in both CTAs' TMEM (same columns), and the commit is multicast
// to the barrier of both CTAs.  Run under `timeout 30`: every wait is bounded and traps instead of hanging.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t sbo_bytes) {
  return (uint64_t)((addr & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)4 << 61);
}
// kind::f16 instruction descriptor: D = f32, A = B = bf16, K-major, N, M (256 for the pair)
__device__ __forceinline__ uint32_t make_idesc(int n, int m) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) k(int n, int iters, long long* out) {
  extern __shared__ unsigned char smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  __shared__ uint32_t tmem_slot; __shared__ unsigned long long bar;
  uint32_t rank; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem_raw)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1)); asm volatile("fence.mbarrier_init.release.cluster;"); }
  if (threadIdx.x < 32) {   // one warp of EACH CTA of the pair takes part in the pair allocation
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;"); __syncthreads(); cluster_sync_all(); asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = tmem_slot;
  long long cyc = 0;
  if (rank == 0 && threadIdx.x == 0) {
    const uint32_t idesc = make_idesc(n, 256);
    const uint64_t bd = make_desc(base + 96 * 1024, 512);
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const uint64_t ad = make_desc(base + (uint32_t)(j * 128 * 64), 512);
        const uint32_t d = tmem + (uint32_t)((j & 1) * 256);
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                     ::"r"(d), "l"(ad), "l"(bd), "r"(idesc), "r"(1u) : "memory");
      }
    }
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(&bar)), "h"((uint16_t)3) : "memory");
    cyc = -t0;
  }
  if (threadIdx.x == 0) {   // both CTAs wait for the multicast commit (bounded)
    uint32_t ok = 0; const long long w0 = clock64();
    while (!ok) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
      if (clock64() - w0 > 4000000000LL) { printf("umma2_rate: timeout (block %d)\n", blockIdx.x); __trap(); }
    }
    if (rank == 0) { cyc += clock64(); if (blockIdx.x == 0) out[0] = cyc; }
  }
  asm volatile("tcgen05.fence::before_thread_sync;"); __syncthreads(); cluster_sync_all(); asm volatile("tcgen05.fence::after_thread_sync;");
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tmem));
}
int main() {
  long long* d; cudaMalloc(&d, 8);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int iters = 2000;
  for (int n : {32, 64, 96, 128, 192, 256}) {   // N % 16 == 0 for M = 256
    k<<<148, 128, 200 * 1024>>>(n, iters, d);
    long long c = 0; cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
    cudaError_t e = cudaGetLastError();
    printf("cta_group::2 M=256 N=%3d: %.1f clk/MMA (single-CTA M=128 costs: 45.5 @ <=32, 48 @ 64, 56 @ 96, N/2 above)  %s\n", n, (double)c / (iters * 8.0),
           e ? cudaGetErrorString(e) : "");
    if (e) break;
  }
  return 0;
}
