// Microbenchmark: HBM read bandwidth when each "pixel" contributes a 256-byte run and pixels are STRIDE bytes apart
// (a 64-channel fp32 slice of an NHWC tensor with 64 .. 1024 channels), 3.28 M pixels.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void rd(const float4* __restrict__ src, long long npix, long long stride16, float* out) {
  float acc = 0.f;
  const int lane16 = threadIdx.x & 15;
  for (long long p = (long long)blockIdx.x * (blockDim.x >> 4) + (threadIdx.x >> 4); p < npix; p += (long long)gridDim.x * (blockDim.x >> 4)) {
    const float4 v = __ldcs(src + p * stride16 + lane16);
    acc += v.x + v.y + v.z + v.w;
  }
  if (acc == 123.456f) out[0] = acc;
}
int main() {
  const long long npix = 3276800;
  float4* buf; cudaMalloc(&buf, npix * 4096 + 4096); cudaMemset(buf, 0, npix * 4096);
  float* out; cudaMalloc(&out, 4);
  for (int stride : {256, 512, 1024, 2048, 4096}) {
    rd<<<148 * 8, 256>>>(buf, npix, stride / 16, out); cudaDeviceSynchronize();
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0); rd<<<148 * 8, 256>>>(buf, npix, stride / 16, out); cudaEventRecord(e1); cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("pixel stride %4d B: %.3f ms, %.0f GB/s of useful bytes\n", stride, ms, npix * 256.0 / ms / 1e6);
  }
  return 0;
}
