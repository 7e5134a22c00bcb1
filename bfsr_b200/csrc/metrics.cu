// Evaluation metrics of the reference's test drivers on the device:
//   PSNR  — LINF-LP/utils.py:132-151 (calc_psnr: plain / 'benchmark' luma + shave / 'div2k' shave)
//   SSIM  — LINF-LP/utils.py:154-193 (11x11 Gaussian window sigma 1.5, 'valid' region, float64, mean over channels).
//           SRFlow-LP's Measure.py:46-49 calls skimage's default SSIM -- 7x7 uniform window, sample covariance -- a different
//           definition: bfsr_metric_ssim_uniform below (same kernel, uniform window, cov_norm = n/(n-1)).
// Both accumulate in fp64 (the reference's SSIM is fp64; its PSNR is an fp32 mean whose rounding we do not reproduce).
#include "common.cuh"
#include "../../include/bfsr_b200.h"
#include <cmath>
#include <string>
#include <vector>

namespace { struct FreeAsync { void* p; cudaStream_t s; ~FreeAsync() { if (p) cudaFreeAsync(p, s); } }; }   // scratch released on every exit path

namespace bfsr {

__device__ __forceinline__ double block_sum(double v) {
  __shared__ double sh[32];
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  v = threadIdx.x < (blockDim.x >> 5) ? sh[threadIdx.x] : 0.0;
  if (threadIdx.x < 32) for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  return v;   // valid in thread 0
}

// sum of squared (optionally luma-weighted) differences over the shaved region
__global__ void psnr_kernel(const float* sr, const float* hr, int B, int C, int H, int W, int luma, int shave, float inv_range,
                            double* acc) {
  const int h = H - 2 * shave, w = W - 2 * shave, Ce = luma ? 1 : C;
  const long long n = (long long)B * Ce * h * w;
  double s = 0.0;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(e % w) + shave; long long t = e / w; const int y = (int)(t % h) + shave; t /= h;
    const int c = (int)(t % Ce); const long long b = t / Ce;
    float d;
    if (luma) {   // diff.mul(convert).sum(dim=1), convert = [65.738, 129.057, 25.064] / 256 (utils.py:137-140)
      const long long i = (b * C * H + y) * W + x, pl = (long long)H * W;
      const float d0 = (sr[i] - hr[i]) * inv_range, d1 = (sr[i + pl] - hr[i + pl]) * inv_range, d2 = (sr[i + 2 * pl] - hr[i + 2 * pl]) * inv_range;
      d = d0 * (65.738f / 256.f) + d1 * (129.057f / 256.f) + d2 * (25.064f / 256.f);
    } else {
      const long long i = ((b * C + c) * H + y) * W + x;
      d = (sr[i] - hr[i]) * inv_range;
    }
    s += (double)d * (double)d;
  }
  s = block_sum(s);
  if (threadIdx.x == 0) atomicAdd(acc, s);
}

// one thread per valid pixel of one channel plane; images are (C,H,W) fp32 scaled by `mul` (255 for [0,1] inputs).
// K x K window `win` (11x11 Gaussian: utils.py; 7x7 uniform: skimage's default), cov_norm = 1 or NP / (NP - 1) (sample covariance)
__global__ void ssim_kernel(const float* a, const float* b, int C, int H, int W, float mul, int K, double cov_norm,
                            const double* __restrict__ win, double* acc) {
  const int h = H - (K - 1), w = W - (K - 1);
  const long long n = (long long)C * h * w;
  const double C1 = (0.01 * 255) * (0.01 * 255), C2 = (0.03 * 255) * (0.03 * 255);
  double s = 0.0;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(e % w); const long long t = e / w; const int y = (int)(t % h); const int c = (int)(t / h);
    const float* pa = a + ((long long)c * H + y) * W + x;
    const float* pb = b + ((long long)c * H + y) * W + x;
    double m1 = 0, m2 = 0, s11 = 0, s22 = 0, s12 = 0;
    for (int i = 0; i < K; ++i)
      for (int j = 0; j < K; ++j) {
        const double wgt = win[i * K + j];
        const double u = (double)(pa[i * W + j] * mul), v = (double)(pb[i * W + j] * mul);
        m1 += wgt * u; m2 += wgt * v; s11 += wgt * u * u; s22 += wgt * v * v; s12 += wgt * u * v;
      }
    const double v1 = cov_norm * (s11 - m1 * m1), v2 = cov_norm * (s22 - m2 * m2), cov = cov_norm * (s12 - m1 * m2);
    s += ((2 * m1 * m2 + C1) * (2 * cov + C2)) / ((m1 * m1 + m2 * m2 + C1) * (v1 + v2 + C2));
  }
  s = block_sum(s);
  if (threadIdx.x == 0) atomicAdd(acc, s);
}

// MATLAB-compatible antialiased bicubic resize (LINF-LP/imresize.py:53-175; the LR-consistency metric of test.py:183-200):
// out[o] = sum_p w[o][p] * in[idx[o][p]] along one dimension, fp64; planes are (C, H, W)
template <typename TI>
__global__ void resize_dim_kernel(const TI* in, double* out, int C, int H, int W, int dim, int olen, int P, const double* wgt,
                                  const int* idx, int round_u8) {
  const int oH = dim == 0 ? olen : H, oW = dim == 1 ? olen : W;
  const long long n = (long long)C * oH * oW;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(e % oW); const long long t = e / oW; const int y = (int)(t % oH); const int c = (int)(t / oH);
    const int o = dim == 0 ? y : x;
    double acc = 0.0;
    for (int p = 0; p < P; ++p) {
      const int i = idx[o * P + p];
      const double v = dim == 0 ? (double)in[((long long)c * H + i) * W + x] : (double)in[((long long)c * H + y) * W + i];
      acc += wgt[o * P + p] * v;
    }
    // uint8 images: every pass ends with np.around(np.clip(., 0, 255)).astype(uint8) (imresize.py:108-110,122-124)
    out[e] = round_u8 ? rint(fmin(fmax(acc, 0.0), 255.0)) : acc;
  }
}
__global__ void f64_to_f32_kernel(const double* in, float* out, long long n) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e < n) out[e] = (float)in[e];
}

static double cubic(double x) {   // imresize.py:30-37
  const double a = std::fabs(x), a2 = a * a, a3 = a2 * a;
  return (a <= 1 ? 1.5 * a3 - 2.5 * a2 + 1 : 0.0) + ((1 < a && a <= 2) ? -0.5 * a3 + 2.5 * a2 - 4 * a + 2 : 0.0);
}
// imresize.py:40-64 (mirror padding through the `aux` index table; all-zero weight columns are kept: they add exact zeros)
static void contributions(int in_len, int out_len, double scale, std::vector<double>& w, std::vector<int>& ind, int& P) {
  const double kw = scale < 1 ? 4.0 / scale : 4.0;
  P = (int)std::ceil(kw) + 2;
  w.assign((size_t)out_len * P, 0.0); ind.assign((size_t)out_len * P, 0);
  const int aux_n = 2 * in_len;
  for (int o = 0; o < out_len; ++o) {
    const double u = (o + 1) / scale + 0.5 * (1 - 1 / scale);
    const double left = std::floor(u - kw / 2);
    double sum = 0;
    for (int p = 0; p < P; ++p) {
      const int i = (int)(left + p - 1);
      const double d = u - i - 1;
      const double v = scale < 1 ? scale * cubic(scale * d) : cubic(d);
      w[(size_t)o * P + p] = v; sum += v;
      int m = i % aux_n; if (m < 0) m += aux_n;
      ind[(size_t)o * P + p] = m < in_len ? m : aux_n - 1 - m;
    }
    for (int p = 0; p < P; ++p) w[(size_t)o * P + p] /= sum;
  }
}

}  // namespace bfsr

namespace bfsr { void set_last_error(const std::string& m); }   // capi.cu: message returned by bfsr_last_error()
using namespace bfsr;
extern "C" {

int bfsr_metric_psnr(const float* sr_dev, const float* hr_dev, int32_t B, int32_t C, int32_t H, int32_t W, int32_t mode,
                     int32_t scale, float rgb_range, double* psnr_out, void* stream) {
  try {
    BFSR_CHECK(sr_dev && hr_dev && psnr_out && B > 0 && C > 0 && H > 0 && W > 0, "psnr: bad arguments");
    BFSR_CHECK(mode >= 0 && mode <= 2, "psnr: mode must be 0 (none), 1 ('benchmark') or 2 ('div2k')");
    const int shave = mode ? scale : 0, luma = (mode == 1 && C > 1) ? 1 : 0;
    BFSR_CHECK(!luma || C == 3, "psnr('benchmark'): luma conversion needs 3 channels");
    BFSR_CHECK(H > 2 * shave && W > 2 * shave, "psnr: image smaller than the shaved border");
    cudaStream_t s = (cudaStream_t)stream;
    double* acc = nullptr;
    CUDA_OK(cudaMallocAsync((void**)&acc, 8, s));
    FreeAsync guard{acc, s};
    CUDA_OK(cudaMemsetAsync(acc, 0, 8, s));
    const long long n = (long long)B * (luma ? 1 : C) * (H - 2 * shave) * (W - 2 * shave);
    psnr_kernel<<<(int)((n + 255) / 256 > 1184 ? 1184 : (n + 255) / 256), 256, 0, s>>>(sr_dev, hr_dev, B, C, H, W, luma, shave, 1.f / rgb_range, acc);
    double sum = 0;
    CUDA_OK(cudaMemcpyAsync(&sum, acc, 8, cudaMemcpyDeviceToHost, s));
    CUDA_OK(cudaStreamSynchronize(s));
    *psnr_out = -10.0 * std::log10(sum / (double)n);
  } catch (const std::exception& ex) { set_last_error(ex.what()); return -1; }
  return 0;
}

static void ssim_run(const float* img1_dev, const float* img2_dev, int C, int H, int W, float mul, int K, const double* win_host,
                     double cov_norm, double* ssim_out, cudaStream_t s) {
  double* buf = nullptr;
  const int nw = K * K;
  CUDA_OK(cudaMallocAsync((void**)&buf, (nw + 1) * 8, s));
  FreeAsync guard{buf, s};
  CUDA_OK(cudaMemcpyAsync(buf, win_host, nw * 8, cudaMemcpyHostToDevice, s));
  CUDA_OK(cudaMemsetAsync(buf + nw, 0, 8, s));
  const long long n = (long long)C * (H - K + 1) * (W - K + 1);
  ssim_kernel<<<(int)((n + 127) / 128 > 2368 ? 2368 : (n + 127) / 128), 128, 0, s>>>(img1_dev, img2_dev, C, H, W, mul, K, cov_norm, buf, buf + nw);
  double sum = 0;
  CUDA_OK(cudaMemcpyAsync(&sum, buf + nw, 8, cudaMemcpyDeviceToHost, s));
  CUDA_OK(cudaStreamSynchronize(s));     // also keeps win_host alive until the copy has been made
  *ssim_out = sum / (double)n;
}

int bfsr_metric_ssim(const float* img1_dev, const float* img2_dev, int32_t C, int32_t H, int32_t W, float mul, double* ssim_out,
                     void* stream) {
  try {
    BFSR_CHECK(img1_dev && img2_dev && ssim_out && C > 0 && H > 10 && W > 10, "ssim: bad arguments (images must exceed the 11x11 window)");
    // cv2.getGaussianKernel(11, 1.5): exp(-(i-5)^2 / (2 sigma^2)) normalised to sum 1; window = outer product (utils.py:160-161)
    double k[11], ksum = 0, win[121];
    for (int i = 0; i < 11; ++i) { k[i] = std::exp(-(double)((i - 5) * (i - 5)) / (2.0 * 1.5 * 1.5)); ksum += k[i]; }
    for (int i = 0; i < 11; ++i) k[i] /= ksum;
    for (int i = 0; i < 11; ++i) for (int j = 0; j < 11; ++j) win[i * 11 + j] = k[i] * k[j];
    ssim_run(img1_dev, img2_dev, C, H, W, mul, 11, win, 1.0, ssim_out, (cudaStream_t)stream);
  } catch (const std::exception& ex) { set_last_error(ex.what()); return -1; }
  return 0;
}

int bfsr_metric_ssim_uniform(const float* img1_dev, const float* img2_dev, int32_t C, int32_t H, int32_t W, float mul,
                             int32_t win_size, int32_t sample_cov, double* ssim_out, void* stream) {
  try {
    BFSR_CHECK(img1_dev && img2_dev && ssim_out && C > 0, "ssim_uniform: bad arguments");
    BFSR_CHECK(win_size >= 3 && win_size <= 31 && (win_size & 1), "ssim_uniform: win_size must be odd, 3..31");
    BFSR_CHECK(H >= win_size && W >= win_size, "ssim_uniform: win_size exceeds image extent");
    std::vector<double> win((size_t)win_size * win_size, 1.0 / (double)(win_size * win_size));
    const double np = (double)(win_size * win_size);
    ssim_run(img1_dev, img2_dev, C, H, W, mul, win_size, win.data(), sample_cov ? np / (np - 1.0) : 1.0, ssim_out, (cudaStream_t)stream);
  } catch (const std::exception& ex) { set_last_error(ex.what()); return -1; }
  return 0;
}

static int imresize_run(const float* img_dev, int32_t C, int32_t H, int32_t W, double scale, float* out_dev, int32_t* out_h,
                        int32_t* out_w, void* stream, int round_u8) {
  try {
    BFSR_CHECK(C > 0 && H > 0 && W > 0 && scale > 0, "imresize: bad arguments");
    const int oh = (int)std::ceil(scale * H), ow = (int)std::ceil(scale * W);     // deriveSizeFromScale, imresize.py:6-10
    if (out_h) *out_h = oh;
    if (out_w) *out_w = ow;
    if (!out_dev) return 0;                                                      // size query
    BFSR_CHECK(img_dev, "imresize: null input");
    cudaStream_t s = (cudaStream_t)stream;
    std::vector<double> w0, w1; std::vector<int> i0, i1; int P0 = 0, P1 = 0;
    contributions(H, oh, scale, w0, i0, P0);
    contributions(W, ow, scale, w1, i1, P1);
    // one scratch block: [w0 | w1 | t0 | t1] doubles, then [i0 | i1] ints
    const size_t nd = w0.size() + w1.size() + (size_t)C * oh * W + (size_t)C * oh * ow;
    double* blk = nullptr;
    CUDA_OK(cudaMallocAsync((void**)&blk, nd * 8 + (i0.size() + i1.size()) * 4, s));
    struct Free { double* p; cudaStream_t s; ~Free() { cudaFreeAsync(p, s); } } guard{blk, s};
    double *dw0 = blk, *dw1 = dw0 + w0.size(), *t0 = dw1 + w1.size(), *t1 = t0 + (size_t)C * oh * W;
    int *di0 = reinterpret_cast<int*>(blk + nd), *di1 = di0 + i0.size();
    CUDA_OK(cudaMemcpyAsync(dw0, w0.data(), w0.size() * 8, cudaMemcpyHostToDevice, s));
    CUDA_OK(cudaMemcpyAsync(dw1, w1.data(), w1.size() * 8, cudaMemcpyHostToDevice, s));
    CUDA_OK(cudaMemcpyAsync(di0, i0.data(), i0.size() * 4, cudaMemcpyHostToDevice, s));
    CUDA_OK(cudaMemcpyAsync(di1, i1.data(), i1.size() * 4, cudaMemcpyHostToDevice, s));
    // equal scales: argsort keeps the order (rows first, then columns), imresize.py:150,163-165
    const long long n0 = (long long)C * oh * W, n1 = (long long)C * oh * ow;
    resize_dim_kernel<float><<<(int)((n0 + 255) / 256 > 2368 ? 2368 : (n0 + 255) / 256), 256, 0, s>>>(img_dev, t0, C, H, W, 0, oh, P0, dw0, di0, round_u8);
    resize_dim_kernel<double><<<(int)((n1 + 255) / 256 > 2368 ? 2368 : (n1 + 255) / 256), 256, 0, s>>>(t0, t1, C, oh, W, 1, ow, P1, dw1, di1, round_u8);
    f64_to_f32_kernel<<<(int)((n1 + 255) / 256), 256, 0, s>>>(t1, out_dev, n1);
    CUDA_OK(cudaStreamSynchronize(s));   // the host weight tables must outlive the copies
    CUDA_OK(cudaGetLastError());
  } catch (const std::exception& ex) { set_last_error(ex.what()); return -1; }
  return 0;
}

int bfsr_imresize_bicubic(const float* img_dev, int32_t C, int32_t H, int32_t W, double scale, float* out_dev, int32_t* out_h,
                          int32_t* out_w, void* stream) {
  return imresize_run(img_dev, C, H, W, scale, out_dev, out_h, out_w, stream, 0);
}

int bfsr_imresize_bicubic_u8(const float* img_dev, int32_t C, int32_t H, int32_t W, double scale, float* out_dev, int32_t* out_h,
                             int32_t* out_w, void* stream) {
  return imresize_run(img_dev, C, H, W, scale, out_dev, out_h, out_w, stream, 1);
}

}  // extern "C"
