// LINF-LP query-side kernels: local Fourier features, the 27-dim conditional flow (both directions, fold + residual
// fused into the inverse), the stride-3 LR embedding conv and a general bilinear resize.
//
// Reference: LINF-LP/models/linf.py:248-407 (~60 ATen kernels per query chunk, 4 grid_sample gathers, an LU solve per
// flow layer per call, flow.py:110-122) — here W^-1 of every NaiveLinear is precomputed in fp64 at load time.
#include "ops.cuh"

namespace bfsr {

// ------------------------------------------------------------------ local Fourier features (linf.py:251-309)
struct FeatArgs {
  View cf;              // (B,h,w,2*hid): [coef | freq]
  const float* coord;   // (B,qh,qw,2) (row, col) in [-1,1]
  const float* cell;    // (B,2)
  const float* phase;   // (hid/2, 2)
  View out;             // (B,qh,qw,4*hid)
  int h, w, qh, qw, hid;
  float shift_y[2], shift_x[2];   // fp32(v*r + 1e-6) for v = -1, +1
  float lo, hi;                   // clamp bounds fp32(-1+1e-6), fp32(1-1e-6)
  float two_ry, v0_ry, two_rx, v0_rx;   // make_coord: c_k = fp32(2r)*k + fp32(-1+r)   (utils.py:105-120)
};

// grid_sample(mode='nearest', align_corners=False) source index: clamp(rne(((c+1)*n-1)/2), 0, n-1)
__device__ __forceinline__ int nearest_index(float c, int n) {
  const float x = __fdiv_rn(__fadd_rn(__fmul_rn(__fadd_rn(c, 1.f), (float)n), -1.f), 2.f);
  int i = (int)nearbyintf(x);
  return i < 0 ? 0 : (i > n - 1 ? n - 1 : i);
}

// One half-block (64 threads) per query; a thread produces channels 2t, 2t+1 of both halves for the four neighbours, so the
// stores into the BF16X2 operand tensor of the MLP are packed 4-byte words (256-byte runs per plane) instead of 2-byte scalars.
__global__ void __launch_bounds__(128) linf_features_kernel(FeatArgs a, long long nq) {
  const long long q = (long long)blockIdx.x * 2 + (threadIdx.x >> 6);
  if (q >= nq) return;
  const int tl = threadIdx.x & 63;
  const int half = a.hid / 2;                      // 128
  const long long t = q / a.qw; const int b = (int)(t / a.qh);
  const float cy = a.coord[q * 2 + 0], cx = a.coord[q * 2 + 1];
  const float cell_y = __fmul_rn(a.cell[b * 2 + 0], (float)a.h), cell_x = __fmul_rn(a.cell[b * 2 + 1], (float)a.w);
  float rel_y[4], rel_x[4], area[4]; long long src[4];
#pragma unroll
  for (int nb = 0; nb < 4; ++nb) {
    const int vx = nb >> 1, vy = nb & 1;           // vx outer (dim 0 = rows), vy inner (linf.py:271-272)
    float sy = __fadd_rn(cy, a.shift_y[vx]), sx = __fadd_rn(cx, a.shift_x[vy]);
    sy = fminf(fmaxf(sy, a.lo), a.hi); sx = fminf(fmaxf(sx, a.lo), a.hi);
    const int iy = nearest_index(sy, a.h), ix = nearest_index(sx, a.w);
    const float qcy = __fadd_rn(__fmul_rn(a.two_ry, (float)iy), a.v0_ry), qcx = __fadd_rn(__fmul_rn(a.two_rx, (float)ix), a.v0_rx);
    rel_y[nb] = __fmul_rn(__fadd_rn(cy, -qcy), (float)a.h);
    rel_x[nb] = __fmul_rn(__fadd_rn(cx, -qcx), (float)a.w);
    area[nb] = __fadd_rn(fabsf(__fmul_rn(rel_y[nb], rel_x[nb])), 1e-9f);
    src[nb] = ((long long)b * a.h + iy) * a.w + ix;
  }
  const float tot = __fadd_rn(__fadd_rn(__fadd_rn(area[0], area[1]), area[2]), area[3]);
  const bool packed = a.out.fmt == BF16X2 && a.cf.fmt == F32 && (a.out.coff & 1) == 0 && (a.out.cs & 1) == 0 && (a.cf.coff & 1) == 0 && (a.cf.cs & 1) == 0;
  for (int k0 = 2 * tl; k0 < half; k0 += 128) {
    float ph[2];
#pragma unroll
    for (int j = 0; j < 2; ++j) ph[j] = fmaf(cell_x, a.phase[(k0 + j) * 2 + 1], __fmul_rn(cell_y, a.phase[(k0 + j) * 2 + 0]));
#pragma unroll
    for (int nb = 0; nb < 4; ++nb) {
      const float wgt = __fdiv_rn(area[3 - nb], tot);          // diagonal swap (linf.py:305-306)
      float oc[2], os[2];
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int k = k0 + j;
        const float fy = ld(a.cf, src[nb], a.hid + k), fx = ld(a.cf, src[nb], a.hid + half + k);
        const float f = __fadd_rn(__fadd_rn(__fmul_rn(fy, rel_y[nb]), __fmul_rn(fx, rel_x[nb])), ph[j]);
        const float ang = __fmul_rn(3.14159265358979323846f, f);
        float sn, cs;
        sincosf(ang, &sn, &cs);
        const float c0 = ld(a.cf, src[nb], k), c1 = ld(a.cf, src[nb], half + k);
        oc[j] = __fmul_rn(__fmul_rn(wgt, c0), cs);
        os[j] = __fmul_rn(__fmul_rn(wgt, c1), sn);
      }
      if (packed) {
        __nv_bfloat16* d = (__nv_bfloat16*)a.out.p + q * a.out.cs + a.out.coff;
#pragma unroll
        for (int hsel = 0; hsel < 2; ++hsel) {
          const float x0 = hsel ? os[0] : oc[0], x1 = hsel ? os[1] : oc[1];
          const __nv_bfloat162 hi = __floats2bfloat162_rn(x0, x1);
          const __nv_bfloat162 lo = __floats2bfloat162_rn(x0 - __low2float(hi), x1 - __high2float(hi));
          const int ch = nb * a.hid + hsel * half + k0;
          *reinterpret_cast<__nv_bfloat162*>(d + ch) = hi;
          *reinterpret_cast<__nv_bfloat162*>(d + a.out.plane + ch) = lo;
        }
      } else {
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          st(a.out, q, nb * a.hid + k0 + j, oc[j]);
          st(a.out, q, nb * a.hid + half + k0 + j, os[j]);
        }
      }
    }
  }
}

void linf_features(const View& cf, const float* coord, const float* cell, const float* phase, const View& out, int qh, int qw,
                   cudaStream_t s) {
  FeatArgs a;
  a.cf = cf; a.coord = coord; a.cell = cell; a.phase = phase; a.out = out;
  a.h = cf.H; a.w = cf.W; a.qh = qh; a.qw = qw; a.hid = cf.C / 2;
  BFSR_CHECK(out.C == 4 * a.hid && out.N == cf.N && out.H == qh && out.W == qw, "linf_features: output shape");
  const double ry = 2.0 / a.h / 2.0, rx = 2.0 / a.w / 2.0;
  for (int v = 0; v < 2; ++v) { a.shift_y[v] = (float)((v ? 1 : -1) * ry + 1e-6); a.shift_x[v] = (float)((v ? 1 : -1) * rx + 1e-6); }
  a.lo = (float)(-1 + 1e-6); a.hi = (float)(1 - 1e-6);
  a.two_ry = (float)(2 * (1.0 / a.h)); a.v0_ry = (float)(-1 + 1.0 / a.h);
  a.two_rx = (float)(2 * (1.0 / a.w)); a.v0_rx = (float)(-1 + 1.0 / a.w);
  const long long nq = (long long)cf.N * qh * qw;
  if (!nq) return;
  snprintf(g_prof_tag, sizeof g_prof_tag, "linf_features q%dx%d", qh, qw);
  ProfScope prof(PK_OTHER, (double)nq * out.C * 4.0, s);
  BFSR_CHECK(a.hid % 4 == 0, "linf_features: hidden width must be a multiple of 4");
  linf_features_kernel<<<(unsigned)((nq + 1) / 2), 128, 0, s>>>(a, nq);
  count_launch();
}

// ------------------------------------------------------------------ conditional flow on D-vectors (flow.py:29-63)
constexpr int FD = 27, FDP = 28;
struct FlowArgs {
  const float* M;      // [n_layers+1][D][D]: forward W_i, inverse W_i^-1 ; index n_layers = `last`
  const float* bias;   // [n_layers+1][D]
  View aff;            // (B,qh,qw, 2*D*n_layers)
  const float* zin;    // NCHW (B,D,qh,qw)
  float* out;          // forward: NCHW (B,D,qh,qw); inverse: NCHW (B,3,OH,OW)
  const float* inp;    // inverse + residual: NCHW (B,3,h,w) LR input (added through bilinear resize), or null
  int n_layers, qh, qw, OH, OW, h, w, ps;
  long long nq;
};

// y[o] (+)= sum_c W[o][c] x[c] in outer-product form: Wt holds the layer's matrix TRANSPOSED ([c][o], rows padded to 28 floats), so
// every 16-byte broadcast load feeds four INDEPENDENT accumulators -- 27 independent FMA chains instead of one 27-long dependent
// chain per output (the kernel is latency-bound).  Per output the products are still added in ascending c: bit-identical.
__device__ __forceinline__ void matvec27(const float* Wt, const float* x, float* y) {
#pragma unroll
  for (int c = 0; c < FD; ++c) {
    const float xc = x[c];
#pragma unroll
    for (int k = 0; k < 7; ++k) {
      const float4 m = *reinterpret_cast<const float4*>(Wt + c * FDP + 4 * k);
      y[4 * k] = fmaf(m.x, xc, y[4 * k]); y[4 * k + 1] = fmaf(m.y, xc, y[4 * k + 1]); y[4 * k + 2] = fmaf(m.z, xc, y[4 * k + 2]);
      if (4 * k + 3 < FD) y[4 * k + 3] = fmaf(m.w, xc, y[4 * k + 3]);
    }
  }
}

// 256 queries per block, two blocks per SM (<= 128 registers): the kernel is bound by instruction latency, so resident warps count
// (ncu, 128-thread blocks at 168 registers: 12 warps per SM, 16 % issue-slot utilisation)
constexpr int FLOW_T = 256;
template <bool INV>
__global__ void __launch_bounds__(FLOW_T, 2) linf_flow_kernel(FlowArgs a) {
  extern __shared__ float sm[];
  // transposed weight rows padded to FDP = 28 floats, read as seven 16-byte broadcast loads (matvec27)
  const int nm = (a.n_layers + 1) * FD * FDP;
  float* Ms = sm; float* bs = sm + nm;
  for (int e = threadIdx.x; e < nm; e += blockDim.x) {           // Ms[layer][c][o] = M[layer][o][c]
    const int r = e / FDP, o = e - r * FDP, l = r / FD, c = r - l * FD;
    Ms[e] = o < FD ? a.M[(l * FD + o) * FD + c] : 0.f;
  }
  for (int e = threadIdx.x; e < (a.n_layers + 1) * FD; e += blockDim.x) bs[e] = a.bias[e];
  // per-layer affine parameters of the block's 128 queries, staged through shared memory: a query's 2*D values of a layer are
  // 216 contiguous bytes of a 2160-byte row, so the block reads them as coalesced runs (a thread walking its own row touched a
  // different cache line per lane and thrashed L1: 2.9 of the 9.6 ms of a config-3 batch went into the two flow passes)
  constexpr int AP = 2 * FD + 1;                                   // padded row: conflict-free per-thread reads
  float* As = bs + (a.n_layers + 1) * FD;                          // [FLOW_T][AP]
  const long long q0 = (long long)blockIdx.x * blockDim.x;
  const int nrow = (int)((a.nq - q0) < (long long)blockDim.x ? (a.nq - q0) : (long long)blockDim.x);
  auto load_layer = [&](int i) {
    __syncthreads();                                               // the previous layer's rows have been consumed
    const float* base = (const float*)a.aff.p + q0 * a.aff.cs + a.aff.coff + i * 2 * FD;
    for (int e = threadIdx.x; e < nrow * 2 * FD; e += blockDim.x) {
      const int r = e / (2 * FD), c = e - r * 2 * FD;
      As[r * AP + c] = __ldg(base + (long long)r * a.aff.cs + c);
    }
    __syncthreads();
  };
  __syncthreads();
  const long long q = q0 + threadIdx.x;
  const bool live = q < a.nq;
  const long long qq = live ? q : a.nq - 1;                        // idle lanes of the last block follow the barriers on a valid row
  const int qx = (int)(qq % a.qw); const long long t = qq / a.qw; const int qy = (int)(t % a.qh); const long long b = t / a.qh;
  const long long plane = (long long)a.qh * a.qw;
  float x[FD], y[FD];
#pragma unroll
  for (int c = 0; c < FD; ++c) x[c] = a.zin[(b * FD + c) * plane + (long long)qy * a.qw + qx];
  const float* af = As + (live ? threadIdx.x : 0) * AP;
  if (!INV) {
    for (int i = 0; i < a.n_layers; ++i) {
      load_layer(i);
      const float* W = Ms + i * FD * FDP;
#pragma unroll
      for (int o = 0; o < FD; ++o) y[o] = bs[i * FD + o];
      matvec27(W, x, y);
#pragma unroll
      for (int c = 0; c < FD; ++c) {
        const float scale = __fdividef(1.f, 1.f + __expf(-(af[c] + 2.f))) + 1e-4f;
        x[c] = fmaf(y[c], scale, af[FD + c]);
      }
    }
    const float* W = Ms + a.n_layers * FD * FDP;
#pragma unroll
    for (int o = 0; o < FD; ++o) y[o] = bs[a.n_layers * FD + o];
    matvec27(W, x, y);
    if (live) {
#pragma unroll
      for (int o = 0; o < FD; ++o) a.out[(b * FD + o) * plane + (long long)qy * a.qw + qx] = y[o];
    }
  } else {
    {   // last^-1
      const float* W = Ms + a.n_layers * FD * FDP;
#pragma unroll
      for (int c = 0; c < FD; ++c) x[c] -= bs[a.n_layers * FD + c];
#pragma unroll
      for (int o = 0; o < FD; ++o) y[o] = 0.f;
      matvec27(W, x, y);
    }
    for (int i = a.n_layers - 1; i >= 0; --i) {
      load_layer(i);
      const float* W = Ms + i * FD * FDP;
#pragma unroll
      for (int c = 0; c < FD; ++c) {
        const float scale = __fdividef(1.f, 1.f + __expf(-(af[c] + 2.f))) + 1e-4f;
        x[c] = __fdividef(y[c] - af[FD + c], scale) - bs[i * FD + c];
      }
#pragma unroll
      for (int o = 0; o < FD; ++o) y[o] = 0.f;
      matvec27(W, x, y);
    }
    // fold 3x3 patches (== pixel_shuffle(3), linf.py:401-406), crop to (OH,OW), add bilinear(inp) (test.py:168-171)
    if (!live) return;
    if (a.ps == 0) {     // stand-alone Flow.inverse: the D-vector itself, same NCHW layout as the forward output
#pragma unroll
      for (int o = 0; o < FD; ++o) a.out[(b * FD + o) * plane + (long long)qy * a.qw + qx] = y[o];
      return;
    }
    const int ps = a.ps;
    for (int c = 0; c < 3; ++c)
      for (int ky = 0; ky < ps; ++ky)
        for (int kx = 0; kx < ps; ++kx) {
          const int oy = qy * ps + ky, ox = qx * ps + kx;
          if (oy >= a.OH || ox >= a.OW) continue;
          float v = y[c * ps * ps + ky * ps + kx];
          if (a.inp) {
            const float sh = (float)a.h / (float)a.OH, sw = (float)a.w / (float)a.OW;
            float fy = sh * (oy + 0.5f) - 0.5f, fx = sw * (ox + 0.5f) - 0.5f;
            fy = fy < 0.f ? 0.f : fy; fx = fx < 0.f ? 0.f : fx;
            const int y0 = (int)fy, x0 = (int)fx;
            const int y1 = y0 + (y0 < a.h - 1 ? 1 : 0), x1 = x0 + (x0 < a.w - 1 ? 1 : 0);
            const float ly = fy - y0, lx = fx - x0;
            const float* sp = a.inp + (b * 3 + c) * (long long)a.h * a.w;
            v += (1.f - ly) * ((1.f - lx) * sp[y0 * a.w + x0] + lx * sp[y0 * a.w + x1]) +
                 ly * ((1.f - lx) * sp[y1 * a.w + x0] + lx * sp[y1 * a.w + x1]);
          }
          a.out[((b * 3 + c) * a.OH + oy) * (long long)a.OW + ox] = v;
        }
  }
}

void linf_flow(bool inverse, const float* M, const float* bias, int n_layers, const View& aff, const float* zin, int B,
               int qh, int qw, float* out, int OH, int OW, const float* inp, int h, int w, int ps, cudaStream_t s) {
  BFSR_CHECK(ps == 0 || 3 * ps * ps == FD, "linf_flow: only patch_size 3 (D = 27) is built");
  BFSR_CHECK(aff.fmt == F32 && aff.C == 2 * FD * n_layers, "linf_flow: affine_info shape");
  FlowArgs a;
  a.M = M; a.bias = bias; a.aff = aff; a.zin = zin; a.out = out; a.inp = inp;
  a.n_layers = n_layers; a.qh = qh; a.qw = qw; a.OH = OH; a.OW = OW; a.h = h; a.w = w; a.ps = ps;
  a.nq = (long long)B * qh * qw;
  if (!a.nq) return;
  const size_t smem = ((size_t)(n_layers + 1) * (FD * FDP + FD) + FLOW_T * (2 * FD + 1)) * 4;
  BFSR_CHECK(smem <= 227 * 1024, "linf_flow: %d flow layers need %zu bytes of shared memory (limit 227 KB)", n_layers, smem);
  if (smem > 48 * 1024) {   // flow_layers >= 16: opt in to the large carve-out (the shipped models have 10 layers = 33 KB)
    CUDA_OK(cudaFuncSetAttribute(linf_flow_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CUDA_OK(cudaFuncSetAttribute(linf_flow_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  const int grid = cdiv(a.nq, FLOW_T);
  snprintf(g_prof_tag, sizeof g_prof_tag, "linf_flow-%s q%dx%d", inverse ? "inv" : "fwd", qh, qw);
  ProfScope prof(PK_OTHER, (double)a.nq * (2.0 * FD * n_layers + 2 * FD) * 4.0, s);
  if (inverse) linf_flow_kernel<true><<<grid, FLOW_T, smem, s>>>(a);
  else linf_flow_kernel<false><<<grid, FLOW_T, smem, s>>>(a);
  count_launch();
}

// ------------------------------------------------------------------ stride-3 3x3 conv, pad 1 (unet.py lr_proj.0) + LeakyReLU
__global__ void conv3x3_s3_kernel(const float* x, int B, int Cin, int h, int w, const float* wgt, const float* bias, View out,
                                  long long n) {
  long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  const int co = (int)(e % out.C); const long long pix = e / out.C;
  const int ox = (int)(pix % out.W); const long long t = pix / out.W; const int oy = (int)(t % out.H); const long long b = t / out.H;
  float acc = bias[co];
  for (int ci = 0; ci < Cin; ++ci)
    for (int ky = 0; ky < 3; ++ky)
      for (int kx = 0; kx < 3; ++kx) {
        const int iy = oy * 3 - 1 + ky, ix = ox * 3 - 1 + kx;
        if (iy < 0 || iy >= h || ix < 0 || ix >= w) continue;
        acc = fmaf(x[((b * Cin + ci) * h + iy) * (long long)w + ix], wgt[((co * Cin + ci) * 3 + ky) * 3 + kx], acc);
      }
  st(out, pix, co, acc > 0.f ? acc : 0.2f * acc);
}
void conv3x3_s3_lrelu(const float* x_nchw, int B, int Cin, int h, int w, const float* w_oihw_dev, const float* bias_dev,
                      const View& out, cudaStream_t s) {
  BFSR_CHECK(out.H == (h + 2 - 3) / 3 + 1 && out.W == (w + 2 - 3) / 3 + 1 && out.N == B, "conv3x3_s3: output shape");
  const long long n = out.npix() * out.C;
  if (!n) return;
  conv3x3_s3_kernel<<<cdiv(n, 128), 128, 0, s>>>(x_nchw, B, Cin, h, w, w_oihw_dev, bias_dev, out, n);
  count_launch();
}

// ------------------------------------------------------------------ F.interpolate(size=..., bilinear, align_corners=False) NHWC
__global__ void bilinear_resize_kernel(View src, View dst, long long n) {
  long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  const int C = dst.C;
  const long long pix = e / C; const int c = (int)(e % C);
  const int x = (int)(pix % dst.W); const long long t = pix / dst.W; const int y = (int)(t % dst.H); const long long b = t / dst.H;
  const float sh = (float)src.H / (float)dst.H, sw = (float)src.W / (float)dst.W;
  float fy = sh * (y + 0.5f) - 0.5f, fx = sw * (x + 0.5f) - 0.5f;
  fy = fy < 0.f ? 0.f : fy; fx = fx < 0.f ? 0.f : fx;
  const int y0 = (int)fy, x0 = (int)fx;
  const int y1 = y0 + (y0 < src.H - 1 ? 1 : 0), x1 = x0 + (x0 < src.W - 1 ? 1 : 0);
  const float ly = fy - y0, lx = fx - x0;
  const long long sb = b * src.H * src.W;
  const float a00 = ld(src, sb + (long long)y0 * src.W + x0, c), a01 = ld(src, sb + (long long)y0 * src.W + x1, c);
  const float a10 = ld(src, sb + (long long)y1 * src.W + x0, c), a11 = ld(src, sb + (long long)y1 * src.W + x1, c);
  st(dst, pix, c, (1.f - ly) * ((1.f - lx) * a00 + lx * a01) + ly * ((1.f - lx) * a10 + lx * a11));
}
void bilinear_resize(const View& src, const View& dst, cudaStream_t s) {
  BFSR_CHECK(src.C == dst.C && src.N == dst.N, "bilinear_resize: channel/batch mismatch");
  const long long n = dst.npix() * dst.C;
  if (!n) return;
  bilinear_resize_kernel<<<cdiv(n, 256), 256, 0, s>>>(src, dst, n);
  count_launch();
}


// ------------------------------------------------------------------ test-time input construction (datasets/wrappers.py:154-238, 516-613)
// F.interpolate(x, size, mode='bilinear', align_corners=False) on NCHW planes (area_pixel_compute_source_index semantics)
__global__ void bilinear_nchw_kernel(const float* src, float* dst, long long n, int hi, int wi, int ho, int wo, float sub, float mul) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  const int x = (int)(e % wo); const long long t = e / wo; const int y = (int)(t % ho); const long long pl = t / ho;
  const float sh = (float)hi / (float)ho, sw = (float)wi / (float)wo;
  float fy = sh * (y + 0.5f) - 0.5f, fx = sw * (x + 0.5f) - 0.5f;
  fy = fy < 0.f ? 0.f : fy; fx = fx < 0.f ? 0.f : fx;
  const int y0 = (int)fy, x0 = (int)fx;
  const int y1 = y0 + (y0 < hi - 1 ? 1 : 0), x1 = x0 + (x0 < wi - 1 ? 1 : 0);
  const float ly = fy - y0, lx = fx - x0;
  const float* sp = src + pl * (long long)hi * wi;
  const float a00 = (sp[y0 * wi + x0] - sub) * mul, a01 = (sp[y0 * wi + x1] - sub) * mul;
  const float a10 = (sp[y1 * wi + x0] - sub) * mul, a11 = (sp[y1 * wi + x1] - sub) * mul;
  dst[e] = (1.f - ly) * ((1.f - lx) * a00 + lx * a01) + ly * ((1.f - lx) * a10 + lx * a11);
}
static void bilinear_nchw(const float* src, float* dst, long long planes, int hi, int wi, int ho, int wo, float sub, float mul, cudaStream_t s) {
  const long long n = planes * ho * wo;
  if (!n) return;
  bilinear_nchw_kernel<<<cdiv(n, 256), 256, 0, s>>>(src, dst, n, hi, wi, ho, wo, sub, mul);
  count_launch();
}
// inp = (lr - 0.5) / 0.5 ; gt_lr_up = unfold_ps(lr_up - up(down(lr_up))) zero padded ; coord = patch-centre coords (0 in the pad) ; cell
__global__ void linf_inputs_kernel(const float* lr01, const float* lr_up, const float* lr_udu, int B, int h, int w, int H, int W,
                                   int ps, int qh, int qw, float* inp, float* coord, float* cell, float* gt, long long n_gt,
                                   long long n_inp, float ay, float by, float ax, float bx, float cy, float cx) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e < n_gt) {   // gt[b][c*ps*ps + ky*ps + kx][a][q]
    const int q = (int)(e % qw); long long t = e / qw; const int a = (int)(t % qh); t /= qh;
    const int D = 3 * ps * ps; const int d = (int)(t % D); const int b = (int)(t / D);
    const int c = d / (ps * ps), ky = (d / ps) % ps, kx = d % ps;
    const int Y = a * ps + ky, X = q * ps + kx;
    float v = 0.f;
    if (Y < H && X < W) { const long long i = (((long long)b * 3 + c) * H + Y) * W + X; v = lr_up[i] - lr_udu[i]; }
    gt[e] = v;
  }
  if (e < n_inp) inp[e] = (lr01[e] - 0.5f) / 0.5f;
  if (e < (long long)B * qh * qw) {
    const int q = (int)(e % qw); const int a = (int)((e / qw) % qh);
    const int Y = a * ps + ps / 2, X = q * ps + ps / 2;
    float vy = 0.f, vx = 0.f;
    if (Y < H && X < W) { vy = __fadd_rn(ay, __fmul_rn(by, (float)Y)); vx = __fadd_rn(ax, __fmul_rn(bx, (float)X)); }   // make_coord, utils.py:105-120
    coord[2 * e] = vy; coord[2 * e + 1] = vx;
  }
  if (e < B) { cell[2 * e] = cy; cell[2 * e + 1] = cx; }
}
void linf_build_inputs(const float* lr01, int B, int h, int w, int H, int W, int ps, int qh, int qw, float* scratch, float* inp,
                       float* coord, float* cell, float* gt, cudaStream_t s) {
  if (B == 0) return;
  const long long nHR = (long long)B * 3 * H * W, nLR = (long long)B * 3 * h * w;
  float* lr_up = scratch; float* down = scratch + nHR; float* udu = down + nLR;
  bilinear_nchw(lr01, lr_up, (long long)B * 3, h, w, H, W, 0.5f, 2.0f, s);     // lr_up = bilinear((lr-0.5)/0.5)  ((x-0.5)*2 == (x-0.5)/0.5 exactly)
  bilinear_nchw(lr_up, down, (long long)B * 3, H, W, h, w, 0.f, 1.f, s);
  bilinear_nchw(down, udu, (long long)B * 3, h, w, H, W, 0.f, 1.f, s);
  const long long n_gt = (long long)B * 3 * ps * ps * qh * qw;
  const long long n = n_gt > nLR ? n_gt : nLR;
  const double ry = 1.0 / H, rx = 1.0 / W;
  linf_inputs_kernel<<<cdiv(n, 256), 256, 0, s>>>(lr01, lr_up, udu, B, h, w, H, W, ps, qh, qw, inp, coord, cell, gt, n_gt, nLR,
                                                (float)(-1.0 + ry), (float)(2.0 * ry), (float)(-1.0 + rx), (float)(2.0 * rx),
                                                (float)(2.0 / H), (float)(2.0 / W));
  count_launch();
}

}  // namespace bfsr
