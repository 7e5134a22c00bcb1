// fp32 CUDA-core convolution (3x3 / 1x1, stride 1, 'same' zero padding) over NHWC views.
//
// This is the exact-arithmetic member of the conv family: it serves the z-dependent
// coupling convs whose inputs must stay fp32 (the flow inverse amplifies operand
// rounding), every small-K conv that cannot fill a tcgen05 tile
// (K = 27..432), and it is the on-device fp32 reference the tcgen05 kernels are
// checked against at full size.  Replaces cuDNN behind nn.Conv2d at
// RRDBNet_arch.py:25-45, flow.py:26-83, unet.py:10-107.
//
// Tiling: one CTA = 16 x TH output pixels x CO_T output channels, 256 threads; a
// thread owns 8 consecutive pixels of one row x 4 output channels (32 fp32
// accumulators).  Input channels are walked in chunks of 8 through shared memory
// ([ci][row][col] so the 8(+2) pixels a thread needs are 2.5 float4 loads); the 9 taps
// reuse the staged halo tile, so each input element is read from L2 once per CTA.
#include "ops.cuh"
#include <vector>
#include <cstring>

namespace bfsr {

thread_local long long g_launches = 0;

struct ConvArgs {
  View in, out, out2, pre, res1, res2;
  const float* w; const float* bias;
  int cin, cin_pad, cout, cout_pad;
  int H, W;           // output spatial dims
  int in_mode, act;
  float eps, alpha, beta1, beta2;
  int tiles_x;
  int vec_in, vec_out;
};

constexpr int CI = 8;
constexpr int TW = 16;

template <int KS, int CO_T>
__global__ void __launch_bounds__(256) conv_fp32_kernel(ConvArgs a) {
  constexpr int HALO = KS / 2;
  constexpr int NCG = CO_T / 4;
  constexpr int NPG = 256 / NCG;
  constexpr int TH = NPG / 2;
  constexpr int ROWS = TH + 2 * HALO;
  constexpr int COLS = TW + 2 * HALO;
  constexpr int ROWP = 20;                 // padded row pitch (floats), keeps float4 alignment
  constexpr int NV = 8 + KS - 1;

  __shared__ __align__(16) float in_s[CI][ROWS][ROWP];
  __shared__ __align__(16) float w_s[KS * KS][CI][CO_T];

  const int tid = threadIdx.x;
  const int tile = blockIdx.x;
  const int ty0 = (tile / a.tiles_x) * TH, tx0 = (tile % a.tiles_x) * TW;
  const int co_base = blockIdx.y * CO_T;
  const int n = blockIdx.z;
  const int pg = tid / NCG, cg = tid % NCG;
  const int r = pg >> 1, x0 = (pg & 1) * 8;

  const int inH = a.in_mode == IN_UP2 ? a.H >> 1 : a.H;
  const int inW = a.in_mode == IN_UP2 ? a.W >> 1 : a.W;
  const long long in_img = (long long)n * inH * inW;

  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int c0 = 0; c0 < a.cin_pad; c0 += CI) {
    // ---- stage the input halo tile: (ROWS x COLS) pixels x CI channels
    for (int e = tid; e < ROWS * COLS * (CI / 4); e += 256) {
      const int q = e % (CI / 4), pix = e / (CI / 4);
      const int yy = pix / COLS, xx = pix % COLS;
      const int gy = ty0 + yy - HALO, gx = tx0 + xx - HALO;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (gy >= 0 && gy < a.H && gx >= 0 && gx < a.W) {
        const int sy = a.in_mode == IN_UP2 ? gy >> 1 : gy, sx = a.in_mode == IN_UP2 ? gx >> 1 : gx;
        const long long p = in_img + (long long)sy * inW + sx;
        const int cb = c0 + q * 4;
        if (a.vec_in && cb + 4 <= a.cin) {
          v = *reinterpret_cast<const float4*>((const float*)a.in.p + p * a.in.cs + a.in.coff + cb);
        } else {
          if (cb + 0 < a.cin) v.x = ld(a.in, p, cb + 0);
          if (cb + 1 < a.cin) v.y = ld(a.in, p, cb + 1);
          if (cb + 2 < a.cin) v.z = ld(a.in, p, cb + 2);
          if (cb + 3 < a.cin) v.w = ld(a.in, p, cb + 3);
        }
      }
      in_s[q * 4 + 0][yy][xx] = v.x;
      in_s[q * 4 + 1][yy][xx] = v.y;
      in_s[q * 4 + 2][yy][xx] = v.z;
      in_s[q * 4 + 3][yy][xx] = v.w;
    }
    // ---- stage the weight chunk
    for (int e = tid; e < KS * KS * CI * (CO_T / 4); e += 256) {
      const int c4 = e % (CO_T / 4);
      const int ci = (e / (CO_T / 4)) % CI;
      const int tap = e / (CO_T / 4) / CI;
      const float4 v = *reinterpret_cast<const float4*>(
          a.w + ((long long)tap * a.cin_pad + c0 + ci) * a.cout_pad + co_base + c4 * 4);
      *reinterpret_cast<float4*>(&w_s[tap][ci][c4 * 4]) = v;
    }
    __syncthreads();
#pragma unroll
    for (int ci = 0; ci < CI; ++ci) {
#pragma unroll
      for (int dy = 0; dy < KS; ++dy) {
        float v[12];
        const float* row = &in_s[ci][r + dy][x0];
        *reinterpret_cast<float4*>(&v[0]) = *reinterpret_cast<const float4*>(row);
        *reinterpret_cast<float4*>(&v[4]) = *reinterpret_cast<const float4*>(row + 4);
        if (KS == 3) *reinterpret_cast<float2*>(&v[8]) = *reinterpret_cast<const float2*>(row + 8);
#pragma unroll
        for (int dx = 0; dx < KS; ++dx) {
          const float4 w4 = *reinterpret_cast<const float4*>(&w_s[dy * KS + dx][ci][cg * 4]);
#pragma unroll
          for (int px = 0; px < 8; ++px) {
            const float x = v[px + dx];
            acc[px][0] = fmaf(x, w4.x, acc[px][0]);
            acc[px][1] = fmaf(x, w4.y, acc[px][1]);
            acc[px][2] = fmaf(x, w4.z, acc[px][2]);
            acc[px][3] = fmaf(x, w4.w, acc[px][3]);
          }
        }
        (void)NV;
      }
    }
    __syncthreads();
  }

  // ---- epilogue
  const int gy = ty0 + r;
  if (gy >= a.H) return;
  const int co = co_base + cg * 4;
  if (co >= a.cout) return;
  float b[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) b[j] = a.bias[co + j];
#pragma unroll
  for (int px = 0; px < 8; ++px) {
    const int gx = tx0 + x0 + px;
    if (gx >= a.W) break;
    const long long p = ((long long)n * a.H + gy) * a.W + gx;
    float o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float v = acc[px][j] + b[j];
      const int c = co + j;
      if (c < a.cout) {
        if (a.pre.p) v += ld(a.pre, p, c);
        if (a.act == ACT_LRELU) v = v > 0.f ? v : 0.2f * v;
        else if (a.act == ACT_RELU) v = fmaxf(v, 0.f);
        else if (a.act == ACT_CROSS_SIGMOID) { if (c & 1) v = 1.f / (1.f + expf(-(v + 2.f))) + a.eps; }
        v *= a.alpha;
        if (a.res1.p) v = fmaf(a.beta1, ld(a.res1, p, c), v);
        if (a.res2.p) v = fmaf(a.beta2, ld(a.res2, p, c), v);
      }
      o[j] = v;
    }
    if (a.vec_out && co + 4 <= a.cout) {
      *reinterpret_cast<float4*>((float*)a.out.p + p * a.out.cs + a.out.coff + co) = make_float4(o[0], o[1], o[2], o[3]);
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (co + j < a.cout) st(a.out, p, co + j, o[j]);
    }
    if (a.out2.p) {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (co + j < a.cout) st(a.out2, p, co + j, o[j]);
    }
  }
}

template <int KS, int CO_T>
static void launch(const ConvArgs& a, int N, cudaStream_t s) {
  constexpr int TH = (256 / (CO_T / 4)) / 2;
  ConvArgs b = a;
  b.tiles_x = cdiv(a.W, TW);
  dim3 grid(b.tiles_x * cdiv(a.H, TH), a.cout_pad / CO_T, N);
  conv_fp32_kernel<KS, CO_T><<<grid, 256, 0, s>>>(b);
  count_launch();
}

// ---- small-Cin variant (Cin <= 16, Cout == 64): the z-dependent first conv of every coupling (Cin = C/2 = 6, 12), the
// RGB head convs and the stems of the priors.  These are HBM-bound (K <= 144): a warp owns 8 consecutive pixels x all 64
// output channels (2 per lane), so the pre-activation rows it adds and the rows it stores are whole 256-byte runs;
// inputs are broadcast shared-memory reads, weights conflict-free 64-bit reads.
template <int CINP>
__global__ void __launch_bounds__(256) conv3x3_small_kernel(ConvArgs a) {
  constexpr int T = 16;                       // 16 x 16 output pixels per CTA, 4 passes of 8 warps x 8 pixels
  constexpr int PITCH = 20;
  extern __shared__ __align__(16) float smem_small[];
  float* w_s = smem_small;                    // [9][CINP][64]
  float* in_s = smem_small + 9 * CINP * 64;   // [CINP][18][PITCH]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tile = blockIdx.x;
  const int ty0 = (tile / a.tiles_x) * T, tx0 = (tile % a.tiles_x) * T;
  const int n = blockIdx.z;
  const long long img = (long long)n * a.H * a.W;
  for (int e = tid; e < 9 * CINP * 16; e += 256) {
    const int c4 = e & 15, ci = (e >> 4) % CINP, tap = (e >> 4) / CINP;
    const float4 v = ci < a.cin_pad ? *reinterpret_cast<const float4*>(a.w + ((long long)tap * a.cin_pad + ci) * a.cout_pad + c4 * 4)
                                    : make_float4(0.f, 0.f, 0.f, 0.f);
    *reinterpret_cast<float4*>(&w_s[(tap * CINP + ci) * 64 + c4 * 4]) = v;
  }
  for (int e = tid; e < 18 * 18 * CINP; e += 256) {
    const int ci = e % CINP, pix = e / CINP;
    const int yy = pix / 18, xx = pix % 18;
    const int gy = ty0 + yy - 1, gx = tx0 + xx - 1;
    float v = 0.f;
    if (ci < a.cin && gy >= 0 && gy < a.H && gx >= 0 && gx < a.W) v = ld(a.in, img + (long long)gy * a.W + gx, ci);
    in_s[(ci * 18 + yy) * PITCH + xx] = v;
  }
  __syncthreads();
  const float2 b2 = *reinterpret_cast<const float2*>(a.bias + 2 * lane);
  for (int pass = 0; pass < 4; ++pass) {
    const int row = pass * 4 + (warp >> 1), x0 = (warp & 1) * 8;
    float acc[8][2];
#pragma unroll
    for (int i = 0; i < 8; ++i) { acc[i][0] = 0.f; acc[i][1] = 0.f; }
#pragma unroll 2
    for (int ci = 0; ci < CINP; ++ci) {
#pragma unroll
      for (int dy = 0; dy < 3; ++dy) {
        float v[12];
        const float* rp = &in_s[(ci * 18 + row + dy) * PITCH + x0];
        *reinterpret_cast<float4*>(&v[0]) = *reinterpret_cast<const float4*>(rp);
        *reinterpret_cast<float4*>(&v[4]) = *reinterpret_cast<const float4*>(rp + 4);
        *reinterpret_cast<float2*>(&v[8]) = *reinterpret_cast<const float2*>(rp + 8);
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) {
          const float2 w2 = *reinterpret_cast<const float2*>(&w_s[((dy * 3 + dx) * CINP + ci) * 64 + 2 * lane]);
#pragma unroll
          for (int px = 0; px < 8; ++px) { acc[px][0] = fmaf(v[px + dx], w2.x, acc[px][0]); acc[px][1] = fmaf(v[px + dx], w2.y, acc[px][1]); }
        }
      }
    }
    const int gy = ty0 + row;
    if (gy >= a.H) continue;
#pragma unroll
    for (int px = 0; px < 8; ++px) {
      const int gx = tx0 + x0 + px;
      if (gx >= a.W) break;
      const long long p = img + (long long)gy * a.W + gx;
      float o0 = acc[px][0] + b2.x, o1 = acc[px][1] + b2.y;
      if (a.pre.p) { const float2 t = *reinterpret_cast<const float2*>((const float*)a.pre.p + p * a.pre.cs + a.pre.coff + 2 * lane); o0 += t.x; o1 += t.y; }
      if (a.act == ACT_LRELU) { o0 = o0 > 0.f ? o0 : 0.2f * o0; o1 = o1 > 0.f ? o1 : 0.2f * o1; }
      else if (a.act == ACT_RELU) { o0 = fmaxf(o0, 0.f); o1 = fmaxf(o1, 0.f); }
      *reinterpret_cast<float2*>((float*)a.out.p + p * a.out.cs + a.out.coff + 2 * lane) = make_float2(o0 * a.alpha, o1 * a.alpha);
    }
  }
}

static bool small_ok(const ConvW& w, const View& in, const View& out, const ConvEpi& epi, int in_mode) {
  auto v2 = [](const View& v) { return v.fmt == F32 && v.cs % 2 == 0 && v.coff % 2 == 0 && ((uintptr_t)v.p % 8) == 0; };
  return w.ks == 3 && w.cin <= 16 && w.cout == 64 && in_mode == IN_DIRECT && !epi.res1 && !epi.res2 && !epi.out2 &&
         epi.act != ACT_CROSS_SIGMOID && in.fmt == F32 && v2(out) && (!epi.pre || v2(*epi.pre));
}

void conv2d_fp32(const ConvW& w, const View& in, const View& out, const ConvEpi& epi, int in_mode, cudaStream_t s) {
  BFSR_CHECK(in.C == w.cin, "conv: input view has %d channels, weights expect %d", in.C, w.cin);
  BFSR_CHECK(out.C == w.cout, "conv: output view has %d channels, weights produce %d", out.C, w.cout);
  BFSR_CHECK(in.N == out.N, "conv: batch mismatch");
  if (in_mode == IN_UP2) BFSR_CHECK(in.H * 2 == out.H && in.W * 2 == out.W, "conv(up2): spatial mismatch");
  else BFSR_CHECK(in.H == out.H && in.W == out.W, "conv: spatial mismatch %dx%d vs %dx%d", in.H, in.W, out.H, out.W);
  ConvArgs a;
  a.in = in; a.out = out; a.out2 = epi.out2 ? *epi.out2 : View();
  a.pre = epi.pre ? *epi.pre : View();
  a.res1 = epi.res1 ? *epi.res1 : View();
  a.res2 = epi.res2 ? *epi.res2 : View();
  a.w = w.w; a.bias = w.bias;
  a.cin = w.cin; a.cin_pad = w.cin_pad; a.cout = w.cout; a.cout_pad = w.cout_pad;
  a.H = out.H; a.W = out.W; a.in_mode = in_mode; a.act = epi.act;
  a.eps = epi.eps; a.alpha = epi.alpha; a.beta1 = epi.beta1; a.beta2 = epi.beta2;
  a.vec_in = (in.fmt == F32 && in.cs % 4 == 0 && in.coff % 4 == 0 && ((uintptr_t)in.p % 16) == 0);
  a.vec_out = (out.fmt == F32 && out.cs % 4 == 0 && out.coff % 4 == 0 && ((uintptr_t)out.p % 16) == 0);
  a.tiles_x = 0;
  if (out.npix() == 0) return;
  snprintf(g_prof_tag, sizeof g_prof_tag, "fp32 k%d %d->%d %dx%d", w.ks, w.cin, w.cout, out.H, out.W);
  ProfScope prof(PK_CONV_FP32, 2.0 * (double)out.npix() * w.cin * w.ks * w.ks * w.cout, s);
  if (small_ok(w, in, out, epi, in_mode)) {
    a.tiles_x = cdiv(a.W, 16);
    dim3 grid(a.tiles_x * cdiv(a.H, 16), 1, out.N);
    if (w.cin <= 8) {
      const size_t smem = (size_t)(9 * 8 * 64 + 8 * 18 * 20) * 4;
      conv3x3_small_kernel<8><<<grid, 256, smem, s>>>(a);
    } else {
      const size_t smem = (size_t)(9 * 16 * 64 + 16 * 18 * 20) * 4;
      CUDA_OK(cudaFuncSetAttribute(conv3x3_small_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      conv3x3_small_kernel<16><<<grid, 256, smem, s>>>(a);
    }
    count_launch();
    return;
  }
#define L(KS, T) launch<KS, T>(a, out.N, s)
  if (w.ks == 3) { if (w.co_tile == 64) L(3, 64); else if (w.co_tile == 32) L(3, 32); else L(3, 16); }
  else           { if (w.co_tile == 64) L(1, 64); else if (w.co_tile == 32) L(1, 32); else L(1, 16); }
#undef L
}

ConvW pack_conv(const float* w_oihw, int cout, int cin_src, int ks, const float* bias, const float* out_scale,
                const std::vector<int>& cin_map, int tc_min_cin) {
  BFSR_CHECK(ks == 1 || ks == 3, "conv kernel size %d unsupported", ks);
  ConvW c;
  c.ks = ks; c.cout = cout;
  c.cin = cin_map.empty() ? cin_src : (int)cin_map.size();
  c.co_tile = cout > 32 ? 64 : (cout > 16 ? 32 : 16);
  c.cout_pad = cdiv(cout, c.co_tile) * c.co_tile;
  c.cin_pad = cdiv(c.cin, CI) * CI;
  const int taps = ks * ks;
  std::vector<float> h((size_t)taps * c.cin_pad * c.cout_pad, 0.f), hb(c.cout_pad, 0.f);
  for (int co = 0; co < cout; ++co) {
    const float sc = out_scale ? out_scale[co] : 1.f;
    for (int ci = 0; ci < c.cin; ++ci) {
      const int src = cin_map.empty() ? ci : cin_map[ci];
      if (src < 0) continue;
      BFSR_CHECK(src < cin_src, "pack_conv: channel map out of range");
      for (int t = 0; t < taps; ++t)
        h[((size_t)t * c.cin_pad + ci) * c.cout_pad + co] = w_oihw[((size_t)co * cin_src + src) * taps + t] * sc;
    }
    hb[co] = bias ? bias[co] : 0.f;   // final bias: callers fold their own scales into it
  }
  CUDA_OK(cudaMalloc((void**)&c.w, h.size() * 4));
  CUDA_OK(cudaMalloc((void**)&c.bias, hb.size() * 4));
  CUDA_OK(cudaMemcpy(c.w, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
  CUDA_OK(cudaMemcpy(c.bias, hb.data(), hb.size() * 4, cudaMemcpyHostToDevice));
  pack_conv_tc(c, h, tc_min_cin);   // split-bf16 image for the tcgen05 path when the shape is eligible
  return c;
}

void free_conv(ConvW& w) {
  if (w.w) cudaFree(w.w);
  if (w.bias) cudaFree(w.bias);
  if (w.w_tc) cudaFree(w.w_tc);
  if (w.w_tc_fold) cudaFree(w.w_tc_fold);
  if (w.w_tc_f3) cudaFree(w.w_tc_f3);
  w.w = w.bias = nullptr; w.w_tc = nullptr; w.w_tc_fold = nullptr; w.w_tc_f3 = nullptr;
}

}  // namespace bfsr
