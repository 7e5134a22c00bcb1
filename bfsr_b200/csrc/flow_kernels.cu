// HBM-bound kernels of the conditional flow: the fused FlowStep (both directions),
// Split2d, latent normalisation, Squeeze2d index maps, layout and resampling.
//
// One launch per FlowStep: ActNorm, InvertibleConv1x1 and both affine couplings are
// fused so the flow state z makes exactly one HBM round trip per step
// (reference: ~15 ATen kernels per step, FlowStep.py:88-129; plus an fp64
// torch.inverse per step per call, Permutations.py:41 — here W^-1 is folded with the
// ActNorm scale once at load time).  Squeeze2d / Unsqueeze2d (flow.py:122-152) never
// run as copies: they are index maps in the load (encode) or store (decode) of the
// adjacent step.
//
// Per-CTA scheme: a tile of PIX pixels x C channels is staged in shared memory with
// coalesced NHWC reads (the elementwise coupling math is applied on the way in), the
// C x C channel mix runs from shared memory with 4 output channels per thread, and the
// result is written with coalesced stores.
#include "ops.cuh"

namespace bfsr {

// ------------------------------------------------------------------ fused flow steps
struct StepArgs {
  View zin, zout, h, hF;
  View z1op;           // optional: bf16 (hi, lo) operand copy of the first C/2 output channels, padded to a multiple of 8
  const float* M;      // [C][C] row-major (out, in)
  const float* MT;     // [C][C] transposed (in, out), or null
  const float* cvec;   // [C]
  int C, H, W;         // dims of the squeezed (level) tensor
  long long npix;
  int squeeze_in, unsqueeze_out, has_h, has_hF;
  int pix_per_block;
};

// element (pix, c) of the level tensor read from the un-squeezed source: c = 4c' + 2fh + fw
__device__ __forceinline__ float ld_squeezed(const View& v, int H, int W, long long pix, int c) {
  const int j = (int)(pix % W); const long long t = pix / W; const int i = (int)(t % H); const long long b = t / H;
  const int cc = c >> 2, fh = (c >> 1) & 1, fw = c & 1;
  const long long sp = (b * (2 * H) + (2 * i + fh)) * (2 * W) + (2 * j + fw);
  return ld(v, sp, cc);
}

template <bool INV>
__global__ void __launch_bounds__(256) flowstep_kernel(StepArgs a) {
  extern __shared__ __align__(16) float sm[];
  const int C = a.C, C4 = C >> 2, ZP = C + 1;
  float* Mt = sm;                      // [C][C] transposed: Mt[ci*C + co]
  float* zs = sm + C * C;              // [PIX][C+1]
  const int tid = threadIdx.x, P = a.pix_per_block;
  const long long p0 = (long long)blockIdx.x * P;

  if (a.MT) { for (int e = tid; e < C * C; e += 256) Mt[e] = a.MT[e]; }
  else { for (int e = tid; e < C * C; e += 256) { const int co = e / C, ci = e % C; Mt[ci * C + co] = a.M[e]; } }

  // ---- phase 1: gather z, apply the elementwise part that precedes the channel mix
  for (int e = tid; e < P * C; e += 256) {
    const int p = e / C, c = e % C;
    const long long pix = p0 + p;
    float z = 0.f;
    if (pix < a.npix) {
      z = a.squeeze_in ? ld_squeezed(a.zin, a.H, a.W, pix, c) : ld(a.zin, pix, c);
      if (a.has_h && c >= C / 2) {
        const int j = c - C / 2;
        const float shift = ld(a.h, pix, 2 * j), scale = ld(a.h, pix, 2 * j + 1);
        z = INV ? z / scale - shift : (z + shift) * scale;
      }
      if (INV && a.has_hF) {
        const float shiftF = ld(a.hF, pix, 2 * c), scaleF = ld(a.hF, pix, 2 * c + 1);
        z = z / scaleF - shiftF;
      }
    }
    zs[p * ZP + c] = z;
  }
  __syncthreads();

  // ---- phase 2: channel mix, register tile of 4 pixels x 4 output channels per thread (5 LDS per 16 FMA)
  const int PQ = P >> 2;
  for (int e = tid; e < PQ * C4; e += 256) {
    const int pq = e / C4, g = e % C4;
    float4 acc[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    const float* zr = zs + (pq * 4) * ZP;
    const float* mr = Mt + 4 * g;
#pragma unroll 4
    for (int ci = 0; ci < C; ++ci) {
      const float4 m = *reinterpret_cast<const float4*>(mr + ci * C);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float zv = zr[k * ZP + ci];
        acc[k].x = fmaf(m.x, zv, acc[k].x); acc[k].y = fmaf(m.y, zv, acc[k].y);
        acc[k].z = fmaf(m.z, zv, acc[k].z); acc[k].w = fmaf(m.w, zv, acc[k].w);
      }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const long long pix = p0 + pq * 4 + k;
      if (pix >= a.npix) continue;
      float o[4] = {acc[k].x, acc[k].y, acc[k].z, acc[k].w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int c = 4 * g + j;
        if (INV) o[j] -= a.cvec[c];
        else {
          o[j] += a.cvec[c];
          if (a.has_hF) { const float shiftF = ld(a.hF, pix, 2 * c), scaleF = ld(a.hF, pix, 2 * c + 1); o[j] = (o[j] + shiftF) * scaleF; }
        }
      }
      if (a.unsqueeze_out) {
        const int jx = (int)(pix % a.W); const long long t = pix / a.W; const int iy = (int)(t % a.H); const long long b = t / a.H;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int fh = j >> 1, fw = j & 1;
          const long long dp = (b * (2 * a.H) + (2 * iy + fh)) * (2 * a.W) + (2 * jx + fw);
          st(a.zout, dp, g, o[j]);
        }
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) st(a.zout, pix, 4 * g + j, o[j]);
      }
    }
  }
}

// ---- small-C variant (C = 12, 24): one pixel per thread, the whole step in registers.  Per pixel the thread streams
// z (C), h (C) and hF (2C) with 128-bit loads (consecutive lanes read consecutive pixels, so the warp's loads tile a
// contiguous span), the C x C mix reads the matrix from shared memory as broadcast float4s, and the result leaves with
// 128-bit stores.  Squeeze2d / Unsqueeze2d are index maps on the load / store side.
template <int C, bool INV>
__global__ void __launch_bounds__(128) flowstep_px_kernel(StepArgs a) {
  __shared__ __align__(16) float Ms[C * C];
  __shared__ float cs[C];
  for (int e = threadIdx.x; e < C * C; e += blockDim.x) Ms[e] = a.M[e];
  for (int e = threadIdx.x; e < C; e += blockDim.x) cs[e] = a.cvec[e];
  __syncthreads();
  const long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= a.npix) return;
  float z[C];
  const int jx = (int)(pix % a.W); const long long t = pix / a.W; const int iy = (int)(t % a.H); const long long b = t / a.H;
  if (a.squeeze_in) {
#pragma unroll
    for (int f = 0; f < 4; ++f) {
      const long long sp = (b * (2 * a.H) + (2 * iy + (f >> 1))) * (2 * a.W) + (2 * jx + (f & 1));
      const float* src = (const float*)a.zin.p + sp * a.zin.cs + a.zin.coff;
#pragma unroll
      for (int cc = 0; cc < C / 4; ++cc) z[4 * cc + f] = src[cc];
    }
  } else {
    const float4* src = reinterpret_cast<const float4*>((const float*)a.zin.p + pix * a.zin.cs + a.zin.coff);
#pragma unroll
    for (int k = 0; k < C / 4; ++k) { const float4 v = src[k]; z[4 * k] = v.x; z[4 * k + 1] = v.y; z[4 * k + 2] = v.z; z[4 * k + 3] = v.w; }
  }
  if (a.has_h) {   // self-conditional coupling on the second half: (shift, scale) pairs
    const float4* hp = reinterpret_cast<const float4*>((const float*)a.h.p + pix * a.h.cs + a.h.coff);
#pragma unroll
    for (int k = 0; k < C / 4; ++k) {
      const float4 v = hp[k];           // pairs for channels C/2 + 2k, C/2 + 2k + 1
      const int c0 = C / 2 + 2 * k;
      if (INV) { z[c0] = z[c0] / v.y - v.x; z[c0 + 1] = z[c0 + 1] / v.w - v.z; }
      else { z[c0] = (z[c0] + v.x) * v.y; z[c0 + 1] = (z[c0 + 1] + v.z) * v.w; }
    }
  }
  const float4* fp = reinterpret_cast<const float4*>((const float*)a.hF.p + pix * a.hF.cs + a.hF.coff);
  if (INV && a.has_hF) {
#pragma unroll
    for (int k = 0; k < C / 2; ++k) { const float4 v = fp[k]; z[2 * k] = z[2 * k] / v.y - v.x; z[2 * k + 1] = z[2 * k + 1] / v.w - v.z; }
  }
  float o[C];
#pragma unroll
  for (int co = 0; co < C; ++co) {
    float acc = 0.f;
#pragma unroll
    for (int k = 0; k < C / 4; ++k) {
      const float4 m = *reinterpret_cast<const float4*>(&Ms[co * C + 4 * k]);
      acc = fmaf(m.x, z[4 * k], acc); acc = fmaf(m.y, z[4 * k + 1], acc);
      acc = fmaf(m.z, z[4 * k + 2], acc); acc = fmaf(m.w, z[4 * k + 3], acc);
    }
    o[co] = INV ? acc - cs[co] : acc + cs[co];
  }
  if (!INV && a.has_hF) {
#pragma unroll
    for (int k = 0; k < C / 2; ++k) { const float4 v = fp[k]; o[2 * k] = (o[2 * k] + v.x) * v.y; o[2 * k + 1] = (o[2 * k + 1] + v.z) * v.w; }
  }
  if (a.z1op.p) {   // operand copy of z1 for the next coupling conv: C/2 channels zero-padded to a multiple of 8, bf16 (hi, lo)
    constexpr int CP = (C / 2 + 7) & ~7;
    __nv_bfloat16* d = (__nv_bfloat16*)a.z1op.p + pix * a.z1op.cs + a.z1op.coff;
#pragma unroll
    for (int k = 0; k < CP / 8; ++k) {
      uint32_t hi[4], lo[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int c0 = 8 * k + 2 * e;
        const float x0 = c0 < C / 2 ? o[c0] : 0.f, x1 = c0 + 1 < C / 2 ? o[c0 + 1] : 0.f;
        const __nv_bfloat162 hh = __floats2bfloat162_rn(x0, x1);
        const __nv_bfloat162 ll = __floats2bfloat162_rn(x0 - __low2float(hh), x1 - __high2float(hh));
        hi[e] = *reinterpret_cast<const uint32_t*>(&hh); lo[e] = *reinterpret_cast<const uint32_t*>(&ll);
      }
      *reinterpret_cast<uint4*>(d + 8 * k) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
      *reinterpret_cast<uint4*>(d + a.z1op.plane + 8 * k) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    }
  }
  if (a.unsqueeze_out) {
#pragma unroll
    for (int f = 0; f < 4; ++f) {
      const long long dp = (b * (2 * a.H) + (2 * iy + (f >> 1))) * (2 * a.W) + (2 * jx + (f & 1));
      float* dst = (float*)a.zout.p + dp * a.zout.cs + a.zout.coff;
#pragma unroll
      for (int cc = 0; cc < C / 4; ++cc) dst[cc] = o[4 * cc + f];
    }
  } else {
    float4* dst = reinterpret_cast<float4*>((float*)a.zout.p + pix * a.zout.cs + a.zout.coff);
#pragma unroll
    for (int k = 0; k < C / 4; ++k) dst[k] = make_float4(o[4 * k], o[4 * k + 1], o[4 * k + 2], o[4 * k + 3]);
  }
}


// ---- wide-C variant (C = 96, level 3 of the shipped topology): tile of 64 pixels per CTA staged in shared memory with
// 128-bit loads (coupling / ft-affine inverse applied on the way in), C x C mix with a 2-pixel x 12-channel register tile
// per thread (5 shared loads per 24 FMA), 128-bit epilogue.  No squeeze folding (the two boundary steps of a level use the
// generic kernel).
template <int C, bool INV>
__global__ void __launch_bounds__(256) flowstep_wide_kernel(StepArgs a) {
  constexpr int P = 64, ZP = C + 4, C4 = C / 4, NG = C / 12;
  static_assert(C % 12 == 0 && C % 8 == 0 && (P / 2) * NG == 256, "tile shape");
  extern __shared__ __align__(16) float sm[];
  float* Mt = sm;                      // [ci][co]
  float* zs = sm + C * C;              // [P][ZP]
  const int tid = threadIdx.x;
  if (a.MT) {     // pre-transposed copy: straight 128-bit copy (a transposing fill is a 32-way bank conflict per store)
    for (int e = tid; e < C * C / 4; e += 256) reinterpret_cast<float4*>(Mt)[e] = __ldg(reinterpret_cast<const float4*>(a.MT) + e);
  } else {
    for (int e = tid; e < C * C; e += 256) { const int co = e / C, ci = e % C; Mt[ci * C + co] = a.M[e]; }
  }
  // persistent over pixel tiles: the 36 KB mix matrix is staged once per CTA, not once per 64 pixels
  const long long ntiles = (a.npix + P - 1) / P;
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
  const long long p0 = tile * P;
  __syncthreads();                       // the previous tile's phase 2 has finished reading zs
  // ---- phase 1
  for (int e = tid; e < P * C4; e += 256) {
    const int p = e / C4, k = e % C4, c = 4 * k;
    const long long pix = p0 + p;
    float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    if (pix < a.npix) {
      z = *reinterpret_cast<const float4*>((const float*)a.zin.p + pix * a.zin.cs + a.zin.coff + c);
      if (a.has_h && c >= C / 2) {
        const float4* hp = reinterpret_cast<const float4*>((const float*)a.h.p + pix * a.h.cs + a.h.coff + 2 * (c - C / 2));
        const float4 h0 = hp[0], h1 = hp[1];          // (shift, scale) pairs of channels c..c+3
        if (INV) { z.x = z.x / h0.y - h0.x; z.y = z.y / h0.w - h0.z; z.z = z.z / h1.y - h1.x; z.w = z.w / h1.w - h1.z; }
        else { z.x = (z.x + h0.x) * h0.y; z.y = (z.y + h0.z) * h0.w; z.z = (z.z + h1.x) * h1.y; z.w = (z.w + h1.z) * h1.w; }
      }
      if (INV && a.has_hF) {
        const float4* fp = reinterpret_cast<const float4*>((const float*)a.hF.p + pix * a.hF.cs + a.hF.coff + 2 * c);
        const float4 f0 = fp[0], f1 = fp[1];
        z.x = z.x / f0.y - f0.x; z.y = z.y / f0.w - f0.z; z.z = z.z / f1.y - f1.x; z.w = z.w / f1.w - f1.z;
      }
    }
    *reinterpret_cast<float4*>(zs + p * ZP + c) = z;
  }
  __syncthreads();
  // ---- phase 2: thread = (pixel pair, group of 12 output channels)
  const int g = tid % NG, pq = tid / NG;
  float acc[2][12];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 12; ++j) acc[i][j] = 0.f;
  const float* z0 = zs + (2 * pq) * ZP;
  const float* mr = Mt + 12 * g;
#pragma unroll 4
  for (int ci = 0; ci < C; ++ci) {
    const float4 m0 = *reinterpret_cast<const float4*>(mr + ci * C), m1 = *reinterpret_cast<const float4*>(mr + ci * C + 4),
                 m2 = *reinterpret_cast<const float4*>(mr + ci * C + 8);
    const float m[12] = {m0.x, m0.y, m0.z, m0.w, m1.x, m1.y, m1.z, m1.w, m2.x, m2.y, m2.z, m2.w};
    const float za = z0[ci], zb = z0[ZP + ci];
#pragma unroll
    for (int j = 0; j < 12; ++j) { acc[0][j] = fmaf(m[j], za, acc[0][j]); acc[1][j] = fmaf(m[j], zb, acc[1][j]); }
  }
  float cv[12];
#pragma unroll
  for (int j = 0; j < 12; ++j) cv[j] = a.cvec[12 * g + j];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const long long pix = p0 + 2 * pq + i;
    if (pix >= a.npix) continue;
    float o[12];
#pragma unroll
    for (int j = 0; j < 12; ++j) o[j] = INV ? acc[i][j] - cv[j] : acc[i][j] + cv[j];
    if (!INV && a.has_hF) {
      const float4* fp = reinterpret_cast<const float4*>((const float*)a.hF.p + pix * a.hF.cs + a.hF.coff + 24 * g);
#pragma unroll
      for (int k = 0; k < 6; ++k) { const float4 v = fp[k]; o[2 * k] = (o[2 * k] + v.x) * v.y; o[2 * k + 1] = (o[2 * k + 1] + v.z) * v.w; }
    }
    float4* dst = reinterpret_cast<float4*>((float*)a.zout.p + pix * a.zout.cs + a.zout.coff + 12 * g);
#pragma unroll
    for (int k = 0; k < 3; ++k) dst[k] = make_float4(o[4 * k], o[4 * k + 1], o[4 * k + 2], o[4 * k + 3]);
    if (a.z1op.p && 12 * g < C / 2) {     // C/2 is a multiple of 12 here: whole groups fall into the first half
      __nv_bfloat16* d = (__nv_bfloat16*)a.z1op.p + pix * a.z1op.cs + a.z1op.coff + 12 * g;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        uint32_t hi[2], lo[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const float x0 = o[4 * k + 2 * e], x1 = o[4 * k + 2 * e + 1];
          const __nv_bfloat162 hh = __floats2bfloat162_rn(x0, x1);
          const __nv_bfloat162 ll = __floats2bfloat162_rn(x0 - __low2float(hh), x1 - __high2float(hh));
          hi[e] = *reinterpret_cast<const uint32_t*>(&hh); lo[e] = *reinterpret_cast<const uint32_t*>(&ll);
        }
        *reinterpret_cast<uint2*>(d + 4 * k) = make_uint2(hi[0], hi[1]);
        *reinterpret_cast<uint2*>(d + a.z1op.plane + 4 * k) = make_uint2(lo[0], lo[1]);
      }
    }
  }
  }   // tile loop
}

static bool vec4_ok(const View& v) {
  return v.fmt == F32 && v.cs % 4 == 0 && v.coff % 4 == 0 && ((uintptr_t)v.p % 16) == 0;
}

static void launch_step(bool inv, const StepW& w, const View& zin, bool sq_in, const View* h, const View* hF,
                        const View& zout, bool unsq_out, cudaStream_t s, const View* z1op) {
  StepArgs a;
  a.C = w.C;
  BFSR_CHECK(w.C % 4 == 0, "flow step: C=%d not divisible by 4", w.C);
  if (sq_in) {
    BFSR_CHECK(zin.C * 4 == w.C && zout.C == w.C && zin.H == 2 * zout.H && zin.W == 2 * zout.W, "flowstep: squeeze shapes");
    a.H = zout.H; a.W = zout.W; a.npix = zout.npix();
  } else {
    BFSR_CHECK(zin.C == w.C, "flowstep: z has %d channels, step expects %d", zin.C, w.C);
    a.H = zin.H; a.W = zin.W; a.npix = zin.npix();
    if (unsq_out) BFSR_CHECK(zout.C * 4 == w.C && zout.H == 2 * zin.H && zout.W == 2 * zin.W, "flowstep: unsqueeze shapes");
    else BFSR_CHECK(zout.C == w.C && zout.npix() == zin.npix(), "flowstep: output shape");
  }
  if (h) BFSR_CHECK(h->C == (w.C - w.C / 2) * 2 && h->npix() == a.npix, "flowstep: h shape");
  if (hF) BFSR_CHECK(hF->C == 2 * w.C && hF->npix() == a.npix, "flowstep: hF shape");
  a.zin = zin; a.zout = zout;
  a.h = h ? *h : View(); a.hF = hF ? *hF : View();
  a.has_h = h != nullptr; a.has_hF = hF != nullptr;
  a.M = inv ? w.Mi : w.Mf; a.cvec = inv ? w.ci : w.cf; a.MT = inv ? w.MiT : w.MfT;
  a.squeeze_in = sq_in; a.unsqueeze_out = unsq_out;
  a.z1op = View();
  bool z1_done = false;
  a.pix_per_block = w.C <= 24 ? 256 : (w.C <= 96 ? 128 : 32);
  if (a.npix == 0) return;
  const size_t smem = ((size_t)w.C * w.C + (size_t)a.pix_per_block * (w.C + 1)) * 4;
  const int grid = cdiv(a.npix, a.pix_per_block);
  // algorithmic HBM bytes: z in + z out (+ h: C, + hF: 2C) fp32 per level-pixel
  snprintf(g_prof_tag, sizeof g_prof_tag, "flowstep%s C%d %dx%d%s%s", inv ? "-inv" : "-fwd", w.C, a.H, a.W, h ? " h" : "", hF ? " hF" : "");
  ProfScope prof(PK_FLOWSTEP, 4.0 * (double)a.npix * w.C * (2 + (h ? 1 : 0) + (hF ? 2 : 0)), s);
  // small-C fast path (levels 1 and 2 of the shipped topology)
  const bool px_ok = (w.C == 12 || w.C == 24) && zin.fmt == F32 && zout.fmt == F32 && (sq_in || vec4_ok(zin)) &&
                     (unsq_out || vec4_ok(zout)) && (!h || vec4_ok(*h)) && (!hF || vec4_ok(*hF));
  // the fused operand copy of z1 needs the output at level resolution and a BF16X2 view with 16-byte pixel rows
  const bool z1_ok = z1op && !unsq_out && z1op->fmt == BF16X2 && z1op->C == ((w.C / 2 + 7) & ~7) && z1op->cs % 8 == 0 &&
                     z1op->coff % 8 == 0 && z1op->plane % 8 == 0 && ((uintptr_t)z1op->p % 16) == 0 && z1op->npix() == a.npix;
  const bool wide_ok = w.C == 96 && !sq_in && !unsq_out && vec4_ok(zin) && vec4_ok(zout) && (!h || vec4_ok(*h)) && (!hF || vec4_ok(*hF));
  if (wide_ok) {
    if (z1_ok) { a.z1op = *z1op; z1_done = true; }
    const size_t smem96 = (size_t)(96 * 96 + 64 * 100) * 4;
    static int nsm = 0;
    if (!nsm) { int dev = 0; CUDA_OK(cudaGetDevice(&dev)); CUDA_OK(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev)); }
    const int tiles = cdiv(a.npix, 64);
    const int g = tiles < 3 * nsm ? tiles : 3 * nsm;       // three CTAs of 62 KB fit an SM
    if (inv) { CUDA_OK(cudaFuncSetAttribute(flowstep_wide_kernel<96, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem96));
               flowstep_wide_kernel<96, true><<<g, 256, smem96, s>>>(a); }
    else { CUDA_OK(cudaFuncSetAttribute(flowstep_wide_kernel<96, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem96));
           flowstep_wide_kernel<96, false><<<g, 256, smem96, s>>>(a); }
    count_launch();
  } else
  if (px_ok) {
    if (z1_ok) { a.z1op = *z1op; z1_done = true; }
    const int g = cdiv(a.npix, 128);
    if (w.C == 12) { if (inv) flowstep_px_kernel<12, true><<<g, 128, 0, s>>>(a); else flowstep_px_kernel<12, false><<<g, 128, 0, s>>>(a); }
    else { if (inv) flowstep_px_kernel<24, true><<<g, 128, 0, s>>>(a); else flowstep_px_kernel<24, false><<<g, 128, 0, s>>>(a); }
    count_launch();
  } else {
    if (inv) {
      CUDA_OK(cudaFuncSetAttribute(flowstep_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      flowstep_kernel<true><<<grid, 256, smem, s>>>(a);
    } else {
      CUDA_OK(cudaFuncSetAttribute(flowstep_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      flowstep_kernel<false><<<grid, 256, smem, s>>>(a);
    }
    count_launch();
  }
  // operand copy of z1 requested but not producible in the step kernel (boundary steps, odd layouts): plain converting copy
  if (z1op && !z1_done) {
    BFSR_CHECK(!unsq_out, "flowstep: z1 operand copy of an unsqueezed output is not supported");
    resample(zout.slice(0, w.C / 2), *z1op, RS_COPY, s);
  }
}

void flowstep_fwd(const StepW& w, const View& z_in, bool squeeze_in, const View* h_prev, const View* hF,
                  const View& z_out, cudaStream_t s, const View* z1op) {
  launch_step(false, w, z_in, squeeze_in, h_prev, hF, z_out, false, s, z1op);
}
void flowstep_inv(const StepW& w, const View& z_in, const View* h, const View* hF, const View& z_out,
                  bool unsqueeze_out, cudaStream_t s, const View* z1op) {
  launch_step(true, w, z_in, false, h, hF, z_out, unsqueeze_out, s, z1op);
}

// ------------------------------------------------------------------ small elementwise flow ops
__global__ void coupling_finish_kernel(View z, View h, View out, long long n, int C) {
  long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  const long long pix = e / C; const int c = (int)(e % C);
  float v = ld(z, pix, c);
  if (c >= C / 2) { const int j = c - C / 2; v = (v + ld(h, pix, 2 * j)) * ld(h, pix, 2 * j + 1); }
  st(out, pix, c, v);
}
void coupling_finish(const View& z, const View& h, const View& z_out, cudaStream_t s) {
  const long long n = z.npix() * z.C;
  if (!n) return;
  snprintf(g_prof_tag, sizeof g_prof_tag, "coupling_finish C%d", z.C);
  ProfScope prof(PK_OTHER, 8.0 * n, s);
  coupling_finish_kernel<<<cdiv(n, 256), 256, 0, s>>>(z, h, z_out, n, z.C);
  count_launch();
}

__global__ void split_fwd_kernel(View z, View h, View z1, View eps, long long n, int Cp, int Cc) {
  long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  const int C = Cp + Cc;
  const long long pix = e / C; const int c = (int)(e % C);
  const float v = ld(z, pix, c);
  if (c < Cp) { st(z1, pix, c, v); return; }
  const int j = c - Cp;
  const float mean = ld(h, pix, 2 * j), logs = ld(h, pix, 2 * j + 1);
  st(eps, pix, j, (v - mean) / expf(logs));
}
void split_fwd(const View& z, const View& h, const View& z1_out, const View& eps_out, cudaStream_t s) {
  const long long n = z.npix() * z.C;
  if (!n) return;
  snprintf(g_prof_tag, sizeof g_prof_tag, "split_fwd");
  ProfScope prof(PK_OTHER, 8.0 * n, s);
  split_fwd_kernel<<<cdiv(n, 256), 256, 0, s>>>(z, h, z1_out, eps_out, n, z1_out.C, eps_out.C);
  count_launch();
}

__global__ void split_inv_kernel(View z1, View h, View eps, View out, long long n, int Cp, int Cc) {
  long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  const int C = Cp + Cc;
  const long long pix = e / C; const int c = (int)(e % C);
  if (c < Cp) { st(out, pix, c, ld(z1, pix, c)); return; }
  const int j = c - Cp;
  const float mean = ld(h, pix, 2 * j), logs = ld(h, pix, 2 * j + 1);
  st(out, pix, c, mean + expf(logs) * ld(eps, pix, j));
}
void split_inv(const View& z1, const View& h, const View& eps, const View& z_out, cudaStream_t s) {
  const long long n = z_out.npix() * z_out.C;
  if (!n) return;
  snprintf(g_prof_tag, sizeof g_prof_tag, "split_inv");
  ProfScope prof(PK_OTHER, 8.0 * n, s);
  split_inv_kernel<<<cdiv(n, 256), 256, 0, s>>>(z1, h, eps, z_out, n, z1.C, eps.C);
  count_launch();
}

__global__ void normalise_kernel(View e, View out, long long npix, int C) {
  long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= npix) return;
  float s = 0.f;
  for (int c = 0; c < C; ++c) s += ld(e, pix, c);
  const float mean = s / C;
  float q = 0.f;
  for (int c = 0; c < C; ++c) { const float d = ld(e, pix, c) - mean; q = fmaf(d, d, q); }
  const float sd = sqrtf(q / (C - 1)) + 1e-8f;
  for (int c = 0; c < C; ++c) st(out, pix, c, (ld(e, pix, c) - mean) / sd);
}
void normalise_latent(const View& e, const View& out, cudaStream_t s) {
  if (!e.npix()) return;
  snprintf(g_prof_tag, sizeof g_prof_tag, "normalise C%d", e.C);
  ProfScope prof(PK_OTHER, 8.0 * e.npix() * e.C, s);
  normalise_kernel<<<cdiv(e.npix(), 128), 128, 0, s>>>(e, out, e.npix(), e.C);
  count_launch();
}

__global__ void squeeze_copy_kernel(View src, View dst, long long n, int C, int H, int W, int dir) {
  long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  // (pix, c) indexes the SQUEEZED tensor (C channels, H x W)
  const long long pix = e / C; const int c = (int)(e % C);
  const int j = (int)(pix % W); const long long t = pix / W; const int i = (int)(t % H); const long long b = t / H;
  const int cc = c >> 2, fh = (c >> 1) & 1, fw = c & 1;
  const long long up = (b * (2 * H) + (2 * i + fh)) * (2 * W) + (2 * j + fw);
  if (dir == 0) st(dst, pix, c, ld(src, up, cc));
  else st(dst, up, cc, ld(src, pix, c));
}
void squeeze_copy(const View& src, const View& dst, cudaStream_t s) {
  const long long n = dst.npix() * dst.C;
  if (!n) return;
  squeeze_copy_kernel<<<cdiv(n, 256), 256, 0, s>>>(src, dst, n, dst.C, dst.H, dst.W, 0);
  count_launch();
}
void unsqueeze_copy(const View& src, const View& dst, cudaStream_t s) {
  const long long n = src.npix() * src.C;
  if (!n) return;
  squeeze_copy_kernel<<<cdiv(n, 256), 256, 0, s>>>(src, dst, n, src.C, src.H, src.W, 1);
  count_launch();
}

// ------------------------------------------------------------------ layout
__global__ void nchw_to_nhwc_kernel(const float* src, View dst, long long n, int C, int HW) {
  long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  const long long pix = e / C; const int c = (int)(e % C);
  const long long b = pix / HW, r = pix % HW;
  st(dst, pix, c, src[(b * C + c) * HW + r]);
}
void nchw_to_nhwc(const float* src, const View& dst, cudaStream_t s) {
  const long long n = dst.npix() * dst.C;
  if (!n) return;
  snprintf(g_prof_tag, sizeof g_prof_tag, "nchw_to_nhwc C%d", dst.C);
  ProfScope prof(PK_OTHER, 8.0 * n, s);
  nchw_to_nhwc_kernel<<<cdiv(n, 256), 256, 0, s>>>(src, dst, n, dst.C, dst.H * dst.W);
  count_launch();
}
__global__ void nhwc_to_nchw_kernel(View src, float* dst, long long n, int C, int HW) {
  long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  // e indexes the NCHW output so the stores coalesce
  const long long r = e % HW; const long long t = e / HW; const int c = (int)(t % C); const long long b = t / C;
  dst[e] = ld(src, b * HW + r, c);
}
void nhwc_to_nchw(const View& src, float* dst, cudaStream_t s) {
  const long long n = src.npix() * src.C;
  if (!n) return;
  snprintf(g_prof_tag, sizeof g_prof_tag, "nhwc_to_nchw C%d", src.C);
  ProfScope prof(PK_OTHER, 8.0 * n, s);
  nhwc_to_nchw_kernel<<<cdiv(n, 256), 256, 0, s>>>(src, dst, n, src.C, src.H * src.W);
  count_launch();
}

// ------------------------------------------------------------------ resampling
__global__ void resample_kernel(View src, View dst, long long n, int mode, int oy, int ox) {
  long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  const int C = dst.C;
  const long long pix = e / C; const int c = (int)(e % C);
  const int x = (int)(pix % dst.W); const long long t = pix / dst.W; const int y = (int)(t % dst.H); const long long b = t / dst.H;
  const long long sb = b * src.H * src.W;
  float v;
  if (mode == RS_COPY) v = c < src.C ? ld(src, pix, c) : 0.f;   // extra dst channels are zero padding
  else if (mode == RS_NEAREST_UP2) v = ld(src, sb + (long long)(y >> 1) * src.W + (x >> 1), c);
  else if (mode == RS_NEAREST_DOWN2) v = ld(src, sb + (long long)(2 * y) * src.W + 2 * x, c);
  else if (mode == RS_AVG_DOWN2) {
    // bilinear x0.5, align_corners=False, recompute_scale_factor=True (RRDBNet_arch.py:134): every weight is 0.5
    const long long p = sb + (long long)(2 * y) * src.W + 2 * x;
    const float t0 = 0.5f * ld(src, p, c) + 0.5f * ld(src, p + 1, c);
    const float t1 = 0.5f * ld(src, p + src.W, c) + 0.5f * ld(src, p + src.W + 1, c);
    v = 0.5f * t0 + 0.5f * t1;
  } else if (mode == RS_MAXPOOL2) {
    const long long p = sb + (long long)(2 * y) * src.W + 2 * x;
    v = fmaxf(fmaxf(ld(src, p, c), ld(src, p + 1, c)), fmaxf(ld(src, p + src.W, c), ld(src, p + src.W + 1, c)));
  } else {  // RS_BILINEAR_UP2_AC: nn.Upsample(x2, bilinear, align_corners=True) then F.pad to dst (unet.py:80-91)
    const int uy = y - oy, ux = x - ox, UH = 2 * src.H, UW = 2 * src.W;
    if (uy < 0 || uy >= UH || ux < 0 || ux >= UW) v = 0.f;
    else {
      const float sh = UH > 1 ? (float)(src.H - 1) / (float)(UH - 1) : 0.f;
      const float sw = UW > 1 ? (float)(src.W - 1) / (float)(UW - 1) : 0.f;
      const float fy = sh * uy, fx = sw * ux;
      const int y0 = (int)fy, x0 = (int)fx;
      const int y1 = y0 + (y0 < src.H - 1 ? 1 : 0), x1 = x0 + (x0 < src.W - 1 ? 1 : 0);
      const float ly = fy - y0, lx = fx - x0;
      const float a00 = ld(src, sb + (long long)y0 * src.W + x0, c), a01 = ld(src, sb + (long long)y0 * src.W + x1, c);
      const float a10 = ld(src, sb + (long long)y1 * src.W + x0, c), a11 = ld(src, sb + (long long)y1 * src.W + x1, c);
      v = (1.f - ly) * ((1.f - lx) * a00 + lx * a01) + ly * ((1.f - lx) * a10 + lx * a11);
    }
  }
  st(dst, pix, c, v);
}
// 8 consecutive channels of one pixel (views with 8-channel-aligned strides / offsets): two 128-bit loads for fp32, one per
// plane for BF16X2
__device__ __forceinline__ void ld8(const View& v, long long pix, int c, float* x) {
  if (v.fmt == F32) {
    const float4* p = reinterpret_cast<const float4*>((const float*)v.p + pix * v.cs + v.coff + c);
    const float4 a = __ldg(p), b = __ldg(p + 1);
    x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w; x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
  } else {
    const __nv_bfloat16* p = (const __nv_bfloat16*)v.p + pix * v.cs + v.coff + c;
    const uint4 h = __ldg(reinterpret_cast<const uint4*>(p)), l = __ldg(reinterpret_cast<const uint4*>(p + v.plane));
    const uint32_t hh[4] = {h.x, h.y, h.z, h.w}, ll[4] = {l.x, l.y, l.z, l.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {   // bf16 -> fp32 is a 16-bit shift
      x[2 * i] = __uint_as_float(hh[i] << 16) + __uint_as_float(ll[i] << 16);
      x[2 * i + 1] = __uint_as_float(hh[i] & 0xffff0000u) + __uint_as_float(ll[i] & 0xffff0000u);
    }
  }
}
__device__ __forceinline__ void st8(const View& v, long long pix, int c, const float* x) {
  if (v.fmt == F32) {
    float4* p = reinterpret_cast<float4*>((float*)v.p + pix * v.cs + v.coff + c);
    p[0] = make_float4(x[0], x[1], x[2], x[3]); p[1] = make_float4(x[4], x[5], x[6], x[7]);
  } else {
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const __nv_bfloat162 hh = __floats2bfloat162_rn(x[2 * i], x[2 * i + 1]);
      const __nv_bfloat162 ll = __floats2bfloat162_rn(x[2 * i] - __low2float(hh), x[2 * i + 1] - __high2float(hh));
      hi[i] = *reinterpret_cast<const uint32_t*>(&hh); lo[i] = *reinterpret_cast<const uint32_t*>(&ll);
    }
    __nv_bfloat16* p = (__nv_bfloat16*)v.p + pix * v.cs + v.coff + c;
    *reinterpret_cast<uint4*>(p) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4*>(p + v.plane) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
  }
}
// channel-vectorised twin of resample_kernel: one thread = 8 channels of one output pixel
__global__ void resample8_kernel(View src, View dst, long long n8, int mode, int oy, int ox) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n8) return;
  const int C8 = dst.C >> 3;
  const long long pix = e / C8; const int c = (int)(e % C8) << 3;
  const int x = (int)(pix % dst.W); const long long t = pix / dst.W; const int y = (int)(t % dst.H); const long long b = t / dst.H;
  const long long sb = b * src.H * src.W;
  float v[8];
  if (mode == RS_COPY) ld8(src, pix, c, v);
  else if (mode == RS_NEAREST_UP2) ld8(src, sb + (long long)(y >> 1) * src.W + (x >> 1), c, v);
  else if (mode == RS_NEAREST_DOWN2) ld8(src, sb + (long long)(2 * y) * src.W + 2 * x, c, v);
  else if (mode == RS_AVG_DOWN2 || mode == RS_MAXPOOL2) {
    const long long p = sb + (long long)(2 * y) * src.W + 2 * x;
    float a[8], bq[8], cq[8], d[8];
    ld8(src, p, c, a); ld8(src, p + 1, c, bq); ld8(src, p + src.W, c, cq); ld8(src, p + src.W + 1, c, d);
#pragma unroll
    for (int i = 0; i < 8; ++i)
      v[i] = mode == RS_MAXPOOL2 ? fmaxf(fmaxf(a[i], bq[i]), fmaxf(cq[i], d[i]))
                                 : 0.5f * (0.5f * a[i] + 0.5f * bq[i]) + 0.5f * (0.5f * cq[i] + 0.5f * d[i]);   // same op order as the scalar kernel
  } else {  // RS_BILINEAR_UP2_AC
    const int uy = y - oy, ux = x - ox, UH = 2 * src.H, UW = 2 * src.W;
    if (uy < 0 || uy >= UH || ux < 0 || ux >= UW) {
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = 0.f;
    } else {
      const float sh = UH > 1 ? (float)(src.H - 1) / (float)(UH - 1) : 0.f;
      const float sw = UW > 1 ? (float)(src.W - 1) / (float)(UW - 1) : 0.f;
      const float fy = sh * uy, fx = sw * ux;
      const int y0 = (int)fy, x0 = (int)fx;
      const int y1 = y0 + (y0 < src.H - 1 ? 1 : 0), x1 = x0 + (x0 < src.W - 1 ? 1 : 0);
      const float ly = fy - y0, lx = fx - x0;
      float a00[8], a01[8], a10[8], a11[8];
      ld8(src, sb + (long long)y0 * src.W + x0, c, a00); ld8(src, sb + (long long)y0 * src.W + x1, c, a01);
      ld8(src, sb + (long long)y1 * src.W + x0, c, a10); ld8(src, sb + (long long)y1 * src.W + x1, c, a11);
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = (1.f - ly) * ((1.f - lx) * a00[i] + lx * a01[i]) + ly * ((1.f - lx) * a10[i] + lx * a11[i]);
    }
  }
  st8(dst, pix, c, v);
}
static bool vec8_ok(const View& v) {
  return v.C % 8 == 0 && v.cs % 8 == 0 && v.coff % 8 == 0 && ((uintptr_t)v.p % 16) == 0 && (v.fmt == F32 || v.plane % 8 == 0);
}

void resample(const View& src, const View& dst, int mode, cudaStream_t s) {
  BFSR_CHECK((src.C == dst.C || (mode == RS_COPY && src.C < dst.C)) && src.N == dst.N,
             "resample: channel/batch mismatch (%d vs %d)", src.C, dst.C);
  int oy = 0, ox = 0;
  switch (mode) {
    case RS_COPY: BFSR_CHECK(src.H == dst.H && src.W == dst.W, "resample copy: shape"); break;
    case RS_NEAREST_UP2: BFSR_CHECK(dst.H == 2 * src.H && dst.W == 2 * src.W, "resample up2: shape"); break;
    case RS_NEAREST_DOWN2: case RS_AVG_DOWN2:
      BFSR_CHECK(src.H == 2 * dst.H && src.W == 2 * dst.W, "resample down2: shape"); break;
    case RS_MAXPOOL2: BFSR_CHECK(dst.H == src.H / 2 && dst.W == src.W / 2, "maxpool2: shape"); break;
    case RS_BILINEAR_UP2_AC:
      BFSR_CHECK(dst.H >= 2 * src.H && dst.W >= 2 * src.W, "bilinear up2: dst smaller than upsampled src");
      oy = (dst.H - 2 * src.H) / 2; ox = (dst.W - 2 * src.W) / 2; break;
    default: BFSR_CHECK(false, "resample: bad mode %d", mode);
  }
  const long long n = dst.npix() * dst.C;
  if (!n) return;
  snprintf(g_prof_tag, sizeof g_prof_tag, "resample m%d C%d %dx%d", mode, dst.C, dst.H, dst.W);
  ProfScope prof(PK_OTHER, 8.0 * n, s);
  if (src.C == dst.C && vec8_ok(src) && vec8_ok(dst)) resample8_kernel<<<cdiv(n / 8, 256), 256, 0, s>>>(src, dst, n / 8, mode, oy, ox);
  else resample_kernel<<<cdiv(n, 256), 256, 0, s>>>(src, dst, n, mode, oy, ox);
  count_launch();
}

// F.interpolate(lr, scale_factor=s, mode='bilinear', align_corners=False)  — SRFlow-LP/code/test.py:137
__global__ void bilinear_up_nchw_kernel(const float* src, int C, int h, int w, int scale, View dst, long long n) {
  long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  const long long pix = e / C; const int c = (int)(e % C);
  const int x = (int)(pix % dst.W); const long long t = pix / dst.W; const int y = (int)(t % dst.H); const long long b = t / dst.H;
  const float rs = 1.f / (float)scale;
  float fy = rs * (y + 0.5f) - 0.5f, fx = rs * (x + 0.5f) - 0.5f;
  fy = fy < 0.f ? 0.f : fy; fx = fx < 0.f ? 0.f : fx;
  const int y0 = (int)fy, x0 = (int)fx;
  const int y1 = y0 + (y0 < h - 1 ? 1 : 0), x1 = x0 + (x0 < w - 1 ? 1 : 0);
  const float ly = fy - y0, lx = fx - x0;
  const float* sp = src + (b * C + c) * (long long)h * w;
  const float a00 = sp[y0 * w + x0], a01 = sp[y0 * w + x1], a10 = sp[y1 * w + x0], a11 = sp[y1 * w + x1];
  st(dst, pix, c, (1.f - ly) * ((1.f - lx) * a00 + lx * a01) + ly * ((1.f - lx) * a10 + lx * a11));
}
void bilinear_up_nchw(const float* src, int N, int C, int h, int w, int scale, const View& dst, cudaStream_t s) {
  BFSR_CHECK(dst.N == N && dst.C == C && dst.H == h * scale && dst.W == w * scale, "bilinear_up: shape");
  const long long n = dst.npix() * C;
  if (!n) return;
  snprintf(g_prof_tag, sizeof g_prof_tag, "bilinear_up");
  ProfScope prof(PK_OTHER, 4.0 * n, s);
  bilinear_up_nchw_kernel<<<cdiv(n, 256), 256, 0, s>>>(src, C, h, w, scale, dst, n);
  count_launch();
}

}  // namespace bfsr
