// tcgen05 implicit-GEMM 3x3 convolution for sm_100a (replaces cuDNN behind every wide nn.Conv2d of the path:
// RRDBNet_arch.py:25-45, the feature-only coupling convs of FlowAffineCouplingsAblation.py:45-55, unet.py:10-107).
//
// GEMM view per CTA: M = 128 output pixels (8 wide x 16 tall), N = NT <= 64 output channels, K = 9 taps x Cin.
//
//  * A operand (pixels x channels).  The fp32 NHWC halo tile (18 x 10 pixels) of one 64-channel chunk is loaded ONCE
//    by the producer warps with coalesced 128-bit loads, split on the fly into bf16 (hi, lo) planes and written to
//    shared memory in the UMMA K-major SWIZZLE_128B layout (one pixel = one 128-byte row).  The 9 taps are then just
//    9 shifted VIEWS of that tile: the smem descriptor's start address moves by (dy*10+dx) rows and its stride-byte-
//    offset (distance between 8-row groups = one image row of the tile) is the halo pitch, 1280 B.  The swizzle is a
//    function of the absolute smem address (Swizzle<3,4,3> o smem_ptr), so shifted views stay consistent.  Zero
//    padding of the convolution and of ragged tiles is written as zeros by the producer.
//  * B operand (weights), pre-packed at load time as [cout tile][chunk][tap] images of the exact smem layout
//    ([W_hi ; W_lo] rows of 128 B, swizzled), streamed through a 4-slot ring with cp.async.bulk (TMA bulk copy,
//    SASS UBLKCP) completing on mbarriers.
//  * fp32-accurate arithmetic on bf16 tensor cores (split-bf16 x3):
//        x*w ~= x_hi*w_hi + x_hi*w_lo + x_lo*w_hi      (dropped term ~2^-16 relative)
//    issued as TWO tcgen05.mma per K=16 step:  A_hi x [W_hi;W_lo] (N = 2*NT, accumulator columns [0,2NT)) and
//    A_lo x W_hi (N = NT, accumulating into columns [0,NT)); the epilogue adds the two column halves.  The fast mode
//    issues only A_hi x W_hi.
//  * Accumulators live in TMEM; one elected thread issues the MMAs; tcgen05.commit arrives on the mbarriers that
//    recycle the A / W slots and that release the epilogue.  Epilogue: tcgen05.ld 32x32b -> bias, pre-activation add,
//    activation, scaled residuals -> fp32 NHWC channel-slice store (same fused epilogue as the fp32 kernel).
//
// Warp roles (192 threads): warps 0-3 = A producers, then epilogue (TMEM lane quarter = warp id); warp 4 = TMEM
// allocator + MMA issuer; warp 5 = weight TMA producer.
#include "ops.cuh"
#include <vector>
#include <cstring>
#include <cmath>

namespace bfsr {

thread_local int g_conv_mode = 0;   // 0 = split-bf16 x3 on tcgen05 (accurate), 1 = bf16 (fast), 2 = fp32 CUDA cores only

namespace tc {
constexpr int TW = 8, TH = 16, PITCH = 10, HROWS = TH + 2, HPIX = HROWS * PITCH;   // 180 halo pixels
constexpr int KC = 64;                       // channels per chunk = one 128-byte bf16 row
constexpr int A_PLANE = 23552;               // >= HPIX*128, multiple of 1024
constexpr int A_SLOT = 2 * A_PLANE;
constexpr int NA = 2, NW = 4;
constexpr int W_SLOT_MAX = 2 * 64 * 128;     // [W_hi;W_lo] for NT = 64
constexpr int SMEM_BYTES = NA * A_SLOT + NW * W_SLOT_MAX + 1024 /*align slack*/ + 256 /*barriers*/;
constexpr int NPROD = 128;
}  // namespace tc

struct TcArgs {
  View in, out, pre, res1, res2;
  const unsigned char* w; const float* bias;
  int cin, cout, nt, n_chunks;
  int H, W, in_mode, act;
  float eps, alpha, beta1, beta2;
  int tiles_x, fast, vec_in, vec_out;
};

// ------------------------------------------------------------------ PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
// bounded wait: a protocol bug traps instead of hanging the device
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) { printf("bfsr conv_tc: mbarrier timeout (block %d,%d,%d thread %d)\n", blockIdx.x, blockIdx.y, blockIdx.z, threadIdx.x); __trap(); }
  }
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t addr, float* v) {
  uint32_t r[16];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                 "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(addr));
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
// K-major SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor, version 1)
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t sbo_bytes) {
  return (uint64_t)((addr & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) |
         ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
// kind::f16 instruction descriptor: D=f32, A=B=bf16, both K-major, M=128
__device__ __forceinline__ uint32_t make_idesc(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&t);
}

__global__ void __launch_bounds__(192, 1) conv_tc_kernel(TcArgs a) {
  using namespace tc;
  extern __shared__ unsigned char smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_smem = base;                               // NA slots of [hi plane | lo plane]
  const uint32_t w_smem = base + NA * A_SLOT;                 // NW slots
  const uint32_t bars = w_smem + NW * W_SLOT_MAX;             // mbarriers (8 B each)
  const uint32_t a_full = bars, a_empty = bars + 8 * NA, w_full = bars + 16 * NA, w_empty = w_full + 8 * NW;
  const uint32_t acc_full = w_empty + 8 * NW, tmem_slot = acc_full + 8;
  unsigned char* smem_gen = smem_raw + (base - smem_u32(smem_raw));

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int tile = blockIdx.x;
  const int ty0 = (tile / a.tiles_x) * TH, tx0 = (tile % a.tiles_x) * TW;
  const int nt = a.nt, co_base = blockIdx.y * nt;
  const int n = blockIdx.z;
  const int w_rows = a.fast ? nt : 2 * nt;
  const uint32_t w_bytes = (uint32_t)w_rows * 128u;
  uint32_t ncols = 32; while ((int)ncols < (a.fast ? nt : 2 * nt)) ncols <<= 1;

  if (tid == 0) {
    for (int i = 0; i < NA; ++i) { mbar_init(a_full + 8 * i, NPROD); mbar_init(a_empty + 8 * i, 1); }
    for (int i = 0; i < NW; ++i) { mbar_init(w_full + 8 * i, 1); mbar_init(w_empty + 8 * i, 1); }
    mbar_init(acc_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp < 4) {
    // ===================== A producers: fp32 halo tile -> (hi, lo) bf16 planes, swizzled =====================
    const int inH = a.in_mode == IN_UP2 ? a.H >> 1 : a.H, inW = a.in_mode == IN_UP2 ? a.W >> 1 : a.W;
    const long long img = (long long)n * inH * inW;
    for (int c = 0; c < a.n_chunks; ++c) {
      const int slot = c % NA;
      mbar_wait(a_empty + 8 * slot, ((c / NA) & 1) ^ 1);
      unsigned char* pl_hi = smem_gen + slot * A_SLOT;
      constexpr int ITEMS = HPIX * 8;
      constexpr int U = 4;
      for (int it0 = tid; it0 < ITEMS; it0 += NPROD * U) {
        float4 v[U][2];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int it = it0 + u * NPROD;
          v[u][0] = make_float4(0.f, 0.f, 0.f, 0.f); v[u][1] = v[u][0];
          if (it < ITEMS) {
            const int q = it >> 3, j = it & 7;
            const int yy = q / PITCH, xx = q - yy * PITCH;
            const int gy = ty0 + yy - 1, gx = tx0 + xx - 1;
            const int cb = c * KC + j * 8;
            if (gy >= 0 && gy < a.H && gx >= 0 && gx < a.W && cb < a.cin) {
              const int sy = a.in_mode == IN_UP2 ? gy >> 1 : gy, sx = a.in_mode == IN_UP2 ? gx >> 1 : gx;
              const long long p = img + (long long)sy * inW + sx;
              if (a.vec_in && cb + 8 <= a.cin) {
                const float4* src = reinterpret_cast<const float4*>((const float*)a.in.p + p * a.in.cs + a.in.coff + cb);
                v[u][0] = __ldg(src); v[u][1] = __ldg(src + 1);
              } else {
                float t[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) t[e] = cb + e < a.cin ? ld(a.in, p, cb + e) : 0.f;
                v[u][0] = make_float4(t[0], t[1], t[2], t[3]); v[u][1] = make_float4(t[4], t[5], t[6], t[7]);
              }
            }
          }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int it = it0 + u * NPROD;
          if (it < ITEMS) {
            const int q = it >> 3, j = it & 7;
            const float x[8] = {v[u][0].x, v[u][0].y, v[u][0].z, v[u][0].w, v[u][1].x, v[u][1].y, v[u][1].z, v[u][1].w};
            float hi[8], lo[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) { hi[e] = __bfloat162float(__float2bfloat16_rn(x[e])); lo[e] = x[e] - hi[e]; }
            const uint32_t off = (uint32_t)q * 128u + (uint32_t)((j ^ (q & 7)) << 4);
            *reinterpret_cast<uint4*>(pl_hi + off) =
                make_uint4(pack_bf16(hi[0], hi[1]), pack_bf16(hi[2], hi[3]), pack_bf16(hi[4], hi[5]), pack_bf16(hi[6], hi[7]));
            if (!a.fast)
              *reinterpret_cast<uint4*>(pl_hi + A_PLANE + off) =
                  make_uint4(pack_bf16(lo[0], lo[1]), pack_bf16(lo[2], lo[3]), pack_bf16(lo[4], lo[5]), pack_bf16(lo[6], lo[7]));
          }
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the tensor core
      mbar_arrive(a_full + 8 * slot);
    }
    // ===================== epilogue: TMEM -> registers -> fused epilogue -> global =====================
    mbar_wait(acc_full, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int row = warp * 32 + lane;                 // accumulator row = TMEM lane
    const int gy = ty0 + (row >> 3), gx = tx0 + (row & 7);
    const bool valid = gy < a.H && gx < a.W;
    const long long p = ((long long)n * a.H + gy) * a.W + gx;
    const uint32_t t_row = tmem_base + ((uint32_t)(warp * 32) << 16);
    for (int n0 = 0; n0 < nt; n0 += 16) {
      float acc[16], acc2[16];
      tmem_ld16(t_row + n0, acc);
      if (!a.fast) tmem_ld16(t_row + nt + n0, acc2);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      if (!a.fast) {
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[i] += acc2[i];
      }
      if (valid) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int co = co_base + n0 + i;
          if (co < a.cout) {
            float v = acc[i] + a.bias[co];
            if (a.pre.p) v += ld(a.pre, p, co);
            if (a.act == ACT_LRELU) v = v > 0.f ? v : 0.2f * v;
            else if (a.act == ACT_RELU) v = fmaxf(v, 0.f);
            else if (a.act == ACT_CROSS_SIGMOID) { if (co & 1) v = 1.f / (1.f + expf(-(v + 2.f))) + a.eps; }
            v *= a.alpha;
            if (a.res1.p) v = fmaf(a.beta1, ld(a.res1, p, co), v);
            if (a.res2.p) v = fmaf(a.beta2, ld(a.res2, p, co), v);
            acc[i] = v;
          }
        }
        const int co = co_base + n0;
        if (a.vec_out && co + 16 <= a.cout) {
          float4* dst = reinterpret_cast<float4*>((float*)a.out.p + p * a.out.cs + a.out.coff + co);
#pragma unroll
          for (int i = 0; i < 4; ++i) dst[i] = make_float4(acc[4 * i], acc[4 * i + 1], acc[4 * i + 2], acc[4 * i + 3]);
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) if (co + i < a.cout) st(a.out, p, co + i, acc[i]);
        }
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  } else if (warp == 4) {
    // ===================== MMA issuer (one elected lane) =====================
    if (lane == 0) {
      const uint32_t idesc_wide = make_idesc(a.fast ? nt : 2 * nt), idesc_nt = make_idesc(nt);
      int wi = 0;
      uint32_t first = 1;
      for (int c = 0; c < a.n_chunks; ++c) {
        const int slot = c % NA;
        mbar_wait(a_full + 8 * slot, (c / NA) & 1);
        const uint32_t a_hi = a_smem + slot * A_SLOT, a_lo = a_hi + A_PLANE;
        int nk = (a.cin - c * KC + 15) >> 4; nk = nk > 4 ? 4 : nk;
        for (int t = 0; t < 9; ++t, ++wi) {
          const int ws = wi % NW;
          mbar_wait(w_full + 8 * ws, (wi / NW) & 1);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t shift = (uint32_t)((t / 3) * PITCH + (t % 3)) * 128u;
          const uint32_t wb = w_smem + ws * W_SLOT_MAX;
          for (int kk = 0; kk < nk; ++kk) {
            const uint64_t bd = make_desc(wb + kk * 32, 1024);
            umma_f16(tmem_base, make_desc(a_hi + shift + kk * 32, PITCH * 128), bd, idesc_wide, first ? 0u : 1u);
            first = 0;
            if (!a.fast) umma_f16(tmem_base, make_desc(a_lo + shift + kk * 32, PITCH * 128), bd, idesc_nt, 1u);
          }
          umma_commit(w_empty + 8 * ws);      // weight slot reusable once these MMAs retire
        }
        umma_commit(a_empty + 8 * slot);      // A slot reusable
      }
      umma_commit(acc_full);
    }
    __syncwarp();
  } else {
    // ===================== weight producer: cp.async.bulk of pre-swizzled [W_hi;W_lo] images =====================
    if (lane == 0) {
      const size_t tap_stride = (size_t)2 * nt * 128;    // packed image always holds both planes
      const unsigned char* wsrc = a.w + (size_t)blockIdx.y * a.n_chunks * 9 * tap_stride;
      const int total = a.n_chunks * 9;
      for (int wi = 0; wi < total; ++wi) {
        const int ws = wi % NW;
        mbar_wait(w_empty + 8 * ws, ((wi / NW) & 1) ^ 1);
        mbar_expect_tx(w_full + 8 * ws, w_bytes);
        bulk_g2s(w_smem + ws * W_SLOT_MAX, wsrc + (size_t)wi * tap_stride, w_bytes, w_full + 8 * ws);
      }
    }
    __syncwarp();
  }
  __syncthreads();
  if (warp == 4) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(ncols) : "memory");
  }
}

// ------------------------------------------------------------------ host side
static inline unsigned short f2bf(float x) {   // round-to-nearest-even
  uint32_t u; memcpy(&u, &x, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return (unsigned short)((u >> 16) | 0x40);
  u += 0x7fffu + ((u >> 16) & 1u);
  return (unsigned short)(u >> 16);
}
static inline float bf2f(unsigned short b) { uint32_t u = (uint32_t)b << 16; float x; memcpy(&x, &u, 4); return x; }

// h: host fp32 packed [tap][cin_pad][cout_pad] (the fp32 kernel's layout)
void pack_conv_tc(ConvW& c, const std::vector<float>& h) {
  using namespace tc;
  if (c.ks != 3 || c.cin < 32) return;
  const int nt = c.cout > 48 ? 64 : (c.cout + 15) / 16 * 16;
  const int n_tiles = (c.cout + nt - 1) / nt, n_chunks = (c.cin + KC - 1) / KC;
  const size_t tap_bytes = (size_t)2 * nt * 128;
  std::vector<unsigned short> img((size_t)n_tiles * n_chunks * 9 * tap_bytes / 2, 0);
  for (int t = 0; t < n_tiles; ++t)
    for (int ch = 0; ch < n_chunks; ++ch)
      for (int tap = 0; tap < 9; ++tap) {
        unsigned short* dst = img.data() + (((size_t)t * n_chunks + ch) * 9 + tap) * tap_bytes / 2;
        for (int r = 0; r < nt; ++r) {
          const int co = t * nt + r;
          for (int k = 0; k < KC; ++k) {
            const int ci = ch * KC + k;
            float w = 0.f;
            if (co < c.cout && ci < c.cin) w = h[((size_t)tap * c.cin_pad + ci) * c.cout_pad + co];
            const unsigned short hi = f2bf(w), lo = f2bf(w - bf2f(hi));
            const int j = k >> 3, e = k & 7;
            dst[(size_t)r * 64 + ((j ^ (r & 7)) << 3) + e] = hi;
            const int r2 = nt + r;
            dst[(size_t)r2 * 64 + ((j ^ (r2 & 7)) << 3) + e] = lo;
          }
        }
      }
  CUDA_OK(cudaMalloc(&c.w_tc, img.size() * 2));
  CUDA_OK(cudaMemcpy(c.w_tc, img.data(), img.size() * 2, cudaMemcpyHostToDevice));
  c.tc_kchunks = n_chunks; c.tc_npad = nt;
}

bool conv_tc_eligible(const ConvW& w, const View& in, const View& out) {
  return w.w_tc != nullptr && g_conv_mode != 2 && in.fmt == F32 && out.npix() >= 128;
}

void conv2d_tc(const ConvW& w, const View& in, const View& out, const ConvEpi& epi, int in_mode, cudaStream_t s) {
  using namespace tc;
  BFSR_CHECK(w.w_tc, "conv_tc: weights not packed for the tcgen05 path");
  BFSR_CHECK(in.C == w.cin && out.C == w.cout && in.N == out.N, "conv_tc: shape mismatch");
  if (in_mode == IN_UP2) BFSR_CHECK(in.H * 2 == out.H && in.W * 2 == out.W, "conv_tc(up2): spatial mismatch");
  else BFSR_CHECK(in.H == out.H && in.W == out.W, "conv_tc: spatial mismatch");
  TcArgs a;
  a.in = in; a.out = out;
  a.pre = epi.pre ? *epi.pre : View(); a.res1 = epi.res1 ? *epi.res1 : View(); a.res2 = epi.res2 ? *epi.res2 : View();
  a.w = (const unsigned char*)w.w_tc; a.bias = w.bias;
  a.cin = w.cin; a.cout = w.cout; a.nt = w.tc_npad; a.n_chunks = w.tc_kchunks;
  a.H = out.H; a.W = out.W; a.in_mode = in_mode; a.act = epi.act;
  a.eps = epi.eps; a.alpha = epi.alpha; a.beta1 = epi.beta1; a.beta2 = epi.beta2;
  a.tiles_x = cdiv(out.W, TW);
  a.fast = g_conv_mode == 1;
  a.vec_in = (in.fmt == F32 && in.cs % 4 == 0 && in.coff % 4 == 0 && ((uintptr_t)in.p % 16) == 0);
  a.vec_out = (out.fmt == F32 && out.cs % 4 == 0 && out.coff % 4 == 0 && ((uintptr_t)out.p % 16) == 0);
  if (out.npix() == 0) return;
  CUDA_OK(cudaFuncSetAttribute(conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
  dim3 grid(a.tiles_x * cdiv(out.H, TH), cdiv(w.cout, w.tc_npad), out.N);
  ProfScope prof(PK_CONV_TC, 2.0 * (double)out.npix() * w.cin * 9 * w.cout, s);
  conv_tc_kernel<<<grid, 192, SMEM_BYTES, s>>>(a);
  count_launch();
}

}  // namespace bfsr
