// tcgen05 implicit-GEMM 3x3 / 1x1 convolution for sm_100a (replaces cuDNN behind every wide nn.Conv2d of the path:
// RRDBNet_arch.py:25-45, the coupling convs of FlowAffineCouplingsAblation.py:45-55, unet.py:10-107, linf.py:228-240).
//
// Persistent, warp-specialised kernel.  One CTA per SM walks a static round-robin list of work tiles; a work tile is a
// MACRO tile of MT sub-tiles (each 8 wide x 16 tall = 128 output pixels = one UMMA M=128 accumulator) times NT output
// channels.  GEMM view per sub-tile: M = 128 pixels, N = NT, K = taps x Cin, walked in 32-channel chunks.
//
//  * A operand (pixels x channels).  One halo tile of the whole macro tile (e.g. 18 x 34 pixels for 2x2 sub-tiles) per
//    32-channel chunk lands in shared memory in the UMMA K-major SWIZZLE_64B layout (one pixel = one 64-byte row), as bf16
//    (hi, lo) planes.  Inputs stored in HBM in that split format (BF16X2 views: every conv-operand-only tensor of the
//    engines) arrive by ONE TMA tile load per plane (kernel variant TMA_IN; zero OOB fill = the conv's padding); fp32 /
//    nearest-2x-upsampled inputs are loaded, split and written by 8 producer warps (variant !TMA_IN).  Every (sub-tile,
//    tap) operand is then a shifted VIEW of that tile: the smem descriptor's start address moves by
//    ((sy*16+dy)*pitch + sx*8+dx) rows and its stride-byte-offset (distance between 8-row groups = one image row of
//    the tile) is the halo pitch.  The swizzle is a function of the absolute smem address, so shifted views stay
//    consistent.
//  * B operand (weights), pre-packed at load time as [cout tile][chunk][tap] images of the exact smem layout
//    ([W_hi ; W_lo] rows of 64 B, swizzled), streamed through a 2-4 stage ring with cp.async.bulk (SASS UBLKCP)
//    completing on mbarriers; one weight stage serves all MT sub-tiles of the macro tile.
//  * fp32-accurate arithmetic on bf16 tensor cores (split-bf16 x3):
//        x*w ~= x_hi*w_hi + x_hi*w_lo + x_lo*w_hi      (dropped term ~2^-16 relative)
//    issued as three N = NT MMAs into the same accumulator columns (round 2: measured never slower than the round-1 "wide" form
//    A_hi x [W_hi;W_lo] as one N = 2*NT MMA + A_lo x W_hi, which stays available with BFSR_TC_WIDE=1).  The fast mode issues A_hi x W_hi.
//  * Accumulators live in TMEM (two stages whenever they fit 512 columns); two warps issue the MMAs (even / odd
//    sub-tiles); tcgen05.commit arrives on the mbarriers that recycle the A / W slots and release the epilogue.
//  * Epilogue (8 warps in the TMA variant, two per TMEM lane quarter; 16 in the lean instantiation used by the short-K convs,
//    whose epilogue is bound by instruction latency): tcgen05.ld 32x32b -> bias / pre-activation /
//    activation / residual passes on whole rows -> swizzled staging tile -> TMA bulk tensor store (fp32 or bf16 planes,
//    optionally a second copy); or a FlowStep applied in place of the store (FlowEpi).
//  * Modes: phase 1/2 = conv over nearest-2x-upsampled channels evaluated per output phase with pre-summed 2x2 taps
//    (phase 2 adds the hi-res channels as TMA parity planes in the same pass); n_pre = a BF16X2 pre-activation tensor
//    enters the GEMM as identity K chunks; fold = 1: nine taps folded into N with a shared-memory shift-add epilogue (opt-in, slower);
//    fold = 2: the three dx taps folded into N over a raster of pitch 32, shift-add by warp shuffles (default for Cout <= 32).
//
// Warp roles (Roles<TMA_IN>): epilogue warps, MMA issuer A (+ TMEM allocator), weight loader, TMA loader or 8 A producers,
// MMA issuer B.
#include "ops.cuh"
#include "tc_ptx.cuh"
#include <cuda.h>
#include <vector>
#include <cstring>
#include <cmath>
#include <cstdlib>

namespace bfsr {

thread_local int g_conv_mode = 0;   // 0 = split-bf16 x3 on tcgen05 (accurate), 1 = bf16 (fast), 2 = fp32 CUDA cores only
thread_local int g_tc_fold = -1;    // tap-folded small-Cout convs: -1 = environment default (off), 0 = off, 1 = on

#ifndef BFSR_EPI_WARPS
#define BFSR_EPI_WARPS 8   // epilogue warps of the TMA-fed variant (multiple of 4)
#endif
namespace tc {
constexpr int NA_MAX = 4, NW_MAX = 4;          // A-operand slots are 2..4 per launch (a.na): deep enough to hide the TMA latency of HBM-bound convs
constexpr int PROD_WARPS = 8, NPROD = PROD_WARPS * 32;
// Warp roles.  TMA-fed variant (input already in bf16 hi/lo planes): 8 epilogue warps (two per TMEM lane quarter -- one
// warp per scheduler cannot hide its own ALU latency and made every small-K conv epilogue-bound), MMA issuer A, weight
// loader, TMA loader for A, MMA issuer B = 12 warps.  Register-producer variant (fp32 / upsampled / phase inputs): 4 epilogue
// warps, MMA issuer A, weight loader, 8 producer warps, MMA issuer B = 15 warps.
template <bool TMA_IN, int EW = BFSR_EPI_WARPS> struct Roles {
  static constexpr int EPI_WARPS = TMA_IN ? EW : 4;
  static constexpr int W_MMA = EPI_WARPS, W_WPROD = EPI_WARPS + 1, W_PROD0 = EPI_WARPS + 2;
  static constexpr int N_PROD = TMA_IN ? 1 : PROD_WARPS;
  static constexpr int W_MMAB = W_PROD0 + N_PROD;                    // first of the extra MMA issuing warps
  static constexpr int N_ISS_MAX = 2;                                // issuing warps incl. W_MMA; 4 measured no faster (the pipe's small-N floor binds)
  static constexpr int NTHREADS = (W_MMAB + N_ISS_MAX - 1) * 32;
};
constexpr int MAX_SMEM = 227 * 1024;
constexpr int STG_WARP = 4096;               // epilogue staging per warp: 32 pixel rows x 128 B
constexpr int stg_bytes(int ew) { return ew * STG_WARP; }
constexpr int bias_bytes(int ew) { return ew * 128 * 4; }   // per-epilogue-warp copy of the cout tile's bias
constexpr int FLOW_BYTES = (24 * 24 + 32) * 4;   // channel-mix matrix + offset vector of a fused FlowStep (C <= 24)
constexpr int MAXI = 10;                     // (pixel, 8-channel) items a producer thread prefetches per chunk
}  // namespace tc

struct TcArgs {
  alignas(64) CUtensorMap tmap;          // BF16X2 input: 5-D (C, W, H, N, plane) tiled map, box = one halo tile x 32 channels
  alignas(64) CUtensorMap tmap_pl[4];    // phase 2: hi-res input seen as four parity planes (py, px): base offset + doubled strides
  alignas(64) CUtensorMap tmap_out;      // output(s): box = 32 channels x 8 x 4 pixels (one epilogue warp's rows of a sub-tile)
  alignas(64) CUtensorMap tmap_out2;
  int tma_out, tma_out2;                 // epilogue leaves through TMA bulk tensor stores (else per-lane direct stores)
  View in, out, out2, pre, res1, res2;
  int tma;                               // A operand arrives by TMA (input already split into bf16 hi/lo planes in HBM)
  int in_bf;                             // register producer reads a BF16X2 view (IN_UP2 / phase modes)
  const unsigned char* w; const float* bias;
  int cin, cout, nt, n_chunks, n_ct;     // n_ct = cout tiles
  int H, W, N, in_mode, act;
  float eps, alpha, beta1, beta2;
  int fast;
  int n_iss;                             // MMA issuing warps (1 or 2)
  int na;                                // A-operand slots in the ring (2..4)
  int wide;                              // accurate mode with NT < 64: A_hi x [W_hi;W_lo] as one N = 2NT MMA (column halves summed by the epilogue)
  int ks, ntaps, halo;                   // 3x3 (9 taps, halo 1) or 1x1 (1 tap, halo 0)
  int phase;                             // 1: conv over a nearest-2x-upsampled input evaluated as four 2x2 phase convs
                                         // 2: single pass over [hi-res channels | nearest2x(low-res channels)]: per output phase
                                         //    the low-res chunks take 4 pre-summed taps, the hi-res channels arrive as four
                                         //    parity planes (stride-2 TMA boxes), each serving the 3x3 taps that land on it
  int n_lo, taps_tile;                   // phase 2: low-res chunks per tile; tap images per (cout tile, phase)
  unsigned short pl_off[4][4][4];        // phase 2: [phase][parity plane][i] view offset (16-byte units) of the plane's i-th tap
  unsigned char pl_nt[4][4];             // phase 2: taps served by a parity plane for an output phase (1, 2, 2 or 4)
  int fold;                              // tap folding (3x3, Cin = 64, Cout <= 24): one GEMM over the HALO tile with N = 9*Cout
                                         // columns (u[r, tap*Cout+co] = x[r,:] . W[tap][:, co]) and a shift-add epilogue
                                         // out[p] = sum_tap u[p + off_tap, tap] -- 9x fewer MMAs for convs that sit on the
                                         // small-N MMA floor
                                         // fold = 2 (dx folding, 3x3, Cout <= 32): the tile is a raster of pitch 32 pixels; per filter
                                         // row dy ONE MMA computes u[r, dx*cb+co] = x[r + 32*dy,:] . W[dy][dx][:, co] for the three dx at
                                         // once (N = 3*cb), M tiles = runs of 128 raster positions = 4 image rows, and the epilogue
                                         // forms out[r] = u0[r] + u1[r+1] + u2[r+2] with two warp shuffles per channel (a warp holds one
                                         // image row of 32 positions, 30 of them valid outputs): 3x fewer MMAs of 3x the width -- the
                                         // small-N floor of the MMA (A operand read from shared memory, ~40 clk) is paid once per dy
  int cb;                                // fold = 2: accumulator column stride of a dx block (16 or 32)
  int pdl;                               // launched with programmatic stream serialisation (griddepcontrol in the kernel)
  int dbg;                               // timing experiments only (BFSR_TC_DBG: 1 = no proxy fence, 2 = no TMA store, 4 = no hi/lo split math): WRONG results
  int stg_bytes;                         // epilogue staging bytes in shared memory (0: the epilogue never stages, e.g. dx-folded head + FlowStep)
  int tile_w, tile_h;                    // output pixels of a macro tile (fold 1: 16x16 or 14x14; fold 2: 30 x 4*mt; else 8*sx x 16*sy)
  FlowEpi flow;                          // flow.C != 0: the epilogue applies the FlowStep instead of storing the conv output
  View in2;                              // phase 2: hi-res part of the input (BF16X2); n_pre: the pre-activation tensor
  int n_pre, n_main;                     // pre-activation folded into the GEMM: n_pre extra 32-channel chunks of `in2` with ONE
                                         // centre tap and identity weights follow the n_main chunks of the conv proper
  int mt, sx, sy;                        // sub-tiles per macro tile and their arrangement (sx * sy = mt)
  int pitch, hrows;                      // halo tile: pitch = 8*sx+2 pixels, hrows = 16*sy+2
  int a_plane, a_slot, w_slot;           // bytes (w_slot = one tap image)
  int tps, nw, w_stage;                  // taps per weight stage, number of stages, bytes per stage
  int w_res;                             // all weight stages fit in smem: loaded once per CTA, never recycled (nw = stages per tile)
  int nacc;                              // TMEM accumulator stages (1, 2 or 4)
  int tiles_x, tiles_y, total_tiles;     // macro tiles per image and total work tiles (incl. cout tiles, batch)
  uint32_t m_nct, m_tx, m_ty;            // ceil(2^32 / d) for d = n_ct, tiles_x, tiles_y (tile_coord)
  uint32_t tmem_cols;
};


// 4 consecutive channels of one pixel -> bf16 (hi, lo) planes of a BF16X2 view (the operand format of the next conv)
__device__ __forceinline__ void store_split(const View& v, long long pix, int c, const float4& o) {
  const float h0 = __bfloat162float(__float2bfloat16_rn(o.x)), h1 = __bfloat162float(__float2bfloat16_rn(o.y));
  const float h2 = __bfloat162float(__float2bfloat16_rn(o.z)), h3 = __bfloat162float(__float2bfloat16_rn(o.w));
  __nv_bfloat16* d = (__nv_bfloat16*)v.p + pix * v.cs + v.coff + c;
  *reinterpret_cast<uint2*>(d) = make_uint2(pack_bf16(h0, h1), pack_bf16(h2, h3));
  *reinterpret_cast<uint2*>(d + v.plane) = make_uint2(pack_bf16(o.x - h0, o.y - h1), pack_bf16(o.z - h2, o.w - h3));
}

// FlowStep on one pixel (one lane): h = accumulator row after bias + cross-sigmoid; Ms / cs = mix matrix and vector in smem
// operands of the step that do not depend on the conv: fetched BEFORE the wait for the accumulator, so the global-load latency hides
// behind the MMAs (with two to four epilogue warps per scheduler it was fully exposed: 63 % long-scoreboard stalls in this code)
template <int C>
__device__ __forceinline__ void flow_prefetch(const FlowEpi& f, long long pix, float4* zq, float4* hq) {
  const float4* zp = reinterpret_cast<const float4*>((const float*)f.z_in.p + pix * f.z_in.cs + f.z_in.coff);
#pragma unroll
  for (int k = 0; k < C / 4; ++k) zq[k] = __ldg(zp + k);
  if (f.hF.p) {
    const float4* fp = reinterpret_cast<const float4*>((const float*)f.hF.p + pix * f.hF.cs + f.hF.coff);
#pragma unroll
    for (int k = 0; k < C / 2; ++k) hq[k] = __ldg(fp + k);
  }
}
// divisions by the coupling scales: sigmoid(.) + 1e-4 lies in [1e-4, 1.0001], far inside the range where the fast division is
// accurate to 2 ulp (the reference divides in fp32 too: FlowAffineCouplingsAblation.py:85,92)
template <int C>
__device__ __forceinline__ void flow_epilogue(const FlowEpi& f, const float* h, const float* Ms, const float* cs, long long pix,
                                              const float4* zq = nullptr, const float4* hq = nullptr) {
  float z[C], o[C];
  const float4* zp = reinterpret_cast<const float4*>((const float*)f.z_in.p + pix * f.z_in.cs + f.z_in.coff);
#pragma unroll
  for (int k = 0; k < C / 4; ++k) { const float4 v = zq ? zq[k] : __ldg(zp + k); z[4 * k] = v.x; z[4 * k + 1] = v.y; z[4 * k + 2] = v.z; z[4 * k + 3] = v.w; }
#pragma unroll
  for (int j = 0; j < C / 2; ++j) {
    if (f.inv) z[C / 2 + j] = __fdividef(z[C / 2 + j], h[2 * j + 1]) - h[2 * j];
    else z[C / 2 + j] = (z[C / 2 + j] + h[2 * j]) * h[2 * j + 1];
  }
  const float4* fp = reinterpret_cast<const float4*>((const float*)f.hF.p + pix * f.hF.cs + f.hF.coff);
  if (f.inv && f.hF.p) {
#pragma unroll
    for (int k = 0; k < C / 2; ++k) {
      const float4 v = hq ? hq[k] : __ldg(fp + k);
      z[2 * k] = __fdividef(z[2 * k], v.y) - v.x; z[2 * k + 1] = __fdividef(z[2 * k + 1], v.w) - v.z;
    }
  }
  if (f.has_mix) {
#pragma unroll
    for (int co = 0; co < C; ++co) {
      float acc = 0.f;
#pragma unroll
      for (int k = 0; k < C / 4; ++k) {
        const float4 m = *reinterpret_cast<const float4*>(Ms + co * C + 4 * k);
        acc = fmaf(m.x, z[4 * k], acc); acc = fmaf(m.y, z[4 * k + 1], acc);
        acc = fmaf(m.z, z[4 * k + 2], acc); acc = fmaf(m.w, z[4 * k + 3], acc);
      }
      o[co] = f.inv ? acc - cs[co] : acc + cs[co];
    }
    if (!f.inv && f.hF.p) {
#pragma unroll
      for (int k = 0; k < C / 2; ++k) { const float4 v = hq ? hq[k] : __ldg(fp + k); o[2 * k] = (o[2 * k] + v.x) * v.y; o[2 * k + 1] = (o[2 * k + 1] + v.z) * v.w; }
    }
  } else {
#pragma unroll
    for (int c = 0; c < C; ++c) o[c] = z[c];
  }
  float4* dst = reinterpret_cast<float4*>((float*)f.z_out.p + pix * f.z_out.cs + f.z_out.coff);
#pragma unroll
  for (int k = 0; k < C / 4; ++k) dst[k] = make_float4(o[4 * k], o[4 * k + 1], o[4 * k + 2], o[4 * k + 3]);
  if (f.z1op.p) {
    constexpr int CP = (C / 2 + 7) & ~7;
    __nv_bfloat16* d = (__nv_bfloat16*)f.z1op.p + pix * f.z1op.cs + f.z1op.coff;
#pragma unroll
    for (int k = 0; k < CP / 8; ++k) {
      uint32_t hi[4], lo[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int c0 = 8 * k + 2 * e;
        const float x0 = c0 < C / 2 ? o[c0] : 0.f, x1 = c0 + 1 < C / 2 ? o[c0 + 1] : 0.f;
        const float h0 = __bfloat162float(__float2bfloat16_rn(x0)), h1 = __bfloat162float(__float2bfloat16_rn(x1));
        hi[e] = pack_bf16(h0, h1); lo[e] = pack_bf16(x0 - h0, x1 - h1);
      }
      *reinterpret_cast<uint4*>(d + 8 * k) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
      *reinterpret_cast<uint4*>(d + f.z1op.plane + 8 * k) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    }
  }
}


#ifdef BFSR_TC_TRACE
#define TR_DECL(...) long long __VA_ARGS__
#define TR_T(x) const long long x = clock64()
#define TR_ADD(acc, t0) acc += clock64() - (t0)
#else
#define TR_DECL(...)
#define TR_T(x)
#define TR_ADD(acc, t0)
#endif

struct TileCoord { int n, ty0, tx0, ct, ph; };
// t / d for the small divisors of the tile walk with a host-computed reciprocal m = ceil(2^32 / d): exact while t * d < 2^32
// (tile counts stay below 2^20, divisors below 2^12); every warp of the CTA decodes every tile, so the three runtime divisions
// were ~90 instructions per warp per tile
__device__ __forceinline__ uint32_t fdiv(uint32_t t, uint32_t m, uint32_t d) { return d == 1 ? t : __umulhi(t, m); }
__device__ __forceinline__ TileCoord tile_coord(const TcArgs& a, int t) {
  TileCoord c;
  uint32_t u = (uint32_t)t, qd;
  qd = fdiv(u, a.m_nct, (uint32_t)a.n_ct); c.ct = (int)(u - qd * a.n_ct); u = qd;
  c.ph = 0;
  if (a.phase) { c.ph = u & 3; u >>= 2; }
  qd = fdiv(u, a.m_tx, (uint32_t)a.tiles_x); const int tx = (int)(u - qd * a.tiles_x); u = qd;
  qd = fdiv(u, a.m_ty, (uint32_t)a.tiles_y); const int ty = (int)(u - qd * a.tiles_y); c.n = (int)qd;
  c.ty0 = ty * a.tile_h; c.tx0 = tx * a.tile_w;
  return c;
}

// CL = 2: launched as 2-CTA clusters.  The two CTAs walk tile PAIRS that share (cout tile, phase) -- hence the whole weight
// stream -- on neighbouring spatial tiles: each CTA fetches half of every weight stage and multicasts it into both shared
// memories, and a stage is recycled when the MMAs of BOTH CTAs have retired.  Halves the L2 -> SM weight traffic of the
// convs whose 0.8 MB-per-tile weight stream binds them (level-1 phase convs).  Everything else is per CTA as for CL = 1.
// EW = epilogue warps of the TMA-fed variant (8 or 16).  The convs with little K per output (the z-dependent coupling convs, the 1x1
// convs, the dx-folded heads) are bound by their epilogue's instruction LATENCY -- ncu (profiles/r2_coupling_l1_summary.md): 19 % of
// the warp slots occupied, 'wait' / 'short scoreboard' stalls, tensor pipe 11-22 %, DRAM 30-59 % -- so they run with four warps per
// TMEM lane quarter instead of two.
template <bool TMA_IN, int CL = 1, int EW = BFSR_EPI_WARPS, bool LEANP = (EW == 16)>
__global__ void __launch_bounds__(tc::Roles<TMA_IN, EW>::NTHREADS, 1) conv_tc_kernel(const __grid_constant__ TcArgs a) {
  using namespace tc;
  using R = Roles<TMA_IN, EW>;
  constexpr int BIAS_BYTES = bias_bytes(TMA_IN ? EW : BFSR_EPI_WARPS);
  // the 16-warp instantiation has 96 registers per thread: it carries only the epilogues of the convs that use it (no tap folding,
  // no C = 24 FlowStep, no residual / fp32 pre-activation / second-output passes; launch_tc checks the same conditions)
  constexpr bool LEAN = LEANP;
  static_assert(CL == 1 || (CL == 2 && TMA_IN), "clusters: TMA-fed variant only");
  // persistent tile walk: CL = 1 strides the tile list by the grid; CL = 2 strides the PAIR list by the cluster count
  // (macros, not hoisted constants: the CL = 1 loops stay on the uniform datapath exactly as before)
#define TC_T_FIRST (CL == 2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x)
#define TC_T_STEP (CL == 2 ? (int)(gridDim.x >> 1) : (int)gridDim.x)
#define TC_T_END (CL == 2 ? a.total_tiles >> 1 : a.total_tiles)
  const int cl_rank = CL == 2 ? (int)(blockIdx.x & 1u) : 0;
  auto tile_of = [&](int p) -> int {
    if (CL == 1) return p;
    const int g = a.n_ct * (a.phase ? 4 : 1);        // tiles that differ only in (cout tile, phase) are consecutive
    const int sp = p / g;
    return (sp * 2 + cl_rank) * g + (p - sp * g);
  };
  // Programmatic dependent launch: the next kernel of the stream may be scheduled as soon as CTAs of this one exit, and runs its
  // prologue (barrier init, TMEM allocation, FlowStep matrices) while our last tiles drain; it blocks at griddepcontrol.wait below
  // until this whole grid has completed and its writes are visible.  Hides ~5-10 us of launch + prologue per conv (860 per step).
  if (a.pdl) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  extern __shared__ unsigned char smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_smem = base;                                   // NA slots of [hi plane | lo plane]
  const uint32_t w_smem = base + a.na * a.a_slot;                 // NW slots
  const uint32_t stg_smem = w_smem + a.nw * a.w_stage;            // epilogue staging (1024-aligned: every slot size is)
  const uint32_t bars = stg_smem + (uint32_t)a.stg_bytes;         // mbarriers (8 B each), then the per-warp bias copies
  const uint32_t a_full = bars, a_empty = bars + 8 * NA_MAX, w_full = bars + 16 * NA_MAX, w_empty = w_full + 8 * NW_MAX;
  const uint32_t acc_full = w_empty + 8 * NW_MAX, acc_empty = acc_full + 32, tmem_slot = acc_empty + 32;   // up to four accumulator stages
  unsigned char* smem_gen = smem_raw + (base - smem_u32(smem_raw));

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);        // provably warp-uniform role index
  const int nt = a.nt;
  const int acc_cols = a.mt * (a.wide ? 2 * nt : nt);            // TMEM columns per accumulator stage

  if (tid == 0) {
    for (int i = 0; i < NA_MAX; ++i) { mbar_init(a_full + 8 * i, TMA_IN ? 1 : NPROD); mbar_init(a_empty + 8 * i, a.n_iss); }
    for (int i = 0; i < NW_MAX; ++i) { mbar_init(w_full + 8 * i, 1); mbar_init(w_empty + 8 * i, a.n_iss * CL); }
    for (int i = 0; i < 4; ++i) { mbar_init(acc_full + 8 * i, a.n_iss); mbar_init(acc_empty + 8 * i, 32 * R::EPI_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  float* flow_s = reinterpret_cast<float*>(smem_gen + (bars + 256 + BIAS_BYTES - base));   // [C*C] mix matrix, then [C] vector
  if (a.flow.C && a.flow.has_mix) {
    for (int i = tid; i < a.flow.C * a.flow.C; i += blockDim.x) flow_s[i] = a.flow.M[i];
    for (int i = tid; i < a.flow.C; i += blockDim.x) flow_s[24 * 24 + i] = a.flow.cvec[i];
  }
  if (warp == R::W_MMA) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(a.tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (CL == 2) cluster_sync_all();       // the peer's barriers exist before any multicast copy / remote commit targets them
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  tmem_base = __shfl_sync(0xffffffffu, tmem_base, 0);
  if (a.pdl) asm volatile("griddepcontrol.wait;" ::: "memory");   // everything below touches activations of the previous kernel

  if (TMA_IN && warp == R::W_PROD0) {
    // ===================== A by TMA: the input already lives in HBM as bf16 (hi, lo) planes =====================
    // One tiled TMA load per plane brings the whole halo tile of a 32-channel chunk straight into the UMMA SWIZZLE_64B
    // layout (one pixel = one 64-byte row); out-of-image pixels and channels past the view are zero-filled by the TMA
    // unit, which is exactly the conv's zero padding.
    {
      const uint32_t box_bytes = (uint32_t)(a.pitch * a.hrows) * ROWB;
      int a_it = 0;
      for (int p = TC_T_FIRST; p < TC_T_END; p += TC_T_STEP) {
        const TileCoord tcd = tile_coord(a, tile_of(p));
        for (int c = 0; c < a.n_chunks; ++c, ++a_it) {
          const int slot = a_it % a.na;
          mbar_wait_relaxed(a_empty + 8 * slot, ((a_it / a.na) & 1) ^ 1);
          if (elect_one()) {
            const uint32_t dst = a_smem + slot * a.a_slot, bar = a_full + 8 * slot;
            mbar_expect_tx(bar, a.fast ? box_bytes : 2 * box_bytes);
            if (a.n_pre && c >= a.n_main) {     // pre-activation chunk: same halo tile geometry, read through the second map
              const int c0 = a.in2.coff + (c - a.n_main) * KC;
              tma_load_5d(dst, &a.tmap_pl[0], bar, c0, tcd.tx0 - a.halo, tcd.ty0 - a.halo, tcd.n, 0);
              if (!a.fast) tma_load_5d(dst + a.a_plane, &a.tmap_pl[0], bar, c0, tcd.tx0 - a.halo, tcd.ty0 - a.halo, tcd.n, 1);
            } else if (a.phase == 2 && c >= a.n_lo) {   // parity plane (py, px) of a hi-res channel chunk: every second pixel
              const int pc = c - a.n_lo, pl = pc & 3, cc = pc >> 2;
              tma_load_5d(dst, &a.tmap_pl[pl], bar, a.in2.coff + cc * KC, tcd.tx0 - 1, tcd.ty0 - 1, tcd.n, 0);
              if (!a.fast) tma_load_5d(dst + a.a_plane, &a.tmap_pl[pl], bar, a.in2.coff + cc * KC, tcd.tx0 - 1, tcd.ty0 - 1, tcd.n, 1);
            } else {
              tma_load_5d(dst, &a.tmap, bar, a.in.coff + c * KC, tcd.tx0 - a.halo, tcd.ty0 - a.halo, tcd.n, 0);
              if (!a.fast) tma_load_5d(dst + a.a_plane, &a.tmap, bar, a.in.coff + c * KC, tcd.tx0 - a.halo, tcd.ty0 - a.halo, tcd.n, 1);
            }
          }
          __syncwarp();
        }
      }
    }
  } else if (!TMA_IN && warp >= R::W_PROD0 && warp < R::W_MMAB) {
    // ===================== A producers: fp32 halo tile -> (hi, lo) bf16 planes, swizzled =====================
    // Each thread owns up to MAXI (pixel, 8-channel) items of a chunk.  The global loads of chunk i+1 are issued
    // into registers BEFORE waiting for its smem slot, so their latency hides behind the MMAs of chunk i-1.
    const int ptid = tid - R::W_PROD0 * 32;
    const int inH = a.in_mode == IN_UP2 ? a.H >> 1 : a.H, inW = a.in_mode == IN_UP2 ? a.W >> 1 : a.W;
    const int items = a.hrows * a.pitch * 4;           // (pixel, 8-channel group) pairs per chunk
    float4 v[MAXI][2];
    auto load_chunk = [&](const TileCoord& tcd, int c) {
      const long long img = (long long)tcd.n * inH * inW;
#pragma unroll
      for (int u = 0; u < MAXI; ++u) {
        const int it = ptid + u * NPROD;
        v[u][0] = make_float4(0.f, 0.f, 0.f, 0.f); v[u][1] = v[u][0];
        if (it < items) {
          const int q = it >> 2, j = it & 3;
          const int yy = q / a.pitch, xx = q - yy * a.pitch;
          const int gy = tcd.ty0 + yy - a.halo, gx = tcd.tx0 + xx - a.halo;
          const int cb = c * KC + j * 8;
          if (gy >= 0 && gy < a.H && gx >= 0 && gx < a.W && cb < a.cin) {
            const int sy_ = a.in_mode == IN_UP2 ? gy >> 1 : gy, sx_ = a.in_mode == IN_UP2 ? gx >> 1 : gx;
            const long long p = img + (long long)sy_ * inW + sx_;
            if (a.in_bf) {   // already split: 8 channels = 16 bytes of the hi plane and 16 of the lo plane
              const __nv_bfloat16* src = (const __nv_bfloat16*)a.in.p + p * a.in.cs + a.in.coff + cb;
              v[u][0] = __ldg(reinterpret_cast<const float4*>(src));
              v[u][1] = __ldg(reinterpret_cast<const float4*>(src + a.in.plane));
            } else if (cb + 8 <= a.cin) {   // eligibility guarantees 128-bit addressable views
              const float4* src = reinterpret_cast<const float4*>((const float*)a.in.p + p * a.in.cs + a.in.coff + cb);
              v[u][0] = __ldg(src); v[u][1] = __ldg(src + 1);
            } else {          // ragged last group (Cin % 8 != 0, e.g. the 6-channel z1 of the first coupling level)
              const float* src = (const float*)a.in.p + p * a.in.cs + a.in.coff + cb;
              float x[8];
              if (cb + 4 <= a.cin) {
                const float4 t4 = __ldg(reinterpret_cast<const float4*>(src));
                x[0] = t4.x; x[1] = t4.y; x[2] = t4.z; x[3] = t4.w;
              } else {
#pragma unroll
                for (int e = 0; e < 4; ++e) x[e] = cb + e < a.cin ? __ldg(src + e) : 0.f;
              }
#pragma unroll
              for (int e = 4; e < 8; ++e) x[e] = cb + e < a.cin ? __ldg(src + e) : 0.f;
              v[u][0] = make_float4(x[0], x[1], x[2], x[3]); v[u][1] = make_float4(x[4], x[5], x[6], x[7]);
            }
          }
        }
      }
    };
    auto store_chunk = [&](unsigned char* pl_hi) {
#pragma unroll
      for (int u = 0; u < MAXI; ++u) {
        const int it = ptid + u * NPROD;
        if (it < items) {
          const int q = it >> 2, j = it & 3;
          const uint32_t off = (uint32_t)q * ROWB + (uint32_t)((j ^ ((q >> 1) & 3)) << 4);
          if (a.in_bf) {
            *reinterpret_cast<float4*>(pl_hi + off) = v[u][0];
            if (!a.fast) *reinterpret_cast<float4*>(pl_hi + a.a_plane + off) = v[u][1];
            continue;
          }
          const float x[8] = {v[u][0].x, v[u][0].y, v[u][0].z, v[u][0].w, v[u][1].x, v[u][1].y, v[u][1].z, v[u][1].w};
          float hi[8], lo[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) { hi[e] = __bfloat162float(__float2bfloat16_rn(x[e])); lo[e] = x[e] - hi[e]; }
          *reinterpret_cast<uint4*>(pl_hi + off) =
              make_uint4(pack_bf16(hi[0], hi[1]), pack_bf16(hi[2], hi[3]), pack_bf16(hi[4], hi[5]), pack_bf16(hi[6], hi[7]));
          if (!a.fast)
            *reinterpret_cast<uint4*>(pl_hi + a.a_plane + off) =
                make_uint4(pack_bf16(lo[0], lo[1]), pack_bf16(lo[2], lo[3]), pack_bf16(lo[4], lo[5]), pack_bf16(lo[6], lo[7]));
        }
      }
    };
    int a_it = 0, t = blockIdx.x, c = 0;
    bool have = t < a.total_tiles;
    TileCoord tcd = tile_coord(a, have ? t : 0);
    TR_DECL(tr_wait = 0, tr_fill = 0); TR_T(tr_start);
    if (have) load_chunk(tcd, 0);
    while (have) {
      const int slot = a_it % a.na;
      TR_T(tr0);
      mbar_wait_relaxed(a_empty + 8 * slot, ((a_it / a.na) & 1) ^ 1);
      TR_ADD(tr_wait, tr0); TR_T(tr1);
      store_chunk(smem_gen + slot * a.a_slot);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the tensor core
      mbar_arrive(a_full + 8 * slot);
      ++a_it;
      if (++c == a.n_chunks) { c = 0; t += gridDim.x; have = t < a.total_tiles; if (have) tcd = tile_coord(a, t); }
      if (have) load_chunk(tcd, c);
      TR_ADD(tr_fill, tr1);
    }
#ifdef BFSR_TC_TRACE
    if (blockIdx.x == 0 && ptid == 0) printf("[tc trace] producer: total %lld wait_a_empty %lld fill %lld (chunks %d)\n", clock64() - tr_start, tr_wait, tr_fill, a_it);
#endif
  } else if (warp < R::EPI_WARPS) {
    const int q = warp & 3, grp = warp >> 2;             // TMEM lane quarter of this warp; block-interleave group
    constexpr int N_GRP = R::EPI_WARPS / 4;
    // ===================== epilogue: TMEM -> registers -> fused epilogue -> swizzled smem -> TMA bulk tensor store ==========
    // tcgen05.ld 32x32b hands every lane one accumulator ROW = the channels of one output pixel.  The lane applies bias /
    // pre-activation / activation / residuals to its pixel, writes the 32-channel piece into a per-warp staging tile in
    // the TMA swizzle pattern (conflict-free) and one lane issues a bulk tensor store of the warp's 8 x 4 pixel box: the
    // TMA unit writes whole lines asynchronously and clips ragged edges / channel tails, the LSU only sees the (few)
    // operand loads.  Views that TMA cannot address (phase outputs, odd strides) fall back to per-lane 16-byte stores.
    const int sub_cols = a.wide ? 2 * nt : nt;
    const float slope = a.act == ACT_LRELU ? 0.2f : (a.act == ACT_RELU ? 0.f : 1.f);
    unsigned char* stg_gen = smem_gen + (stg_smem - base) + warp * STG_WARP;
    const uint32_t stg_u32 = stg_smem + warp * STG_WARP;
    float* bias_s = reinterpret_cast<float*>(smem_gen + (bars + 256 - base)) + warp * 128;
    int cur_ct = -1, t_it = 0;
    // stage one 32-pixel x 32-channel block of this warp and launch its bulk store (out and out2 share the two buffers)
    TR_DECL(tr_stw = 0, tr_stg = 0, tr_ld = 0, tr_math = 0);
    auto stage_store = [&](const View& v, const CUtensorMap* tm, const float* o, int ncol, int c0, int x0, int y0, int n) {
      TR_T(trs0);
      if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // the previous store has drained the buffer
      __syncwarp();
      TR_ADD(tr_stw, trs0); TR_T(trs1);
      unsigned char* buf = stg_gen;
      if (v.fmt == F32) {
#pragma unroll
        for (int k = 0; k < 8; ++k)
          if (4 * k < ncol)
            *reinterpret_cast<float4*>(buf + lane * 128 + ((k ^ (lane & 7)) << 4)) = make_float4(o[4 * k], o[4 * k + 1], o[4 * k + 2], o[4 * k + 3]);
      } else {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (8 * k < ncol) {
            uint32_t hi[4], lo[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {   // packed conversion; bf16 -> fp32 of the hi part is a shift / mask of the packed word
              const float x0f = o[8 * k + 2 * e], x1f = o[8 * k + 2 * e + 1];
              hi[e] = pack_bf16(x0f, x1f);
              lo[e] = pack_bf16(x0f - __uint_as_float(hi[e] << 16), x1f - __uint_as_float(hi[e] & 0xffff0000u));
            }
            const uint32_t off = lane * 64 + ((k ^ ((lane >> 1) & 3)) << 4);
            *reinterpret_cast<uint4*>(buf + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
            *reinterpret_cast<uint4*>(buf + 2048 + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
          }
      }
      if (!(a.dbg & 1)) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0 && !(a.dbg & 2)) {
        const uint32_t src = stg_u32;
        if (v.fmt == F32) tma_store_4d(tm, src, v.coff + c0, x0, y0, n);
        else { tma_store_5d(tm, src, v.coff + c0, x0, y0, n, 0); tma_store_5d(tm, src + 2048, v.coff + c0, x0, y0, n, 1); }
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
      TR_ADD(tr_stg, trs1);
    };
    auto direct_store = [&](const View& v, long long p, int co, const float4& o) {
      if (v.fmt == F32) *reinterpret_cast<float4*>((float*)v.p + p * v.cs + v.coff + co) = o;
      else store_split(v, p, co, o);
    };
    TR_DECL(tr_wait = 0, tr_work = 0); TR_T(tr_start);
    for (int p = TC_T_FIRST; p < TC_T_END; p += TC_T_STEP, ++t_it) {
      const TileCoord tcd = tile_coord(a, tile_of(p));
      const int as = t_it % a.nacc;
      const int co_base = tcd.ct * nt;
      if (tcd.ct != cur_ct) {   // bias of this cout tile -> the warp's smem copy, fetched before the accumulator wait
        cur_ct = tcd.ct;
        __syncwarp();
        float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (lane * 4 < nt && co_base + lane * 4 < a.cout) b4 = __ldg(reinterpret_cast<const float4*>(a.bias + co_base + lane * 4));
        *reinterpret_cast<float4*>(bias_s + lane * 4) = b4;
        __syncwarp();
      }
      // dx-folded head with a fused FlowStep: this warp's (single) M tile of the work tile is known now -- fetch z / hF early
      float4 zq[LEAN ? 3 : 6], hq[LEAN ? 6 : 12];
      const bool pf = TMA_IN && a.fold == 2 && a.flow.C && a.mt <= N_GRP && grp < a.mt;
      if (pf) {
        const int gy = tcd.ty0 + 4 * grp + q, gx = tcd.tx0 + lane;
        if (lane < 30 && gy < a.H && gx < a.W) {
          const long long p = ((long long)tcd.n * a.H + gy) * a.W + gx;
          if (LEAN || a.flow.C == 12) flow_prefetch<12>(a.flow, p, zq, hq); else flow_prefetch<LEAN ? 12 : 24>(a.flow, p, zq, hq);
        }
      }
      TR_T(tr0);
      mbar_wait_relaxed(acc_full + 8 * as, (t_it / a.nacc) & 1);
      TR_ADD(tr_wait, tr0); TR_T(tr1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (TMA_IN && !LEAN && a.fold == 1) {
        // ---- tap-folded conv: TMEM holds u[halo row][tap*Cout + co]; shift-add the nine taps through a shared-memory
        // accumulator (fixed tap order, block barriers between taps: deterministic), then bias / activation / FlowStep
        const int Cc = a.cout, P = a.pitch, nrows = a.pitch * a.hrows;
        float* oacc = reinterpret_cast<float*>(smem_gen + (stg_smem - base));     // [tile_h*tile_w][Cc], reuses the staging area
        const int n_out = a.tile_h * a.tile_w;
        for (int i = tid; i < n_out * Cc; i += 32 * R::EPI_WARPS) oacc[i] = 0.f;
        asm volatile("bar.sync 1, %0;" ::"n"(32 * R::EPI_WARPS) : "memory");
        for (int tp = 0; tp < 9; ++tp) {
          const int dy = tp / 3, dx = tp - 3 * dy;
          for (int m = grp; m < a.mt; m += N_GRP) {
            const int r = m * 128 + q * 32 + lane;           // halo-tile raster position of this accumulator row
            const int ry = r / P, rx = r - ry * P;
            const int oy = ry - dy, ox = rx - dx;
            float u[32];
            const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + as * acc_cols + m * nt + tp * Cc;
            tmem_ld16(t_row, u);
            if (Cc > 16) tmem_ld16(t_row + 16, u + 16);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (r < nrows && oy >= 0 && oy < a.tile_h && ox >= 0 && ox < a.tile_w) {
              float4* dst = reinterpret_cast<float4*>(oacc + (oy * a.tile_w + ox) * Cc);
#pragma unroll
              for (int k = 0; k < 6; ++k)
                if (4 * k < Cc) { float4 v = dst[k]; v.x += u[4 * k]; v.y += u[4 * k + 1]; v.z += u[4 * k + 2]; v.w += u[4 * k + 3]; dst[k] = v; }
            }
          }
          asm volatile("bar.sync 1, %0;" ::"n"(32 * R::EPI_WARPS) : "memory");
        }
        // the accumulators are consumed: the MMA warps may start the next tile while the outputs are finished
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        mbar_arrive(acc_empty + 8 * as);
        for (int pix = tid; pix < n_out; pix += 32 * R::EPI_WARPS) {
          const int oy = pix / a.tile_w, ox = pix - oy * a.tile_w;
          const int gy = tcd.ty0 + oy, gx = tcd.tx0 + ox;
          if (gy >= a.H || gx >= a.W) continue;
          const long long p = ((long long)tcd.n * a.H + gy) * a.W + gx;
          float hv[24];
#pragma unroll
          for (int k = 0; k < 6; ++k)
            if (4 * k < Cc) {
              const float4 v = *reinterpret_cast<const float4*>(oacc + pix * Cc + 4 * k), b4 = *reinterpret_cast<const float4*>(bias_s + 4 * k);
              hv[4 * k] = v.x + b4.x; hv[4 * k + 1] = v.y + b4.y; hv[4 * k + 2] = v.z + b4.z; hv[4 * k + 3] = v.w + b4.w;
            }
          if (a.act == ACT_CROSS_SIGMOID) {
#pragma unroll
            for (int i = 1; i < 24; i += 2) if (i < Cc) hv[i] = 1.f / (1.f + expf(-(hv[i] + 2.f))) + a.eps;
          }
          if (a.flow.C == 12) flow_epilogue<12>(a.flow, hv, flow_s, flow_s + 24 * 24, p);
          else if (a.flow.C == 24) flow_epilogue<24>(a.flow, hv, flow_s, flow_s + 24 * 24, p);
          else {
            float4* dst = reinterpret_cast<float4*>((float*)a.out.p + p * a.out.cs + a.out.coff);
#pragma unroll
            for (int k = 0; k < 6; ++k) if (4 * k < Cc) dst[k] = make_float4(hv[4 * k], hv[4 * k + 1], hv[4 * k + 2], hv[4 * k + 3]);
          }
        }
        asm volatile("bar.sync 1, %0;" ::"n"(32 * R::EPI_WARPS) : "memory");   // oacc is reused by the next tile
        TR_ADD(tr_work, tr1);
        continue;
      }
      if (TMA_IN && a.fold == 2) {
        // ---- dx-folded 3x3: the lane holds raster position (row 4m+q, column lane) of the tile; its accumulator row carries the
        // three dx partial sums; positions lane+1 / lane+2 of the same image row are the next lanes of this warp
        const int CB = a.cb;
        for (int m = grp; m < a.mt; m += N_GRP) {
          const int gy = tcd.ty0 + 4 * m + q, gx = tcd.tx0 + lane;
          const bool valid = lane < 30 && gy < a.H && gx < a.W;
          const long long p = ((long long)tcd.n * a.H + gy) * a.W + gx;
          const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + as * acc_cols + m * sub_cols;
          float acc[32];
#pragma unroll
          for (int g = 0; g < 2; ++g) {
            if (g * 16 < CB) {
              float u0[16], u1[16], u2[16];
              tmem_ld16(t_row + g * 16, u0); tmem_ld16(t_row + CB + g * 16, u1); tmem_ld16(t_row + 2 * CB + g * 16, u2);
              if (a.wide) {
                float v0[16], v1[16], v2[16];
                tmem_ld16(t_row + nt + g * 16, v0); tmem_ld16(t_row + nt + CB + g * 16, v1); tmem_ld16(t_row + nt + 2 * CB + g * 16, v2);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int i = 0; i < 16; ++i) { u0[i] += v0[i]; u1[i] += v1[i]; u2[i] += v2[i]; }
              } else {
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
              }
#pragma unroll
              for (int i = 0; i < 16; ++i)
                acc[g * 16 + i] = u0[i] + __shfl_down_sync(0xffffffffu, u1[i], 1) + __shfl_down_sync(0xffffffffu, u2[i], 2) + bias_s[g * 16 + i];
            } else {
#pragma unroll
              for (int i = 0; i < 16; ++i) acc[g * 16 + i] = 0.f;
            }
          }
          if (a.act == ACT_CROSS_SIGMOID) {
#pragma unroll
            for (int i = 1; i < 32; i += 2) acc[i] = __fdividef(1.f, 1.f + __expf(-(acc[i] + 2.f))) + a.eps;
          } else if (a.act == ACT_RELU) {
#pragma unroll
            for (int i = 0; i < 32; ++i) acc[i] = fmaxf(acc[i], 0.f);
          } else if (a.act != ACT_NONE) {
#pragma unroll
            for (int i = 0; i < 32; ++i) acc[i] = fmaxf(acc[i], 0.f) + slope * fminf(acc[i], 0.f);
          }
          if (a.flow.C) {
            if (valid) {
              if (LEAN || a.flow.C == 12) flow_epilogue<12>(a.flow, acc, flow_s, flow_s + 24 * 24, p, pf ? zq : nullptr, pf ? hq : nullptr);
              else flow_epilogue<LEAN ? 12 : 24>(a.flow, acc, flow_s, flow_s + 24 * 24, p, pf ? zq : nullptr, pf ? hq : nullptr);
            }
          } else if (a.tma_out) {
            if (gy < a.H) stage_store(a.out, &a.tmap_out, acc, (a.cout + 7) & ~7, 0, tcd.tx0, gy, tcd.n);   // box = 32 channels x 30 pixels x 1 row
          } else if (valid) {
#pragma unroll
            for (int k = 0; k < 8; ++k)
              if (4 * k < a.cout) direct_store(a.out, p, 4 * k, make_float4(acc[4 * k], acc[4 * k + 1], acc[4 * k + 2], acc[4 * k + 3]));
          }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        mbar_arrive(acc_empty + 8 * as);
        TR_ADD(tr_work, tr1);
        continue;
      }
      // 32-channel blocks of this cout tile that hold real channels, and the flattened (sub-tile, block) sequence
      const int nblk = min((nt + 31) >> 5, (a.cout - co_base + 31) >> 5);
      const int nb_total = a.mt * nblk;
      int sub = 0, blk = grp;                            // (sub-tile, block) of item b, advanced without divisions
      for (int b = grp; b < nb_total; b += N_GRP, blk += N_GRP) {      // the warps sharing a lane quarter take alternate blocks
        while (blk >= nblk) { blk -= nblk; ++sub; }
        const int n0 = blk << 5;
        {
        const int idx = q * 32 + lane;                     // accumulator row = pixel of the 8 x 16 sub-tile
        const int sub_y = a.sx == 2 ? sub >> 1 : sub, sub_x = a.sx == 2 ? sub & 1 : 0;
        const int sy0 = tcd.ty0 + sub_y * 16, sx0 = tcd.tx0 + sub_x * 8;
        const int gy = sy0 + (idx >> 3), gx = sx0 + (idx & 7);
        const bool valid = gy < a.H && gx < a.W;
        const long long p = a.phase ? ((long long)tcd.n * 2 * a.H + 2 * gy + (tcd.ph >> 1)) * (2 * a.W) + 2 * gx + (tcd.ph & 1)
                                    : ((long long)tcd.n * a.H + gy) * a.W + gx;
        const long long pl = valid ? p : 0;                // rows past the image edge read pixel 0 (their results are clipped)
        const float* prep = (const float*)a.pre.p + pl * a.pre.cs + a.pre.coff;
        const float* r1p = (const float*)a.res1.p + pl * a.res1.cs + a.res1.coff;
        const float* r2p = (const float*)a.res2.p + pl * a.res2.cs + a.res2.coff;
        const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + as * acc_cols + sub * sub_cols;
        {
          const int ncol = nt - n0 < 32 ? 16 : 32;
          float acc[32];
          TR_T(trl0);
          tmem_ld16(t_row + n0, acc);
          if (ncol == 32) tmem_ld16(t_row + n0 + 16, acc + 16);
          if (a.wide) {
            float acc2[32];
            tmem_ld16(t_row + nt + n0, acc2);
            if (ncol == 32) tmem_ld16(t_row + nt + n0 + 16, acc2 + 16);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int i = 0; i < 32; ++i) acc[i] += acc2[i];
          } else {
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          }
          TR_ADD(tr_ld, trl0); TR_T(trm0);
          // The fused epilogue runs as a few whole-row passes with warp-uniform branches around them (a per-group
          // formulation costs ~2000 issue slots per 32x32 block and made every small-K conv epilogue-bound).
          const int ng = ncol >> 2;                        // 4-channel groups in this block; Cout % 4 == 0 (eligibility)
          const int co0 = co_base + n0;
#pragma unroll
          for (int k = 0; k < 8; ++k)
            if (k < ng) {
              const float4 b4 = *reinterpret_cast<const float4*>(bias_s + n0 + 4 * k);
              acc[4 * k] += b4.x; acc[4 * k + 1] += b4.y; acc[4 * k + 2] += b4.z; acc[4 * k + 3] += b4.w;
            }
          if (!LEAN && a.pre.p) {
#pragma unroll
            for (int k = 0; k < 8; ++k)
              if (k < ng && co0 + 4 * k < a.cout) {
                const float4 t4 = __ldg(reinterpret_cast<const float4*>(prep + co0 + 4 * k));
                acc[4 * k] += t4.x; acc[4 * k + 1] += t4.y; acc[4 * k + 2] += t4.z; acc[4 * k + 3] += t4.w;
              }
          }
          if (a.act == ACT_CROSS_SIGMOID) {                // co0 % 4 == 0: the odd channels are the scales
#pragma unroll
            for (int i = 1; i < 32; i += 2) acc[i] = __fdividef(1.f, 1.f + __expf(-(acc[i] + 2.f))) + a.eps;
          } else if (a.act == ACT_RELU) {
#pragma unroll
            for (int i = 0; i < 32; ++i) acc[i] = fmaxf(acc[i], 0.f);
          } else if (a.act != ACT_NONE) {
#pragma unroll
            for (int i = 0; i < 32; ++i) acc[i] = fmaxf(acc[i], 0.f) + slope * fminf(acc[i], 0.f);
          }
          if (a.alpha != 1.f) {
#pragma unroll
            for (int i = 0; i < 32; ++i) acc[i] *= a.alpha;
          }
          if (!LEAN && a.res1.p) {
#pragma unroll
            for (int k = 0; k < 8; ++k)
              if (k < ng && co0 + 4 * k < a.cout) {
                const float4 t4 = __ldg(reinterpret_cast<const float4*>(r1p + co0 + 4 * k));
                acc[4 * k] = fmaf(a.beta1, t4.x, acc[4 * k]); acc[4 * k + 1] = fmaf(a.beta1, t4.y, acc[4 * k + 1]);
                acc[4 * k + 2] = fmaf(a.beta1, t4.z, acc[4 * k + 2]); acc[4 * k + 3] = fmaf(a.beta1, t4.w, acc[4 * k + 3]);
              }
          }
          if (!LEAN && a.res2.p) {
#pragma unroll
            for (int k = 0; k < 8; ++k)
              if (k < ng && co0 + 4 * k < a.cout) {
                const float4 t4 = __ldg(reinterpret_cast<const float4*>(r2p + co0 + 4 * k));
                acc[4 * k] = fmaf(a.beta2, t4.x, acc[4 * k]); acc[4 * k + 1] = fmaf(a.beta2, t4.y, acc[4 * k + 1]);
                acc[4 * k + 2] = fmaf(a.beta2, t4.z, acc[4 * k + 2]); acc[4 * k + 3] = fmaf(a.beta2, t4.w, acc[4 * k + 3]);
              }
          }
          TR_ADD(tr_math, trm0);
          if (TMA_IN && a.flow.C) {   // fused FlowStep: the conv output (shift, scale pairs) is consumed here and never stored
            if (valid) {
              if (LEAN || a.flow.C == 12) flow_epilogue<12>(a.flow, acc, flow_s, flow_s + 24 * 24, p);
              else flow_epilogue<LEAN ? 12 : 24>(a.flow, acc, flow_s, flow_s + 24 * 24, p);
            }
            continue;
          }
          if ((!a.tma_out || (a.out2.p && !a.tma_out2)) && valid) {   // views the TMA unit cannot address: per-lane stores
#pragma unroll
            for (int k = 0; k < 8; ++k)
              if (k < ng && co0 + 4 * k < a.cout) {
                const float4 o = make_float4(acc[4 * k], acc[4 * k + 1], acc[4 * k + 2], acc[4 * k + 3]);
                if (!a.tma_out) direct_store(a.out, p, co0 + 4 * k, o);
                if (a.out2.p && !a.tma_out2) direct_store(a.out2, p, co0 + 4 * k, o);
              }
          }
          // phase outputs: the map strides over every second pixel, the box starts at the phase's own (fy, fx) offset
          const int ox = a.phase ? 2 * sx0 + (tcd.ph & 1) : sx0, oy = a.phase ? 2 * (sy0 + 4 * q) + (tcd.ph >> 1) : sy0 + 4 * q;
          if (a.tma_out) stage_store(a.out, &a.tmap_out, acc, ncol, co_base + n0, ox, oy, tcd.n);
          if (!LEAN && a.tma_out2) stage_store(a.out2, &a.tmap_out2, acc, ncol, co_base + n0, ox, oy, tcd.n);
        }
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      mbar_arrive(acc_empty + 8 * as);                 // every epilogue thread arrives: the accumulator stage is free
      TR_ADD(tr_work, tr1);
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // all bulk stores of this warp have completed
    __syncwarp();
#ifdef BFSR_TC_TRACE
    if (blockIdx.x == 0 && tid == 0) printf("[tc trace] epilogue: total %lld wait_acc_full %lld work %lld (tiles %d) | tmem_ld %lld math %lld store_wait %lld stage+issue %lld\n", clock64() - tr_start, tr_wait, tr_work, t_it, tr_ld, tr_math, tr_stw, tr_stg);
#endif
  } else if (warp == R::W_MMA || (warp >= R::W_MMAB && warp < R::W_MMAB + R::N_ISS_MAX - 1)) {
    // ===================== MMA issuers: whole warp walks the (uniform) loop, one elected lane issues =====================
    // A single thread sustains one small MMA per ~45-55 clk, the tensor pipe accepts one per ~40 (tools/micro/umma_rate.cu),
    // so macro tiles with several sub-tiles are split between two issuing warps (even / odd sub-tiles); every barrier
    // that recycles operand slots or releases the epilogue counts one tcgen05.commit per issuer.
    const int iss = warp == R::W_MMA ? 0 : warp - R::W_MMAB + 1;
    if (iss < a.n_iss) {
    const uint32_t idesc_wide = make_idesc(a.wide ? 2 * nt : nt), idesc_nt = make_idesc(nt);
    const uint32_t lo_rows = (uint32_t)nt * (ROWB >> 4);               // descriptor offset of the W_lo rows inside a tap image
    const uint32_t sbo = a.fold ? 8u * ROWB : (uint32_t)a.pitch * ROWB;   // fold: M tiles are runs of 128 consecutive halo-tile rows
    const int sub_cols = a.wide ? 2 * nt : nt;
    const bool merged = !a.fast && !a.wide;                             // three N = NT MMAs into the same accumulator columns
    // constant descriptor fields (LBO=1, SBO, version 1, SWIZZLE_64B); the start address is added per operand
    const uint64_t desc_a_hi = make_desc(0, sbo), desc_b_hi = make_desc(0, 8 * ROWB);
    uint32_t sub_off[4];
#pragma unroll
    for (int sub = 0; sub < 4; ++sub)
      sub_off[sub] = a.fold ? (uint32_t)(sub * 128) * (ROWB >> 4) : (uint32_t)((sub / a.sx) * 16 * a.pitch + (sub % a.sx) * 8) * (ROWB >> 4);
    const int kw = a.phase ? 2 : a.ks;                                   // taps per filter row (fold 2: ks = 1, the three 'taps' are the filter rows)
    const uint32_t row_step = (uint32_t)(a.pitch - (kw - 1)) * (ROWB >> 4);   // from the last tap of a row to the first of the next
    const int stages = a.ntaps / a.tps;
    int a_it = 0, w_it = 0, t_it = 0;
    TR_DECL(tr_acc = 0, tr_a = 0, tr_w = 0, tr_issue = 0); TR_T(tr_start);
    if (a.w_res && (int)blockIdx.x < a.total_tiles) mbar_wait(w_full, 0);     // resident weights: one wait for the whole set
    for (int p = TC_T_FIRST; p < TC_T_END; p += TC_T_STEP, ++t_it) {
      const int t = tile_of(p);
      const int as = t_it % a.nacc;
      const int ph = a.phase ? (t / a.n_ct) & 3 : 0;
      if (a.w_res) w_it = 0;                                                    // stage index = position inside the tile
      const uint32_t tap_base = a.phase ? (uint32_t)((ph >> 1) * a.pitch + (ph & 1)) * (ROWB >> 4) : 0u;
      TR_T(tr0);
      mbar_wait_relaxed(acc_empty + 8 * as, ((t_it / a.nacc) & 1) ^ 1);
      TR_ADD(tr_acc, tr0);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t d_base = tmem_base + as * acc_cols;
      for (int c = 0; c < a.n_chunks; ++c, ++a_it) {
        const int slot = a_it % a.na;
        TR_T(tr1);
        mbar_wait(a_full + 8 * slot, (a_it / a.na) & 1);
        TR_ADD(tr_a, tr1);
        const uint32_t a_hi = a_smem + slot * a.a_slot, a_lo = a_hi + a.a_plane;
        const uint64_t a_hi_d = desc_a_hi | (uint64_t)((a_hi & 0x3FFFF) >> 4), a_lo_d = desc_a_hi | (uint64_t)((a_lo & 0x3FFFF) >> 4);
        int nk = (a.cin - c * KC + 15) >> 4; nk = nk > 2 ? 2 : nk;
        // running tap geometry (no divisions on the issue path): kw taps per filter row, start offset of the phase
        int tdx = 0;
        uint32_t tap_off = tap_base;
        int stages_c = stages, pl = -1;
        if (a.n_pre) {                                       // one tap per weight stage; identity chunks take the centre tap only
          if (c >= a.n_main) { nk = 2; stages_c = 1; tap_off = (uint32_t)(a.halo * a.pitch + a.halo) * (ROWB >> 4); }
        } else if (a.phase == 2) {                           // one tap per weight stage; hi-res plane chunks use the tap table
          nk = 2;
          if (c < a.n_lo) stages_c = 4;
          else { pl = (c - a.n_lo) & 3; stages_c = a.pl_nt[ph][pl]; }
        }
        for (int st = 0; st < stages_c; ++st, ++w_it) {
          if (pl >= 0) tap_off = a.pl_off[ph][pl][st];
          const int ws = w_it % a.nw;
          TR_T(tr2);
          if (!a.w_res) mbar_wait(w_full + 8 * ws, (w_it / a.nw) & 1);
          TR_ADD(tr_w, tr2); TR_T(tr3);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          // One elected lane issues the whole stage.  Descriptors differ only in the 14-bit start-address field, so each
          // operand is the chunk/slot base descriptor plus a small running offset (uniform-datapath adds).
          uint64_t bd0 = desc_b_hi | (uint64_t)((((w_smem + ws * a.w_stage)) & 0x3FFFF) >> 4);
          for (int tt = 0; tt < a.tps; ++tt) {             // uniform loop; one elected lane issues each tap
            const uint32_t accf0 = (c | st | tt) ? 1u : 0u;
            if (elect_one()) {
#pragma unroll
              for (int sub = 0; sub < 4; ++sub) {
                if (sub < a.mt && (sub & (a.n_iss - 1)) == iss) {
                  const uint32_t d = d_base + sub * sub_cols;
                  const uint64_t ah = a_hi_d + tap_off + sub_off[sub], al = a_lo_d + tap_off + sub_off[sub];
                  umma_f16(d, ah, bd0, idesc_wide, accf0);
                  if (merged) umma_f16(d, ah, bd0 + lo_rows, idesc_nt, 1u);
                  if (!a.fast) umma_f16(d, al, bd0, idesc_nt, 1u);
                  if (nk > 1) {
                    umma_f16(d, ah + 2, bd0 + 2, idesc_wide, 1u);
                    if (merged) umma_f16(d, ah + 2, bd0 + lo_rows + 2, idesc_nt, 1u);
                    if (!a.fast) umma_f16(d, al + 2, bd0 + 2, idesc_nt, 1u);
                  }
                }
              }
            }
            bd0 += (uint32_t)a.w_slot >> 4;
            if (++tdx == kw) { tdx = 0; tap_off += row_step; } else tap_off += ROWB >> 4;
          }
          if (!a.w_res && elect_one()) {                                   // weight stage reusable once these MMAs retire
            if (CL == 2) umma_commit_mc(w_empty + 8 * ws, (uint16_t)3);    // ... in both CTAs of the cluster
            else umma_commit(w_empty + 8 * ws);
          }
          __syncwarp();
          TR_ADD(tr_issue, tr3);
        }
        if (elect_one()) umma_commit(a_empty + 8 * slot);      // A slot reusable
        __syncwarp();
      }
      if (elect_one()) umma_commit(acc_full + 8 * as);
      __syncwarp();
    }
#ifdef BFSR_TC_TRACE
    if (blockIdx.x == 0 && lane == 0 && iss == 0) printf("[tc trace] mma: total %lld wait_acc_empty %lld wait_a_full %lld wait_w_full %lld issue %lld (taps %d)\n", clock64() - tr_start, tr_acc, tr_a, tr_w, tr_issue, w_it);
#endif
    }
  } else {
    // ===================== weight producer: cp.async.bulk of pre-swizzled [W_hi;W_lo] images =====================
    const size_t tap_stride = (size_t)2 * nt * ROWB;    // packed image always holds both planes
    const uint32_t w_bytes = (uint32_t)(a.fast ? nt : 2 * nt) * ROWB;
    int w_it = 0;
    TR_DECL(tr_wait = 0); TR_T(tr_start);
    if (a.w_res) {     // the whole weight set of the (single) cout tile stays resident: one barrier, a.nw bulk copies, done
      if ((int)blockIdx.x < a.total_tiles && elect_one()) {
        mbar_expect_tx(w_full, w_bytes * a.tps * a.nw);
        for (int wi = 0; wi < a.nw; ++wi)
          bulk_g2s(w_smem + wi * a.w_stage, a.w + (size_t)wi * a.tps * tap_stride, w_bytes * a.tps, w_full);
      }
      __syncwarp();
    } else
    for (int p = TC_T_FIRST; p < TC_T_END; p += TC_T_STEP) {
      const int t = tile_of(p);
      const int ct = t % a.n_ct;
      const int wsel = a.phase ? ct * 4 + ((t / a.n_ct) & 3) : ct;       // phase mode: [cout tile][phase][chunk][tap]
      const unsigned char* wsrc = a.w + (size_t)wsel * (a.phase == 2 ? a.taps_tile : a.n_chunks * a.ntaps) * tap_stride;
      const int total = a.phase == 2 ? a.taps_tile : (a.n_pre ? a.n_main * a.ntaps + a.n_pre : a.n_chunks * (a.ntaps / a.tps));
      for (int wi = 0; wi < total; ++wi, ++w_it) {
        const int ws = w_it % a.nw;
        TR_T(tr0);
        mbar_wait_relaxed(w_empty + 8 * ws, ((w_it / a.nw) & 1) ^ 1);
        TR_ADD(tr_wait, tr0);
        if (elect_one()) {
          const uint32_t dst = w_smem + ws * a.w_stage;
          const unsigned char* src = wsrc + (size_t)wi * a.tps * tap_stride;
          mbar_expect_tx(w_full + 8 * ws, w_bytes * a.tps);
          if (CL == 2) {      // this CTA's half of every copy, delivered to both CTAs (each one's w_full counts the whole stage)
            if (!a.fast) { const uint32_t hb = (w_bytes * a.tps) >> 1; bulk_g2s_mc(dst + cl_rank * hb, src + (size_t)cl_rank * hb, hb, w_full + 8 * ws, (uint16_t)3); }
            else for (int tt = 0; tt < a.tps; ++tt) {
              const uint32_t hb = w_bytes >> 1;
              bulk_g2s_mc(dst + tt * a.w_slot + cl_rank * hb, src + (size_t)tt * tap_stride + (size_t)cl_rank * hb, hb, w_full + 8 * ws, (uint16_t)3);
            }
          } else
          if (!a.fast) bulk_g2s(dst, src, w_bytes * a.tps, w_full + 8 * ws);        // taps are contiguous: one bulk copy
          else for (int tt = 0; tt < a.tps; ++tt) bulk_g2s(dst + tt * a.w_slot, src + (size_t)tt * tap_stride, w_bytes, w_full + 8 * ws);
        }
        __syncwarp();
      }
    }
    if (CL == 2) {   // tail: every stage this CTA filled has been released by both CTAs, i.e. no remote arrive is still in flight
      for (int i = 0; i < a.nw && i < w_it; ++i) { const int it = w_it - 1 - i; mbar_wait_relaxed(w_empty + 8 * (it % a.nw), (it / a.nw) & 1); }
    }
#ifdef BFSR_TC_TRACE
    if (blockIdx.x == 0 && lane == 0) printf("[tc trace] wprod: total %lld wait_w_empty %lld (taps %d)\n", clock64() - tr_start, tr_wait, w_it);
#endif
  }
  __syncthreads();
  if (CL == 2) cluster_sync_all();       // neither CTA leaves while the other may still signal its barriers
  if (warp == R::W_MMA) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(a.tmem_cols) : "memory");
  }
}

// ------------------------------------------------------------------ host side

static int pick_nt(int cout) {
  const int r = (cout + 15) / 16 * 16;
  return r > 128 ? 128 : r;
}

// h: host fp32 packed [tap][cin_pad][cout_pad] (the fp32 kernel's layout)
void pack_conv_tc(ConvW& c, const std::vector<float>& h, int min_cin_arg) {
  using namespace tc;
  static const int min_cin_env = getenv("BFSR_TC_MIN_CIN") ? atoi(getenv("BFSR_TC_MIN_CIN")) : 32;
  if (c.cin < (min_cin_arg >= 0 ? min_cin_arg : min_cin_env)) return;
  const int taps = c.ks * c.ks;
  const int nt = pick_nt(c.cout);
  const int n_tiles = (c.cout + nt - 1) / nt, n_chunks = (c.cin + KC - 1) / KC;
  const size_t tap_elems = (size_t)2 * nt * (ROWB / 2);
  // single-tile convs with Cout % 32 == 0 carry Cout/32 identity tap images after the conv's own taps: a BF16X2 pre-activation
  // tensor is then added INSIDE the GEMM as extra K chunks (hi*1 + lo*1 is exact in the fp32 accumulator) instead of by
  // epilogue loads
  const int n_id = (n_tiles == 1 && c.cout % KC == 0) ? c.cout / KC : 0;
  std::vector<unsigned short> img(((size_t)n_tiles * n_chunks * taps + n_id) * tap_elems, 0);
  for (int j = 0; j < n_id; ++j) {
    unsigned short* dst = img.data() + ((size_t)n_chunks * taps + j) * tap_elems;
    for (int k = 0; k < KC; ++k) {
      const int r = j * KC + k;             // output channel r takes input channel k of identity chunk j
      dst[(size_t)r * 32 + (((k >> 3) ^ ((r >> 1) & 3)) << 3) + (k & 7)] = 0x3F80;   // bf16(1.0) in the W_hi rows; W_lo stays 0
    }
  }
  c.tc_n_id = n_id;
  for (int t = 0; t < n_tiles; ++t)
    for (int ch = 0; ch < n_chunks; ++ch)
      for (int tap = 0; tap < taps; ++tap) {
        unsigned short* dst = img.data() + (((size_t)t * n_chunks + ch) * taps + tap) * tap_elems;
        for (int r = 0; r < nt; ++r) {
          const int co = t * nt + r;
          for (int k = 0; k < KC; ++k) {
            const int ci = ch * KC + k;
            float w = 0.f;
            if (co < c.cout && ci < c.cin) w = h[((size_t)tap * c.cin_pad + ci) * c.cout_pad + co];
            const unsigned short hi = f2bf(w), lo = f2bf(w - bf2f(hi));
            const int j = k >> 3, e = k & 7;
            dst[(size_t)r * 32 + ((j ^ ((r >> 1) & 3)) << 3) + e] = hi;
            const int r2 = nt + r;
            dst[(size_t)r2 * 32 + ((j ^ ((r2 >> 1) & 3)) << 3) + e] = lo;
          }
        }
      }
  CUDA_OK(cudaMalloc(&c.w_tc, img.size() * 2));
  CUDA_OK(cudaMemcpy(c.w_tc, img.data(), img.size() * 2, cudaMemcpyHostToDevice));
  c.tc_kchunks = n_chunks; c.tc_npad = nt;
  if (c.ks == 3 && c.cin == 64 && c.cout <= 24 && c.cout % 4 == 0) {   // tap-folded image (see TcArgs::fold)
    const int np = (9 * c.cout + 15) / 16 * 16;
    std::vector<unsigned short> f((size_t)2 * 2 * np * (ROWB / 2), 0);
    for (int ch = 0; ch < 2; ++ch)
      for (int tap = 0; tap < 9; ++tap)
        for (int co = 0; co < c.cout; ++co) {
          const int r = tap * c.cout + co;
          for (int k = 0; k < KC; ++k) {
            const float w = h[((size_t)tap * c.cin_pad + ch * KC + k) * c.cout_pad + co];
            const unsigned short hi = f2bf(w), lo = f2bf(w - bf2f(hi));
            unsigned short* dst = f.data() + (size_t)ch * 2 * np * 32;
            const int j = k >> 3, e = k & 7, r2 = np + r;
            dst[(size_t)r * 32 + ((j ^ ((r >> 1) & 3)) << 3) + e] = hi;
            dst[(size_t)r2 * 32 + ((j ^ ((r2 >> 1) & 3)) << 3) + e] = lo;
          }
        }
    CUDA_OK(cudaMalloc(&c.w_tc_fold, f.size() * 2));
    CUDA_OK(cudaMemcpy(c.w_tc_fold, f.data(), f.size() * 2, cudaMemcpyHostToDevice));
    c.tc_fold_np = np;
  }
  if (c.ks == 3 && c.cout <= 32 && c.cout % 4 == 0) {   // dx-folded image (see TcArgs::fold = 2): [chunk][dy] images of [hi ; lo] rows dx*cb + co
    const int cb = c.cout <= 16 ? 16 : 32, np = 3 * cb;
    std::vector<unsigned short> f((size_t)n_chunks * 3 * 2 * np * (ROWB / 2), 0);
    for (int ch = 0; ch < n_chunks; ++ch)
      for (int dy = 0; dy < 3; ++dy) {
        unsigned short* dst = f.data() + ((size_t)ch * 3 + dy) * 2 * np * 32;
        for (int dx = 0; dx < 3; ++dx)
          for (int co = 0; co < c.cout; ++co) {
            const int r = dx * cb + co, r2 = np + r;
            for (int k = 0; k < KC; ++k) {
              const int ci = ch * KC + k;
              const float w = ci < c.cin ? h[((size_t)(dy * 3 + dx) * c.cin_pad + ci) * c.cout_pad + co] : 0.f;
              const unsigned short hi = f2bf(w), lo = f2bf(w - bf2f(hi));
              const int j = k >> 3, e = k & 7;
              dst[(size_t)r * 32 + ((j ^ ((r >> 1) & 3)) << 3) + e] = hi;
              dst[(size_t)r2 * 32 + ((j ^ ((r2 >> 1) & 3)) << 3) + e] = lo;
            }
          }
      }
    CUDA_OK(cudaMalloc(&c.w_tc_f3, f.size() * 2));
    CUDA_OK(cudaMemcpy(c.w_tc_f3, f.data(), f.size() * 2, cudaMemcpyHostToDevice));
    c.tc_f3_cb = cb;
  }
}

static bool vec4(const View& v);
static void make_tmap_out(CUtensorMap* tm, const View& v, int estride, int bw = 8, int bh = 4);
static bool tma_out_ok(const View& v);
static bool vec4(const View& v) { return v.fmt == F32 && v.cs % 4 == 0 && v.coff % 4 == 0 && ((uintptr_t)v.p % 16) == 0; }
// BF16X2 operand views: TMA needs 16-byte global strides and base; the register producer needs 16-byte channel groups
static bool bf_in_ok(const View& v) {
  return v.fmt == BF16X2 && v.cs % 8 == 0 && v.coff % 8 == 0 && v.plane % 8 == 0 && ((uintptr_t)v.p % 16) == 0;
}
static bool out_ok(const View& v) {
  return vec4(v) || (v.fmt == BF16X2 && v.cs % 4 == 0 && v.coff % 4 == 0 && v.plane % 4 == 0 && ((uintptr_t)v.p % 8) == 0);
}
static bool shapes_ok(const ConvW& w, const View& in, const View& out, const ConvEpi& epi) {
  return w.cout % 4 == 0 && (bf_in_ok(in) || (vec4(in) && (w.cin % 8 == 0 || w.cin <= 16))) && out_ok(out) &&
         (!epi.out2 || out_ok(*epi.out2)) && (!epi.pre || vec4(*epi.pre) || (bf_in_ok(*epi.pre) && bf_in_ok(in) && w.tc_n_id * tc::KC == w.cout)) && (!epi.res1 || vec4(*epi.res1)) &&
         (!epi.res2 || vec4(*epi.res2));
}

static void make_tmap_plane(CUtensorMap* tm, const View& v, int py, int px, int pitch, int hrows);

// Shapes the tensor-core kernel takes (everything else runs on the fp32 CUDA-core kernels).  Batch-independent, so results
// do not depend on chunking.  The kernel has no scalar fallbacks (keeps its instruction footprint small).
bool conv_tc_eligible(const ConvW& w, const View& in, const View& out, const ConvEpi& epi) {
  return w.w_tc != nullptr && !w.tc_phase && g_conv_mode != 2 && shapes_ok(w, in, out, epi);
}

// parity plane (py, px) of a BF16X2 NHWC view with even H, W: pixels (2Y+py, 2X+px) as a dense (C, W/2, H/2, N, plane) tensor
static void make_tmap_plane(CUtensorMap* tm, const View& v, int py, int px, int pitch, int hrows) {
  BFSR_CHECK(v.H % 2 == 0 && v.W % 2 == 0, "parity planes need even spatial dims");
  const cuuint64_t dims[5] = {(cuuint64_t)(v.coff + v.C), (cuuint64_t)(v.W / 2), (cuuint64_t)(v.H / 2), (cuuint64_t)v.N, 2};
  const cuuint64_t strides[4] = {(cuuint64_t)v.cs * 4, (cuuint64_t)v.W * v.cs * 4, (cuuint64_t)v.H * v.W * v.cs * 2,
                                 (cuuint64_t)v.plane * 2};
  const cuuint32_t box[5] = {(cuuint32_t)tc::KC, (cuuint32_t)pitch, (cuuint32_t)hrows, 1, 1};
  const cuuint32_t es[5] = {1, 1, 1, 1, 1};
  void* base = (void*)((__nv_bfloat16*)v.p + ((long long)py * v.W + px) * v.cs);
  const CUresult r = encode_tiled()(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, base, dims, strides, box, es,
                                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
                                    CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  BFSR_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(plane) failed (%d) for view C=%d cs=%d %dx%dx%d", (int)r, v.C, v.cs, v.N, v.H, v.W);
}
static bool tma_out_ok(const View& v) {
  if (((uintptr_t)v.p % 16) != 0) return false;
  return v.fmt == F32 ? v.cs % 4 == 0 : (v.cs % 8 == 0 && v.plane % 8 == 0);
}
// output map: box = [32 channels, 8, 4 pixels] = the accumulator rows of one epilogue warp; fp32 rows are 128 B
// (SWIZZLE_128B), bf16 rows 64 B per plane (SWIZZLE_64B); the channel extent stops at the end of the view (tail clipped)
static void make_tmap_out(CUtensorMap* tm, const View& v, int estride, int bw, int bh) {
  const int es_b = v.fmt == F32 ? 4 : 2;
  const cuuint64_t dims[5] = {(cuuint64_t)(v.coff + v.C), (cuuint64_t)v.W, (cuuint64_t)v.H, (cuuint64_t)v.N, 2};
  const cuuint64_t strides[4] = {(cuuint64_t)v.cs * es_b, (cuuint64_t)v.W * v.cs * es_b, (cuuint64_t)v.H * v.W * v.cs * es_b,
                                 (cuuint64_t)v.plane * es_b};
  const cuuint32_t box[5] = {32, (cuuint32_t)(bw * estride), (cuuint32_t)(bh * estride), 1, 1};   // estride 2: every second pixel (phase outputs); dx-folded convs: 30 x 1
  const cuuint32_t es[5] = {1, (cuuint32_t)estride, (cuuint32_t)estride, 1, 1};
  const CUresult r = v.fmt == F32
      ? encode_tiled()(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, v.p, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)
      : encode_tiled()(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, v.p, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  BFSR_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(out) failed (%d) for view C=%d cs=%d %dx%dx%d", (int)r, v.C, v.cs, v.N, v.H, v.W);
}

static int g_num_sms = 0;

// phase = false: out (N,H,W) = conv_ks(in) (+ optional nearest-2x folded into the loader).
// phase = true : out (N,2H,2W) (+)= conv3x3(nearest2x(in)) evaluated on the LOW-RES grid as four 2x2 phase convs with
//                pre-summed weights (exact in real arithmetic, 16/36 of the MACs); `in` is the low-res tensor.
static void launch_tc(const ConvW& w, const View& in, const View& out, const ConvEpi& epi, int in_mode, int phase, cudaStream_t s,
                      const View* in2 = nullptr, int fold = 0, bool allow16 = true) {
  using namespace tc;
  BFSR_CHECK(w.w_tc, "conv_tc: weights not packed for the tcgen05 path");
  BFSR_CHECK(in.C + (in2 ? in2->C : 0) == w.cin && out.C == w.cout && in.N == out.N, "conv_tc: shape mismatch");
  BFSR_CHECK(shapes_ok(w, in, out, epi), "conv_tc: operand views are not addressable by the tensor-core kernel "
             "(16-byte aligned fp32 or BF16X2 views, Cout %% 4 == 0)");
  if (in_mode == IN_UP2 || phase) BFSR_CHECK(in.H * 2 == out.H && in.W * 2 == out.W, "conv_tc(up2): spatial mismatch");
  else BFSR_CHECK(in.H == out.H && in.W == out.W, "conv_tc: spatial mismatch");
  if (out.npix() == 0) return;
  const int gH = phase ? in.H : out.H, gW = phase ? in.W : out.W;      // grid the GEMM rows live on
  TcArgs a;
  a.in = in; a.out = out; a.out2 = epi.out2 ? *epi.out2 : View();
  const bool pre_gemm = epi.pre && epi.pre->fmt == BF16X2;    // pre-activation enters as identity K chunks (checked below)
  a.pre = (epi.pre && !pre_gemm) ? *epi.pre : View(); a.res1 = epi.res1 ? *epi.res1 : View(); a.res2 = epi.res2 ? *epi.res2 : View();
  a.w = (const unsigned char*)(fold == 2 ? w.w_tc_f3 : (fold ? w.w_tc_fold : w.w_tc)); a.bias = w.bias;
  a.cin = w.cin; a.cout = w.cout; a.nt = fold == 2 ? 3 * w.tc_f3_cb : (fold ? w.tc_fold_np : w.tc_npad); a.n_chunks = w.tc_kchunks;
  a.cb = w.tc_f3_cb;
  a.n_ct = cdiv(w.cout, w.tc_npad);
  a.H = gH; a.W = gW; a.N = out.N; a.in_mode = phase ? (int)IN_DIRECT : in_mode; a.act = epi.act;
  a.eps = epi.eps; a.alpha = epi.alpha; a.beta1 = epi.beta1; a.beta2 = epi.beta2;
  a.fast = g_conv_mode == 1;
  a.phase = phase;
  // macro tile: as many 128-pixel sub-tiles as fit in 512 TMEM columns (and the image); weights shared by all of them
  // accurate mode.  Measured M=128,K=16 SS-mode MMA cost on B200 (tools/micro/umma_rate.cu): 45.5 clk for N <= 32, 48 @ 64,
  // 56 @ 96, then N/2 (64 @ 128, 128 @ 256): an MMA narrower than N = 128 is bound by its fixed cost, so NT <= 64 issues
  // A_hi x [W_hi;W_lo] as ONE N = 2NT MMA plus A_lo x W_hi (e.g. NT = 64: 64 + 48 clk instead of 3 x 48); NT > 64 issues the
  // three products as separate N = NT MMAs into the same columns (same MMA time, half the TMEM -> two accumulator stages)
  // Round 2, measured on the whole step (profiles/r2_step_knobs.md): the three-MMA form is never slower than the wide one and up to
  // 23 % faster on the small-K convs (z-dependent coupling conv 0.53 -> 0.41 ms, 1x1 0.33 -> 0.29 ms) -- the wide form doubles the
  // TMEM loads and adds a pass to an epilogue that is the bottleneck there, and halves the sub-tiles per weight stream everywhere
  // else.  BFSR_TC_WIDE=1 restores the wide form for NT <= 64.
  static const int force_wide = getenv("BFSR_TC_WIDE") ? atoi(getenv("BFSR_TC_WIDE")) : 0;
  a.wide = (!a.fast && fold != 1 && ((a.nt <= 64 && force_wide == 1) || (force_wide == 3 && fold == 2 && a.nt <= 128))) ? 1 : 0;
  // sixteen epilogue warps for the epilogue-latency-bound convs: TMA-fed, few (chunk, tap) MMA groups per output tile (see
  // conv_tc_kernel); the geometry below must then give at least four (sub-tile, 32-channel block) items per work tile
  static const int ew16_maxk = getenv("BFSR_TC_EW16_MAXK") ? atoi(getenv("BFSR_TC_EW16_MAXK")) : 2;
  const bool tma_in = in.fmt == BF16X2 && in_mode == IN_DIRECT && phase != 1;
  const int k_steps = w.tc_kchunks * (fold == 2 ? 3 : w.ks * w.ks) + (pre_gemm ? w.tc_n_id : 0);
  static const int ew16_flow = getenv("BFSR_TC_EW16_FLOW") ? atoi(getenv("BFSR_TC_EW16_FLOW")) : 1;   // C = 12 heads: 21.3 -> 17.1 ms per step
  bool ew16 = allow16 && tma_in && !a.fast && phase == 0 && fold != 1 && (k_steps <= ew16_maxk || (ew16_flow && fold == 2 && epi.flow)) && a.n_ct == 1 && !epi.res1 && !epi.res2 && !epi.out2 &&
              (!epi.pre || pre_gemm) && (!epi.flow || epi.flow->C == 12);
  if (ew16 && fold == 2 && a.nt <= 64) a.wide = 0;                   // 3*cb columns per M tile: four M tiles in two stages
  const int sub_cols = a.wide ? 2 * a.nt : a.nt;
  int mt = 256 / sub_cols; mt = mt >= 4 ? 4 : (mt >= 2 ? 2 : 1);      // two accumulator stages whenever they fit
  static const int mt_cap = getenv("BFSR_TC_MT_MAX") ? atoi(getenv("BFSR_TC_MT_MAX")) : 4;
  if (mt > mt_cap) mt = mt_cap;
  // The level-1 phase convs (NT = 128, 50 tap images = 0.8 MB of weights per tile) are bound by L2 -> SM traffic, not by the
  // MMAs: four sub-tiles per weight stream (one accumulator stage, exposed epilogue) trade 12 % of overlap for 31 % less traffic
  static const bool mt4_phase = getenv("BFSR_TC_MT4") && atoi(getenv("BFSR_TC_MT4")) == 1;
  if (mt4_phase && phase == 2 && sub_cols == 128) mt = 4;
  if (!g_num_sms) { int dev = 0; CUDA_OK(cudaGetDevice(&dev)); CUDA_OK(cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev)); }
  auto arrange = [&](int m, int& sx, int& sy) {          // sub-tile arrangement of a macro tile of m sub-tiles
    sx = 1; sy = 1;
    if (m == 4) {
      if (gW > 8 && gH > 16) { sx = 2; sy = 2; }
      else if (gH > 16) { sy = 2; }
      else if (gW > 8) { sx = 2; }
    } else if (m == 2) {
      if (gH > 16) sy = 2; else if (gW > 8) sx = 2;
    }
  };
  // Wave quantisation: one CTA per SM walks the tile list, so a launch with 200 tiles takes two rounds of which the second is a
  // third full (4 tiles of 160x160 per GPU at 8 GPUs: the RDB's conv5 ran at 64 % of its batch-32 efficiency).  Cost model:
  // rounds x (sub-tiles per tile + 0.6 for the halo / weight stream / fixed cost of a tile); the cheapest macro tile wins, ties go to
  // the larger one.  Tile shape never changes the arithmetic of an output pixel, so results stay independent of the batch size.
  if (fold == 2 && !ew16 && mt > 2 && !(getenv("BFSR_F3_MT") && atoi(getenv("BFSR_F3_MT")) == 4)) mt = 2;   // dx-folded: two M tiles per work tile
  if (fold != 1) {
    const long long per = (long long)out.N * a.n_ct * (phase ? 4 : 1);
    int best = mt; double best_cost = 1e300;
    for (int m = mt; m >= 1; m >>= 1) {
      long long tiles;
      if (fold == 2) tiles = (long long)cdiv(gW, 30) * cdiv(gH, 4 * m) * per;
      else { int sx, sy; arrange(m, sx, sy); tiles = (long long)cdiv(gW, 8 * sx) * cdiv(gH, 16 * sy) * per; }
      const double cost = (double)((tiles + g_num_sms - 1) / g_num_sms) * (m + 0.6);
      if (cost < best_cost * (1.0 - 1e-9)) { best_cost = cost; best = m; }
    }
    mt = best;
  }
  arrange(mt, a.sx, a.sy);
  a.mt = a.sx * a.sy;
  a.ks = w.ks; a.ntaps = phase ? 4 : w.ks * w.ks; a.halo = w.ks / 2;
  a.fold = fold; a.tile_w = 8 * a.sx; a.tile_h = 16 * a.sy;
  int fold_p = 0, fold_r = 0;
  if (fold == 1) {
    BFSR_CHECK(w.w_tc_fold && phase == 0 && in.fmt == BF16X2 && in_mode == IN_DIRECT && !a.fast && !epi.pre && !epi.res1 && !epi.res2 &&
               !epi.out2 && epi.alpha == 1.f && (epi.act == ACT_NONE || epi.act == ACT_CROSS_SIGMOID) && (epi.flow || vec4(out)),
               "conv_tc(fold): unsupported epilogue / operand combination");
    // one GEMM "tap" per chunk over the whole halo tile; M tiles = runs of 128 halo-raster rows; N = 9*Cout (padded to 16)
    a.tile_w = a.tile_h = w.cout <= 12 ? 16 : 14;            // 3 x 112 / 2 x 224 accumulator columns
    fold_p = a.tile_w + 2; fold_r = a.tile_h + 2;
    a.mt = cdiv(fold_p * fold_r, 128); a.sx = a.sy = 1;
    a.ks = 1; a.ntaps = 1; a.halo = 1;
    BFSR_CHECK(a.mt * a.nt <= 512 && a.mt <= 4, "conv_tc(fold): accumulators do not fit TMEM");
  }
  if (fold == 2) {
    BFSR_CHECK(w.w_tc_f3 && phase == 0 && bf_in_ok(in) && in_mode == IN_DIRECT && !epi.pre && !epi.res1 && !epi.res2 && !epi.out2 &&
               epi.alpha == 1.f && (epi.flow || out_ok(out)), "conv_tc(dx-fold): unsupported epilogue / operand combination");
    // raster of pitch 32: M tile = 4 image rows of 32 positions (30 valid outputs each); the halo tile adds one row above / below
    static const int f3_mt = getenv("BFSR_F3_MT") ? atoi(getenv("BFSR_F3_MT")) : 0;
    a.mt = ew16 ? mt : (mt >= 2 ? 2 : 1);
    if (f3_mt == 4 && 4 * sub_cols <= 512) a.mt = 4;
    if (a.mt > mt) a.mt = mt;
    while (a.mt > 1 && 4 * (a.mt - 1) >= gH) --a.mt;              // small images: no M tile entirely below the image
    a.sx = 1; a.sy = a.mt;
    a.tile_w = 30; a.tile_h = 4 * a.mt;
    fold_p = 32; fold_r = a.tile_h + 2;
    a.ks = 1; a.ntaps = 3; a.halo = 1;
  }
  a.flow = FlowEpi();
  if (epi.flow) {
    const FlowEpi& f = *epi.flow;
    BFSR_CHECK((f.C == 12 || f.C == 24) && w.cout == f.C && a.n_ct == 1 && phase == 0 && in.fmt == BF16X2 && in_mode == IN_DIRECT && epi.act == ACT_CROSS_SIGMOID && !epi.out2 &&
               vec4(f.z_in) && vec4(f.z_out) && f.z_in.C == f.C && f.z_out.C == f.C && f.z_in.npix() == out.npix() &&
               f.z_out.npix() == out.npix() && (!f.hF.p || (vec4(f.hF) && f.hF.C == 2 * f.C)) && (!f.has_mix || (f.M && f.cvec)) &&
               (!f.z1op.p || (f.z1op.fmt == BF16X2 && f.z1op.cs % 8 == 0 && f.z1op.coff % 8 == 0 && f.z1op.plane % 8 == 0 &&
                              f.z1op.C == ((f.C / 2 + 7) & ~7))),
               "conv_tc: fused FlowStep epilogue needs C in {12,24}, a single cout tile and 16-byte addressable fp32 z / hF views");
    a.flow = f;
  }
  a.n_lo = 0; a.taps_tile = 0; a.in2 = in2 ? *in2 : View();
  a.n_pre = 0; a.n_main = a.n_chunks;
  if (pre_gemm) {
    BFSR_CHECK(phase == 0 && in.fmt == BF16X2 && in_mode == IN_DIRECT && w.tc_n_id * KC == w.cout && epi.pre->C == w.cout &&
               bf_in_ok(*epi.pre) && epi.pre->npix() == out.npix(),
               "conv_tc: a BF16X2 pre-activation needs a TMA-fed single-tile conv with Cout %% 32 == 0");
    a.n_pre = w.tc_n_id; a.n_chunks = a.n_main + a.n_pre; a.in2 = *epi.pre;
  }
  memset(a.pl_off, 0, sizeof a.pl_off); memset(a.pl_nt, 0, sizeof a.pl_nt);
  static const int max_iss = getenv("BFSR_TC_ISSUERS") ? atoi(getenv("BFSR_TC_ISSUERS")) : 2;
  a.n_iss = (a.mt >= 2 && max_iss >= 2) ? 2 : 1;
  a.pitch = 8 * a.sx + 2 * a.halo; a.hrows = 16 * a.sy + 2 * a.halo;
  if (fold) { a.pitch = fold_p; a.hrows = fold_r; }
  a.a_plane = ((fold == 1 ? a.mt * 128 : a.pitch * a.hrows) * ROWB + 1023) / 1024 * 1024;
  a.a_slot = (a.fast ? 1 : 2) * a.a_plane;
  a.w_slot = 2 * a.nt * ROWB;
  a.nacc = 2 * a.mt * sub_cols <= 512 ? 2 : 1;
  // four stages when they fit (small accumulators): a deeper MMA -> epilogue queue smooths the bubbles of the short-K convs whose
  // three pipeline stages (TMA, MMA, epilogue) are each ~50 % busy (BFSR_TC_NACC4=1; measured neutral on the whole step, so two stages stay the default)
  static const bool nacc4 = getenv("BFSR_TC_NACC4") && atoi(getenv("BFSR_TC_NACC4")) == 1;   // measured neutral (and worse with the smaller tiles it needs): opt-in
  if (nacc4 && fold != 1 && 4 * a.mt * sub_cols <= 512) a.nacc = 4;
  uint32_t cols = 32; while ((int)cols < a.nacc * a.mt * sub_cols) cols <<= 1;
  a.tmem_cols = cols;
  a.tiles_x = cdiv(gW, a.tile_w); a.tiles_y = cdiv(gH, a.tile_h);
  a.total_tiles = a.tiles_x * a.tiles_y * out.N * a.n_ct * (phase ? 4 : 1);
  {
    auto magic = [](int d) { return (uint32_t)((((uint64_t)1 << 32) + (uint64_t)d - 1) / (uint64_t)d); };
    a.m_nct = magic(a.n_ct); a.m_tx = magic(a.tiles_x); a.m_ty = magic(a.tiles_y);
    BFSR_CHECK((uint64_t)a.total_tiles * (uint64_t)std::max(std::max(a.n_ct, a.tiles_x), a.tiles_y) < ((uint64_t)1 << 32) && a.total_tiles < (1 << 24),
               "conv_tc: tile count out of range for the reciprocal tile decode");
  }
  // weight stages: as many taps per stage as fit ~48 KB (fewer barrier round trips on the MMA issue path), 2-4 stages
  a.tps = 1;
  if (a.n_pre) {
    // one tap per weight stage (the identity chunks have a single tap)
  } else if (phase == 2) {
    BFSR_CHECK(in2 && bf_in_ok(in) && bf_in_ok(*in2) && in.C % KC == 0 && in2->C % KC == 0 && in2->H == out.H && in2->W == out.W &&
               in2->N == out.N && in_mode == IN_DIRECT && w.ks == 3, "conv_tc(single-pass phase): operand views");
    a.n_lo = in.C / KC;
    a.n_chunks = a.n_lo + 4 * (in2->C / KC);
    a.taps_tile = a.n_lo * 4 + (in2->C / KC) * 9;
    for (int ph = 0; ph < 4; ++ph)
      for (int pl = 0; pl < 4; ++pl) {
        int n = 0;
        for (int dy = 0; dy < 3; ++dy)
          for (int dx = 0; dx < 3; ++dx) {
            const int u = (ph >> 1) + dy - 1, v = (ph & 1) + dx - 1;               // hi-res offset of the tap from (2y, 2x)
            if ((u & 1) != (pl >> 1) || (v & 1) != (pl & 1)) continue;
            const int sy_ = u < 0 ? -1 : u >> 1, sx_ = v < 0 ? -1 : v >> 1;        // floor(u / 2): row / col shift inside the plane
            a.pl_off[ph][pl][n++] = (unsigned short)(((1 + sy_) * a.pitch + (1 + sx_)) * (ROWB >> 4));
          }
        a.pl_nt[ph][pl] = (unsigned char)n;
      }
  } else
  for (int cand : {9, 4, 3, 2}) if (a.ntaps % cand == 0 && cand * a.w_slot <= 48 * 1024) { a.tps = cand; break; }
  a.w_stage = a.tps * a.w_slot;
  a.na = 2;
  ew16 = ew16 && a.mt * (fold == 2 ? 1 : cdiv(a.nt > a.cout ? a.cout : a.nt, 32)) >= 4;
  const int ew = ew16 ? 16 : BFSR_EPI_WARPS;
  a.stg_bytes = (fold == 2 && epi.flow) ? 0 : stg_bytes(ew);       // the dx-folded head with a FlowStep epilogue stores per lane
  const int fixed = a.na * a.a_slot + 1024 + a.stg_bytes + 256 + bias_bytes(ew) + (epi.flow ? FLOW_BYTES : 0);
  a.nw = (MAX_SMEM - fixed) / a.w_stage; a.nw = a.nw > NW_MAX ? NW_MAX : a.nw;
  while (a.nw < 2 && a.tps > 1) {   // not enough room for double buffering: shrink the stage
    int next = 1; for (int cand : {4, 3, 2}) if (cand < a.tps && a.ntaps % cand == 0) { next = cand; break; }
    a.tps = next; a.w_stage = a.tps * a.w_slot; a.nw = (MAX_SMEM - fixed) / a.w_stage; a.nw = a.nw > NW_MAX ? NW_MAX : a.nw;
  }
  if (a.nw < 2 && ew16) { launch_tc(w, in, out, epi, in_mode, phase, s, in2, fold, false); return; }   // 16-warp geometry does not fit: 8 warps
  BFSR_CHECK(a.nw >= 2, "conv_tc: no room for two weight stages");
  // weight-stationary when the whole set fits next to two A slots (tiny-K convs re-streamed as many weight bytes per tile
  // from L2 as activations: the z-dependent coupling conv moved 5.9 TB/s L2->SM under ncu)
  a.w_res = 0;
  {
    static const bool no_res = getenv("BFSR_TC_WRES") && atoi(getenv("BFSR_TC_WRES")) == 0;
    const int stages_tile = a.n_pre ? a.n_main * a.ntaps + a.n_pre : a.n_chunks * (a.ntaps / a.tps);
    if (!no_res && phase == 0 && fold != 1 && !a.fast && a.n_ct == 1 && fixed + stages_tile * a.w_stage <= MAX_SMEM) { a.w_res = 1; a.nw = stages_tile; }
  }
  // left-over shared memory deepens the A ring (convs with little MMA work per chunk are bound by TMA latency otherwise)
  // the dx-folded heads with a FlowStep epilogue keep TWO A slots: their epilogue's own global loads (z, hF) are latency-critical
  // and queue behind a deeper TMA prefetch (measured 0.62 -> 0.54 ms per level-1 step)
  static const int na_env = getenv("BFSR_TC_NA") ? atoi(getenv("BFSR_TC_NA")) : 0;
  const int na_max = na_env ? na_env : ((fold == 2 && epi.flow) ? 2 : NA_MAX);
  while (a.na < na_max && a.na < NA_MAX && fixed + (a.na - 1) * a.a_slot + a.nw * a.w_stage <= MAX_SMEM) ++a.na;
  const int smem = fixed + (a.na - 2) * a.a_slot + a.nw * a.w_stage;
  BFSR_CHECK(smem <= MAX_SMEM, "conv_tc: smem budget exceeded (%d)", smem);
  BFSR_CHECK(a.hrows * a.pitch * 4 <= MAXI * NPROD, "conv_tc: producer item budget exceeded");
  a.tma = (in.fmt == BF16X2 && in_mode == IN_DIRECT && phase != 1) ? 1 : 0;
  a.in_bf = in.fmt == BF16X2 ? 1 : 0;
  memset(&a.tmap, 0, sizeof a.tmap);
  if (a.tma) make_tmap(&a.tmap, in, a.pitch, a.hrows);
  memset(a.tmap_pl, 0, sizeof a.tmap_pl);
  if (phase == 2) for (int pl = 0; pl < 4; ++pl) make_tmap_plane(&a.tmap_pl[pl], *in2, pl >> 1, pl & 1, a.pitch, a.hrows);
  if (a.n_pre) make_tmap(&a.tmap_pl[0], a.in2, a.pitch, a.hrows);
  memset(&a.tmap_out, 0, sizeof a.tmap_out); memset(&a.tmap_out2, 0, sizeof a.tmap_out2);
  static const bool no_tma_out = getenv("BFSR_NO_TMA_OUT") && atoi(getenv("BFSR_NO_TMA_OUT"));
  a.tma_out = (!no_tma_out && tma_out_ok(out) && !epi.flow && fold != 1) ? 1 : 0;
  a.tma_out2 = (epi.out2 && !no_tma_out && tma_out_ok(*epi.out2)) ? 1 : 0;
  if (a.tma_out) { if (fold == 2) make_tmap_out(&a.tmap_out, out, 1, 30, 1); else make_tmap_out(&a.tmap_out, out, phase ? 2 : 1); }
  if (a.tma_out2) make_tmap_out(&a.tmap_out2, *epi.out2, phase ? 2 : 1);
  if (!g_num_sms) { int dev = 0; CUDA_OK(cudaGetDevice(&dev)); CUDA_OK(cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev)); }
  CUDA_OK(cudaFuncSetAttribute(conv_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, MAX_SMEM));
  CUDA_OK(cudaFuncSetAttribute(conv_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, MAX_SMEM));
  int grid = a.total_tiles < g_num_sms ? a.total_tiles : g_num_sms;
  // 2-CTA clusters with multicast weight stages (opt-in until measured: BFSR_TC_CLUSTER=1 level-1 phase convs, 2 = every
  // TMA-fed conv that streams its weights): needs an even number of spatial tiles so that the pair list is complete
  static const int cl_env = getenv("BFSR_TC_CLUSTER") ? atoi(getenv("BFSR_TC_CLUSTER")) : 0;
  const bool cl2 = cl_env > 0 && (phase == 2 || cl_env >= 2) && a.tma && !a.w_res && fold != 1 && ((a.tiles_x * a.tiles_y * out.N) & 1) == 0 &&
                   a.total_tiles >= 4 && (a.w_stage & 31) == 0;
  if (cl2) {
    CUDA_OK(cudaFuncSetAttribute(conv_tc_kernel<true, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, MAX_SMEM));
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.blockDim = dim3(Roles<true>::NTHREADS); cfg.dynamicSmemBytes = smem; cfg.stream = s; cfg.attrs = at; cfg.numAttrs = 1;
    static int max_clusters = 0;       // co-schedulable CTA pairs (a GPC with an odd SM count leaves one SM out)
    if (!max_clusters) {
      cfg.gridDim = dim3(g_num_sms & ~1); cfg.dynamicSmemBytes = MAX_SMEM;
      CUDA_OK(cudaOccupancyMaxActiveClusters(&max_clusters, conv_tc_kernel<true, 2>, &cfg));
      BFSR_CHECK(max_clusters > 0, "conv_tc: no 2-CTA cluster can be scheduled");
      cfg.dynamicSmemBytes = smem;
    }
    const int pairs = a.total_tiles / 2;
    grid = 2 * (pairs < max_clusters ? pairs : max_clusters);
    cfg.gridDim = dim3(grid);
    snprintf(g_prof_tag, sizeof g_prof_tag, "tc%s-cl2 k%d %d->%d %dx%d", phase == 2 ? "-phase1p" : (fold == 2 ? "-f3" : ""), w.ks, w.cin, w.cout, out.H, out.W);
    a.pdl = 0;
    ProfScope prof(PK_CONV_TC, 2.0 * (double)out.npix() * w.cin * w.ks * w.ks * w.cout, s);
    CUDA_OK(cudaLaunchKernelEx(&cfg, conv_tc_kernel<true, 2>, a));
    count_launch();
    return;
  }
  // algorithmic FLOPs are those of the 3x3 conv over the upsampled tensor (what the reference computes)
  snprintf(g_prof_tag, sizeof g_prof_tag, "tc%s k%d %d->%d %dx%d%s", fold == 2 ? "-f3" : fold ? "-fold" : (phase == 2 ? "-phase1p" : (phase ? "-phase" : "")), w.ks, w.cin, w.cout, out.H, out.W,
           in_mode == IN_UP2 ? " up2" : "");
  ProfScope prof(PK_CONV_TC, 2.0 * (double)out.npix() * w.cin * w.ks * w.ks * w.cout, s);
  // Programmatic dependent launch is wired (griddepcontrol in the kernel) but OFF by default: measured on the whole step it costs
  // 3 % (batch 32: 300.3 vs 290.0 ms; batch 4: 43.7 vs 42.4 ms) -- the early-scheduled CTAs of the next conv take issue slots and
  // barrier traffic from the draining one without shortening its tail.  BFSR_PDL=1 enables it.
  // lean 8-warp instantiation (epilogue without the residual / fp32 pre-activation / second output / tap-fold / C = 24 FlowStep paths)
  static const bool lean8_on = getenv("BFSR_TC_LEAN8") && atoi(getenv("BFSR_TC_LEAN8")) == 1;
  const bool lean8 = lean8_on && !ew16 && a.tma && phase == 0 && fold != 1 && !epi.res1 && !epi.res2 && !epi.out2 && (!epi.pre || pre_gemm) &&
                     (!epi.flow || epi.flow->C == 12);
  static const int dbg_env = getenv("BFSR_TC_DBG") ? atoi(getenv("BFSR_TC_DBG")) : 0;
  a.dbg = dbg_env;
  // BFSR_PDL_SMALL=<pixels>: programmatic dependent launch only for launches with at most that many output pixels (the level-3 convs of
  // 80x80 tiles are dominated by launch + prologue latency; on the big convs PDL costs more than it hides)
  static const bool pdl_all = getenv("BFSR_PDL") && atoi(getenv("BFSR_PDL")) == 1;
  // Measured on the whole config-2 step: <= 1 M pixels (everything but the 320x320 launches) 257.6 -> 255.8 ms; all launches: 3 % slower.
  static const long long pdl_small = getenv("BFSR_PDL_SMALL") ? atoll(getenv("BFSR_PDL_SMALL")) : 1000000;
  const bool use_pdl = pdl_all || (pdl_small > 0 && out.npix() <= pdl_small);
  a.pdl = use_pdl ? 1 : 0;
  cudaLaunchConfig_t cfg = {};
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.gridDim = dim3(grid); cfg.dynamicSmemBytes = smem; cfg.stream = s; cfg.attrs = at; cfg.numAttrs = use_pdl ? 1 : 0;
  if (a.tma && ew16) {
    CUDA_OK(cudaFuncSetAttribute(conv_tc_kernel<true, 1, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, MAX_SMEM));
    cfg.blockDim = dim3(Roles<true, 16>::NTHREADS);
    CUDA_OK(cudaLaunchKernelEx(&cfg, conv_tc_kernel<true, 1, 16>, a));
  } else if (a.tma && lean8) {
    CUDA_OK(cudaFuncSetAttribute(conv_tc_kernel<true, 1, 8, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, MAX_SMEM));
    cfg.blockDim = dim3(Roles<true>::NTHREADS);
    CUDA_OK(cudaLaunchKernelEx(&cfg, conv_tc_kernel<true, 1, 8, true>, a));
  } else if (a.tma) {
    cfg.blockDim = dim3(Roles<true>::NTHREADS);
    CUDA_OK(cudaLaunchKernelEx(&cfg, conv_tc_kernel<true>, a));
  } else {
    cfg.blockDim = dim3(Roles<false>::NTHREADS);
    CUDA_OK(cudaLaunchKernelEx(&cfg, conv_tc_kernel<false>, a));
  }
  count_launch();
}

void conv2d_tc(const ConvW& w, const View& in, const View& out, const ConvEpi& epi, int in_mode, cudaStream_t s) {
  BFSR_CHECK(!w.tc_phase, "conv_tc: phase-packed weights need conv2d_tc_up2_phase");
  // Tap folding is correct but not a win on B200 as built (measured 64->12 @ 320x320: 0.99 ms folded vs 0.56 ms): the nine-tap
  // shift-add through shared memory costs what the 9x fewer MMAs save, and its 336-448 accumulator columns leave no second
  // TMEM stage to hide it.  Opt-in: BFSR_TC_FOLD=1 (or g_tc_fold = 1 from the per-op test entry).
  static const bool env_fold = getenv("BFSR_TC_FOLD") && atoi(getenv("BFSR_TC_FOLD")) == 1;
  const bool fold = (g_tc_fold == 1 || (g_tc_fold < 0 && env_fold)) && w.w_tc_fold && g_conv_mode == 0 && in.fmt == BF16X2 && in_mode == IN_DIRECT && !epi.pre && !epi.res1 &&
                    !epi.res2 && !epi.out2 && epi.alpha == 1.f && (epi.act == ACT_NONE || epi.act == ACT_CROSS_SIGMOID) &&
                    (epi.flow || (out.fmt == F32 && out.cs % 4 == 0 && out.coff % 4 == 0 && ((uintptr_t)out.p % 16) == 0));
  // dx folding (TcArgs::fold = 2): every TMA-fed 3x3 conv with <= 32 output channels and a plain epilogue -- the RRDB's dense convs,
  // the (shift, scale) heads of the couplings (with the fused FlowStep) and of fFeatures.  BFSR_TC_F3=0 keeps the per-tap MMAs.
  static const bool env_f3 = !(getenv("BFSR_TC_F3") && atoi(getenv("BFSR_TC_F3")) == 0);
  const bool f3 = !fold && (g_tc_fold == 2 || (g_tc_fold < 0 && env_f3)) && w.w_tc_f3 && g_conv_mode != 2 && bf_in_ok(in) && in_mode == IN_DIRECT &&
                  !epi.pre && !epi.res1 && !epi.res2 && !epi.out2 && epi.alpha == 1.f && (epi.flow || out_ok(out)) && out.W >= 16;
  launch_tc(w, in, out, epi, in_mode, 0, s, nullptr, f3 ? 2 : (fold ? 1 : 0));
}
void conv2d_tc_up2_phase(const ConvW& w, const View& in_lowres, const View& out, const ConvEpi& epi, cudaStream_t s) {
  BFSR_CHECK(w.tc_phase == 1, "conv_tc: weights are not phase-packed");
  launch_tc(w, in_lowres, out, epi, IN_DIRECT, 1, s);
}

void conv2d_tc_phase1(const ConvW& w, const View& in_hi, const View& in_lo, const View& out, const ConvEpi& epi, cudaStream_t s) {
  BFSR_CHECK(w.tc_phase == 2, "conv_tc: weights are not packed for the single-pass phase evaluation");
  launch_tc(w, in_lo, out, epi, IN_DIRECT, 2, s, &in_hi);
}

// Single-pass packing of a 3x3 conv over [hi (hi_cn channels at full resolution) | nearest2x(lo) (lo_cn channels)]
// (the level-1 conditioning tensor, SRFlowNet_arch.py:136).  Per (cout tile, output phase) the tap images are laid out in
// the order the kernel walks them: low-res chunks x 4 pre-summed taps, then per hi chunk the four parity planes with the
// original 3x3 taps that land on each plane (dy-major).  w_oihw: [cout][cin_src][3][3]; bias is the final bias.
ConvW pack_conv_tc_phase1(const float* w_oihw, int cout, int cin_src, int hi_c0, int hi_cn, int lo_c0, int lo_cn,
                          const float* out_scale, const float* bias) {
  using namespace tc;
  BFSR_CHECK(hi_cn % KC == 0 && lo_cn % KC == 0, "pack_conv_tc_phase1: channel counts must be multiples of %d", KC);
  ConvW c;
  c.ks = 3; c.cin = hi_cn + lo_cn; c.cout = cout; c.cin_pad = c.cin; c.co_tile = 64; c.cout_pad = (cout + 63) / 64 * 64;
  c.tc_phase = 2;
  const int nt = pick_nt(cout);
  const int n_tiles = (cout + nt - 1) / nt, n_lo = lo_cn / KC, n_hi = hi_cn / KC;
  const int taps_tile = n_lo * 4 + n_hi * 9;
  const size_t tap_elems = (size_t)2 * nt * (ROWB / 2);
  std::vector<unsigned short> img((size_t)n_tiles * 4 * taps_tile * tap_elems, 0);
  auto members = [](int f, int a, int* d) {   // taps of the 3-tap filter that land on low-res offset a for phase f
    if (f == 0) { if (a == 0) { d[0] = 0; return 1; } d[0] = 1; d[1] = 2; return 2; }
    if (a == 0) { d[0] = 0; d[1] = 1; return 2; } d[0] = 2; return 1;
  };
  auto put = [&](unsigned short* dst, int t, int ci0, int ny, const int* dys, int nx, const int* dxs) {
    for (int r = 0; r < nt; ++r) {
      const int co = t * nt + r;
      for (int k = 0; k < KC; ++k) {
        double acc = 0.0;
        if (co < cout)
          for (int iy = 0; iy < ny; ++iy)
            for (int ix = 0; ix < nx; ++ix)
              acc += (double)w_oihw[(((size_t)co * cin_src + ci0 + k) * 3 + dys[iy]) * 3 + dxs[ix]];
        const float w = (float)(acc * (out_scale && co < cout ? (double)out_scale[co] : 1.0));
        const unsigned short hi = f2bf(w), lo = f2bf(w - bf2f(hi));
        const int j = k >> 3, e = k & 7;
        dst[(size_t)r * 32 + ((j ^ ((r >> 1) & 3)) << 3) + e] = hi;
        const int r2 = nt + r;
        dst[(size_t)r2 * 32 + ((j ^ ((r2 >> 1) & 3)) << 3) + e] = lo;
      }
    }
  };
  for (int t = 0; t < n_tiles; ++t)
    for (int ph = 0; ph < 4; ++ph) {
      unsigned short* dst = img.data() + ((size_t)t * 4 + ph) * taps_tile * tap_elems;
      for (int ch = 0; ch < n_lo; ++ch)
        for (int tap = 0; tap < 4; ++tap, dst += tap_elems) {
          int dys[2], dxs[2];
          const int ny = members(ph >> 1, tap >> 1, dys), nx = members(ph & 1, tap & 1, dxs);
          put(dst, t, lo_c0 + ch * KC, ny, dys, nx, dxs);
        }
      for (int ch = 0; ch < n_hi; ++ch)
        for (int pl = 0; pl < 4; ++pl)
          for (int dy = 0; dy < 3; ++dy)
            for (int dx = 0; dx < 3; ++dx) {
              const int u = (ph >> 1) + dy - 1, v = (ph & 1) + dx - 1;
              if ((u & 1) != (pl >> 1) || (v & 1) != (pl & 1)) continue;
              put(dst, t, hi_c0 + ch * KC, 1, &dy, 1, &dx);
              dst += tap_elems;
            }
    }
  CUDA_OK(cudaMalloc(&c.w_tc, img.size() * 2));
  CUDA_OK(cudaMemcpy(c.w_tc, img.data(), img.size() * 2, cudaMemcpyHostToDevice));
  std::vector<float> hb(c.cout_pad + 128, 0.f);
  if (bias) for (int i = 0; i < cout; ++i) hb[i] = bias[i];
  CUDA_OK(cudaMalloc((void**)&c.bias, hb.size() * 4));
  CUDA_OK(cudaMemcpy(c.bias, hb.data(), hb.size() * 4, cudaMemcpyHostToDevice));
  c.tc_kchunks = n_lo + 4 * n_hi; c.tc_npad = nt;
  return c;
}

// Phase packing of a 3x3 conv applied to nearest2x(x) (SRFlowNet_arch.py:136 feeds such tensors to every level-1 coupling):
//   out(2y+fy, 2x+fx) = sum_{a,b in {0,1}} W'[fy,fx,a,b] . x(y+fy-1+a, x+fx-1+b),
//   W'[f=0]: a=0 <- {d=0}, a=1 <- {d=1,2};   W'[f=1]: a=0 <- {d=0,1}, a=1 <- {d=2}   (same rule along x).
// w_oihw: [cout][cin_src][3][3]; uses source channels [c0, c0+cn); out_scale multiplies per output channel.
ConvW pack_conv_tc_phase(const float* w_oihw, int cout, int cin_src, int c0, int cn, const float* out_scale) {
  using namespace tc;
  ConvW c;
  c.ks = 3; c.cin = cn; c.cout = cout; c.cin_pad = cn; c.co_tile = 64; c.cout_pad = (cout + 63) / 64 * 64;
  c.tc_phase = 1;
  const int nt = pick_nt(cout);
  const int n_tiles = (cout + nt - 1) / nt, n_chunks = (cn + KC - 1) / KC;
  const size_t tap_elems = (size_t)2 * nt * (ROWB / 2);
  std::vector<unsigned short> img((size_t)n_tiles * 4 * n_chunks * 4 * tap_elems, 0);
  auto members = [](int f, int a, int* d) {   // taps of the 3-tap filter that land on low-res offset a for phase f
    if (f == 0) { if (a == 0) { d[0] = 0; return 1; } d[0] = 1; d[1] = 2; return 2; }
    if (a == 0) { d[0] = 0; d[1] = 1; return 2; } d[0] = 2; return 1;
  };
  for (int t = 0; t < n_tiles; ++t)
    for (int ph = 0; ph < 4; ++ph)
      for (int ch = 0; ch < n_chunks; ++ch)
        for (int tap = 0; tap < 4; ++tap) {
          unsigned short* dst = img.data() + ((((size_t)t * 4 + ph) * n_chunks + ch) * 4 + tap) * tap_elems;
          int dys[2], dxs[2];
          const int ny = members(ph >> 1, tap >> 1, dys), nx = members(ph & 1, tap & 1, dxs);
          for (int r = 0; r < nt; ++r) {
            const int co = t * nt + r;
            for (int k = 0; k < KC; ++k) {
              const int ci = ch * KC + k;
              double acc = 0.0;
              if (co < cout && ci < cn)
                for (int iy = 0; iy < ny; ++iy)
                  for (int ix = 0; ix < nx; ++ix)
                    acc += (double)w_oihw[(((size_t)co * cin_src + c0 + ci) * 3 + dys[iy]) * 3 + dxs[ix]];
              const float w = (float)(acc * (out_scale && co < cout ? (double)out_scale[co] : 1.0));
              const unsigned short hi = f2bf(w), lo = f2bf(w - bf2f(hi));
              const int j = k >> 3, e = k & 7;
              dst[(size_t)r * 32 + ((j ^ ((r >> 1) & 3)) << 3) + e] = hi;
              const int r2 = nt + r;
              dst[(size_t)r2 * 32 + ((j ^ ((r2 >> 1) & 3)) << 3) + e] = lo;
            }
          }
        }
  CUDA_OK(cudaMalloc(&c.w_tc, img.size() * 2));
  CUDA_OK(cudaMemcpy(c.w_tc, img.data(), img.size() * 2, cudaMemcpyHostToDevice));
  std::vector<float> zb(c.cout_pad + 128, 0.f);
  CUDA_OK(cudaMalloc((void**)&c.bias, zb.size() * 4));
  CUDA_OK(cudaMemcpy(c.bias, zb.data(), zb.size() * 4, cudaMemcpyHostToDevice));
  c.tc_kchunks = n_chunks; c.tc_npad = nt;
  return c;
}

}  // namespace bfsr
