// One kernel per coupling FlowStep of a C = 12 or C = 24 level (SRFlow levels 1 and 2): the z-dependent affine sub-net
//     fAffine.0 (3x3, z1 -> 64, + feature-only pre-activation) -> ReLU -> fAffine.2 (1x1, 64 -> 64) -> ReLU -> fAffine.4 (3x3, 64 -> C,
//     cross-sigmoid) -> affine coupling + the FlowStep's ActNorm / InvConv1x1 / feature-affine
// (FlowAffineCouplingsAblation.py:78-96, 114-135; FlowStep.py:88-129) with the two 64-channel hidden maps kept in TENSOR MEMORY, and -- TAIL
// variant -- the feature-only tail fFeatures.2 (1x1) -> ReLU -> fFeatures.4 (3x3, cross-sigmoid) = hF through the same machinery.
// The three-launch chain it replaces (conv_tc: z-conv, 1x1, dx-folded head + FlowEpi) moved 1.3 KB per level pixel through HBM for
// ~450 B of compulsory traffic; here the only HBM traffic is the pre-activation (256 B), z / hF (48 + 48 + 96 B) and the z1 operand (C = 12).
//
// Work decomposition: a work item is a vertical STRIP of 28 output columns x seg_rows output rows of one image.  The CTA marches down
// the strip in BLOCKS of 128 raster positions = 4 rows x 32 columns (one UMMA M = 128 tile); column c of the raster is image column
// x0 - 1 + c, so the 30 columns the head conv needs (28 outputs + 1 halo each side) are columns 0..29 and columns 30, 31 are padding.
// Per block b (h rows yb .. yb+3), C = 12 numbers:
//   M1: acc1 = pre-activation (identity MMAs over the TMA-loaded BF16X2 tile) + conv3x3(z1): nine shifted views of the 6 x 32 z1 halo tile;
//       z1 arrives as ONE bf16 plane [hi(8) | lo(8)] per pixel, so  z_hi.W_hi + z_lo.W_hi  is one K = 16 MMA and  z_hi.W_lo  a second one
//   E1: h1 = relu(acc1) -> packed bf16 (hi, lo) written back IN PLACE into the accumulator's own tensor-memory columns (tcgen05.st): the A
//       operand of the next GEMM; the hidden maps never touch shared memory.  (First version: operand tiles in shared memory -- 4550 clk per
//       block, bound by shared-memory bandwidth: every SS-mode MMA re-reads its 4 KB A slice, 440 KB per block; with A in TMEM an MMA only
//       reads its weights.)
//   M2: acc2 = h1 . W2 (split-bf16 x3, A from TMEM)   E2: h2 = relu(acc2 + b2), zeroed outside the image (the head conv's zero padding), in place
//   M3: acc3[r, tap*12 + co] = h2[r, :] . W3[tap][:, co]  -- all nine taps folded into N = 112 (C = 24: 224), no halo of h2 needed anywhere
//   E3: out(y, x) = sum_{dy,dx} acc3[(y-1+dy, x-1+dx), tap]: the dx sum by warp shuffles (a warp holds one raster row); the dy = 0 and dy = 1
//       tap-row sums of every raster row go to a ring in shared memory, and the warp holding row y finalises output row y - 1 from the
//       rings (rows y-2, y-1) and its own dy = 2 sums -- rows are finalised with a lag of one row and carried across blocks; then bias,
//       cross-sigmoid, and the FlowStep epilogue per pixel (z, hF in; z, z1 operand out) or, TAIL, the store of hF.
// Two blocks are in flight; issuer A runs M1 up to two blocks ahead, issuer B interleaves M2(b), M3(b-1); E1/E2 run on 8 warps, E3 on three groups
// of 4 warps that take blocks in rotation.  Arithmetic is that of the three-launch chain (same split-bf16 products, fp32 accumulation; in the
// bf16 single-pass mode only the hi x hi products).  Measurements and the path to this form: profiles/r2_coupling_fused_summary.md.
#include "ops.cuh"
#include "tc_ptx.cuh"
#include <vector>
#include <cstdlib>
#include <cmath>

#ifdef BFSR_TC_TRACE
#define TR_DECL(...) long long __VA_ARGS__
#define TR_T(x) const long long x = clock64()
#define TR_ADD(acc, t0) acc += clock64() - (t0)
#else
#define TR_DECL(...)
#define TR_T(x)
#define TR_ADD(acc, t0)
#endif

namespace bfsr {
namespace cf {
constexpr int ROWB = 64;
constexpr int PLANE = 128 * ROWB;                      // one (chunk, plane) operand tile: 128 rows x 64 B
constexpr int Z1_ROWS = 6 * 32;
constexpr int Z1_BYTES = Z1_ROWS * ROWB;               // 12288
constexpr int PRE_BYTES = 4 * PLANE;                   // [chunk][hi, lo]
constexpr int W2_BYTES = 2 * 128 * ROWB;               // per 32-channel chunk [W_hi (64 rows) ; W_lo (64 rows)]
constexpr int ID_BYTES = 32 * ROWB;                    // 32 x 32 identity (pre-activation as K chunks)
constexpr int NG = 3;                                  // E3 groups (4 warps each)
constexpr int W_ISSUE_A = 8 + 4 * NG, W_ISSUE_B = W_ISSUE_A + 1, W_LOAD = W_ISSUE_B + 1;
constexpr int NTHREADS = (W_LOAD + 1) * 32;            // 8 E1/E2 warps, 12 E3 warps, two MMA issuers, loader
// tensor memory: fp32 accumulators of the three GEMMs.  E1 / E2 convert an accumulator IN PLACE into the packed bf16
// A operand of the next GEMM: the 16 fp32 columns of a 16-channel group become 8 columns of (hi, hi) pairs + 8 columns of (lo, lo) pairs
constexpr int TM_ACC1 = 0, TM_ACC2 = 128, TM_ACC3 = 256, TM_COLS = 512;
constexpr int OUT_W = 28;                              // output columns per strip
// mbarrier indices
enum { B_WFULL = 0, B_INFULL = 1, B_PREFULL = 3, B_PREEMPTY = 5, B_ACC1FULL = 6, B_H1READY = 8, B_ACC2FULL = 10, B_H2READY = 12,
       B_ACC3FULL = 14 /* one per (E3 group, head group): a waiter must see every phase of a barrier it polls */,
       B_ACC3EMPTY = B_ACC3FULL + 2 * NG, B_BARW = B_ACC3EMPTY + 2 /* tap-row sums of a block written */,
       B_BARR = B_BARW + NG /* ... and read (C = 24: the small ring is recycled block by block) */, B_COUNT = B_BARR + NG };
// Per level: C = 12 (z1: 6 channels -> one K = 16 step [hi(8) | lo(8)], head N = 9 x 12 -> 112, two accumulator / pre-activation stages,
// 16-row tap-sum ring) or C = 24 (z1: 12 channels -> K = 32 [hi(16) | lo(16)], head N = 9 x 24 -> 224 in ONE accumulator stage, one
// pre-activation stage and a 6-row ring recycled block by block: its 130 KB of weights leave no more shared memory)
// TAIL: the feature-only tail of a coupling (fFeatures.2 1x1 -> ReLU -> fFeatures.4 3x3 -> cross-sigmoid = the (shiftF, scaleF) pairs hF of
// a level's step) runs through the same machinery without M1 / E1 and without the FlowStep: C = the tail's 2 x C_level output channels.
template <int C, bool TAIL = false> struct Cfg {
  static_assert(C == 12 || C == 24, "coupling_fused: C = 12 or 24");
  static_assert(!TAIL || C == 24, "coupling_fused: the tail variant is built for 24 output channels (C = 12 levels)");
  static constexpr int ZP = C == 12 ? 8 : 16;                      // z1 channels, padded
  static constexpr int NSUB = TAIL ? 2 : 1;                        // head MMA groups per block (TAIL: two halves of 12 output channels, each with its
                                                                   // own accumulator slot, released as soon as E3 has summed it: the next block's head
                                                                   // MMAs run under the other half's sums)
  static constexpr int CH = C / NSUB;                              // output channels per head group / accumulator slot
  static constexpr int N3 = CH == 12 ? 112 : 224;                  // 9 taps x CH columns (+ padding to a multiple of 16)
  // W1: C = 12: per tap 64 rows, bytes 0..31 = [W_hi | W_hi], bytes 32..63 = [W_lo | 0]
  //     C = 24: per tap 64 rows [W_hi(16) | W_hi(16)], then five images holding [W_lo(16)] of two taps per row (bytes 0..31 / 32..63)
  static constexpr int W1_BYTES = TAIL ? 0 : (C == 12 ? 9 * 64 * ROWB : (9 + 5) * 64 * ROWB);
  static constexpr int W3_BYTES = 2 * NSUB * 2 * N3 * ROWB;        // per 32-channel chunk and head group [W_hi (N3 rows) ; W_lo (N3 rows)]
  static constexpr int W_BYTES = W1_BYTES + W2_BYTES + W3_BYTES + ID_BYTES;
  static constexpr int NPRE = (C == 12 || TAIL) ? 2 : 1;           // pre-activation (TAIL: input) stages
  static constexpr int NACC3 = (C == 12 || TAIL) ? 2 : 1;          // head accumulator slots of N3 columns (C = 12: block parity; TAIL: head half)
  static constexpr int RING = C == 12 ? 16 : (TAIL ? 14 : 6);      // (14: the smallest ring whose aliasing blocks are three apart = same E3 group)                    // rows of tap-row sums kept for the neighbouring rows (two arrays: dy = 0, 1)
  static constexpr bool CHAIN = C != 12 && !TAIL;                           // ring too small to run ahead: block b writes after block b-1 has read
  static constexpr int EXCH_BYTES = 2 * RING * C * 32 * 4;         // [dy][row % RING][C channels][32 lanes] fp32
  static constexpr int OFF_Z1 = W_BYTES, OFF_PRE = OFF_Z1 + (TAIL ? 0 : 2 * Z1_BYTES), OFF_EXCH = OFF_PRE + NPRE * PRE_BYTES, OFF_BARS = OFF_EXCH + EXCH_BYTES;
  static constexpr int SMEM_BYTES = OFF_BARS + 512 + 1024;         // + alignment slack
  static_assert(SMEM_BYTES <= 227 * 1024, "coupling_fused: shared memory budget");
  static_assert(W_BYTES % 1024 == 0 && OFF_PRE % 1024 == 0, "operand tiles must stay 1024-byte aligned");
  static_assert(TM_ACC3 + NACC3 * N3 <= TM_COLS, "coupling_fused: tensor memory budget");
};
}  // namespace cf

struct CfArgs {
  alignas(64) CUtensorMap tm_z1;     // (2 ZP ch, W, H, N, 1) bf16, box = 32 ch (zero-filled past 2 ZP) x 32 px x 6 rows
  alignas(64) CUtensorMap tm_pre;    // (C, W, H, N, plane) bf16 BF16X2 view, box = 32 ch x 32 px x 4 rows
  const unsigned char* w;
  int pre_coff;
  int H, W, N;
  int strips, segs, seg_rows, nblk, total_items;
  float eps;
  int fast;                          // bf16 single-pass mode (precision 1): only the hi x hi products; z1 / h1 / h2 carry zero lo halves
  int dbg;                           // timing experiments only (BFSR_CF_DBG bit mask: skip 1 = conv taps, 2 = identity, 4 = M2, 8 = M3 MMAs, 16 = flow epilogue): WRONG results
  int inv, has_mix, has_hF;
  View z_in, z_out, hF;
  __nv_bfloat16* z1_out;             // [npix][2 ZP] = [hi(ZP) | lo(ZP)] of the first C/2 output channels (next step's conv operand) or null
  float bias2[64], bias3[32];        // biases of fAffine.2 and fAffine.4 (fAffine.0's rides in the pre-activation: its z part has none)
  float M[576], cvec[24];            // [C][C] row-major, [C]
};

// wait of an epilogue role: a short fixed sleep between polls -- 20 warps spinning on try_wait take the issue slots the working warps need
// (the kernel is close to instruction-issue bound), while the 64..512 ns back-off of mbar_wait_relaxed is too coarse for the hand-offs
// on the M1 -> E1 -> M2 -> E2 -> M3 chain
__device__ __forceinline__ void mbar_wait_nap(uint32_t bar, uint32_t parity) {
  if (mbar_try(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try(bar, parity)) {
    __nanosleep(32);
    if (clock64() - t0 > 8000000000LL) {
      printf("bfsr coupling_fused: mbarrier timeout (block %d thread %d bar %u)\n", blockIdx.x, threadIdx.x, bar);
      __trap();
    }
  }
}

struct Blk { int n, x0, yb, y0, y1; };   // yb = image row of raster row 0 of the block; [y0, y1) = output rows of the item
__device__ __forceinline__ Blk blk_coord(const CfArgs& a, int b) {
  const int ii = b / a.nblk, jb = b - ii * a.nblk;
  int item = (int)blockIdx.x + ii * (int)gridDim.x;
  const int seg = item % a.segs; item /= a.segs;
  const int sx = item % a.strips;
  Blk k;
  k.n = item / a.strips;
  k.x0 = sx * cf::OUT_W;
  k.y0 = seg * a.seg_rows;
  k.y1 = min(k.y0 + a.seg_rows, a.H);
  k.yb = k.y0 - 1 + 4 * jb;
  return k;
}

// Walks the CTA's block stream without per-block divisions (each role steps through it in order; the three runtime divisions of
// blk_coord were ~150 instructions per block per warp in a kernel that is close to instruction-issue bound)
struct BlkIter {
  int ii = 0, jb = 0; Blk k;
  __device__ __forceinline__ void init(const CfArgs& a, int b) { ii = b / a.nblk; jb = b - ii * a.nblk; k = blk_coord(a, b); }
  __device__ __forceinline__ void advance(const CfArgs& a, int n) {
    jb += n; k.yb += 4 * n;
    if (jb >= a.nblk) { while (jb >= a.nblk) { jb -= a.nblk; ++ii; } k = blk_coord(a, ii * a.nblk + jb); }
  }
};

// 32 channels of one accumulator row -> packed bf16 (hi, lo) words: word i = channels (2i, 2i+1), even channel in the low half
__device__ __forceinline__ void split_pack32(const float* o, uint32_t* hi, uint32_t* lo) {
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const float x0 = o[2 * i], x1 = o[2 * i + 1];
    hi[i] = pack_bf16(x0, x1);
    lo[i] = pack_bf16(x0 - __uint_as_float(hi[i] << 16), x1 - __uint_as_float(hi[i] & 0xffff0000u));
  }
}
__device__ __forceinline__ void tmem_st8(uint32_t addr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
               ::"r"(addr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t addr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
               ::"r"(addr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
                 "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t addr, float* v) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(addr));
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld4(uint32_t addr, float* v) {
  uint32_t r[4];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
#pragma unroll
  for (int i = 0; i < 4; ++i) v[i] = __uint_as_float(r[i]);
}
// A operand from tensor memory (128 lanes x 8 columns of packed bf16 pairs per K = 16 step), B from shared memory
__device__ __forceinline__ void umma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
               ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(acc) : "memory");
}

// FlowStep on one pixel: h = (shift, scale) pairs of the coupling (FlowEpi semantics, ops.cuh); matrices from the kernel parameters.
// zq: the pixel's z, prefetched by the caller; hq: the first NHQ of its C/2 hF quads, prefetched too (the rest is loaded here: register budget).
template <int C, int NHQ>
__device__ __forceinline__ void flow_apply(const CfArgs& a, const float* h, long long pix, const float4* zq, const float4* hq) {
  float z[C], o[C];
  const float4* fp = reinterpret_cast<const float4*>((const float*)a.hF.p + pix * a.hF.cs + a.hF.coff);
#pragma unroll
  for (int k = 0; k < C / 4; ++k) { const float4 v = zq[k]; z[4 * k] = v.x; z[4 * k + 1] = v.y; z[4 * k + 2] = v.z; z[4 * k + 3] = v.w; }
#pragma unroll
  for (int j = 0; j < C / 2; ++j) {
    if (a.inv) z[C / 2 + j] = __fdividef(z[C / 2 + j], h[2 * j + 1]) - h[2 * j];
    else z[C / 2 + j] = (z[C / 2 + j] + h[2 * j]) * h[2 * j + 1];
  }
  if (a.inv && a.has_hF) {
#pragma unroll
    for (int k = 0; k < C / 2; ++k) {
      const float4 v = k < NHQ ? hq[k] : __ldg(fp + k);
      z[2 * k] = __fdividef(z[2 * k], v.y) - v.x; z[2 * k + 1] = __fdividef(z[2 * k + 1], v.w) - v.z;
    }
  }
  if (a.has_mix) {
#pragma unroll
    for (int co = 0; co < C; ++co) {                      // packed fp32x2 FMAs (sm_100 FFMA2): even / odd k partial sums
      float2 acc = make_float2(0.f, 0.f);
#pragma unroll
      for (int k = 0; k < C; k += 2) acc = __ffma2_rn(make_float2(a.M[co * C + k], a.M[co * C + k + 1]), make_float2(z[k], z[k + 1]), acc);
      o[co] = a.inv ? (acc.x + acc.y) - a.cvec[co] : (acc.x + acc.y) + a.cvec[co];
    }
    if (!a.inv && a.has_hF) {
#pragma unroll
      for (int k = 0; k < C / 2; ++k) {
        const float4 v = k < NHQ ? hq[k] : __ldg(fp + k);
        o[2 * k] = (o[2 * k] + v.x) * v.y; o[2 * k + 1] = (o[2 * k + 1] + v.z) * v.w;
      }
    }
  } else {
#pragma unroll
    for (int c = 0; c < C; ++c) o[c] = z[c];
  }
  float4* dst = reinterpret_cast<float4*>((float*)a.z_out.p + pix * a.z_out.cs + a.z_out.coff);
#pragma unroll
  for (int k = 0; k < C / 4; ++k) dst[k] = make_float4(o[4 * k], o[4 * k + 1], o[4 * k + 2], o[4 * k + 3]);
  if (a.z1_out) {
    constexpr int ZP = cf::Cfg<C>::ZP;
    uint32_t hi[ZP / 2], lo[ZP / 2];
#pragma unroll
    for (int e = 0; e < ZP / 2; ++e) {
      const float x0 = 2 * e < C / 2 ? o[2 * e] : 0.f, x1 = 2 * e + 1 < C / 2 ? o[2 * e + 1] : 0.f;
      hi[e] = pack_bf16(x0, x1);
      lo[e] = a.fast ? 0u : pack_bf16(x0 - __uint_as_float(hi[e] << 16), x1 - __uint_as_float(hi[e] & 0xffff0000u));
    }
    uint4* d = reinterpret_cast<uint4*>(a.z1_out + pix * (2 * ZP));
#pragma unroll
    for (int e = 0; e < ZP / 8; ++e) {
      d[e] = make_uint4(hi[4 * e], hi[4 * e + 1], hi[4 * e + 2], hi[4 * e + 3]);
      d[ZP / 8 + e] = make_uint4(lo[4 * e], lo[4 * e + 1], lo[4 * e + 2], lo[4 * e + 3]);
    }
  }
}

template <int C, bool TAIL>
__global__ void __launch_bounds__(cf::NTHREADS, 1) coupling_fused_kernel(const __grid_constant__ CfArgs a) {
  using namespace cf;
  using K = Cfg<C, TAIL>;
  constexpr int N3 = K::N3;
  extern __shared__ unsigned char smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  unsigned char* sgen = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t w1 = base, w2 = w1 + K::W1_BYTES, w3 = w2 + W2_BYTES, wid = w3 + K::W3_BYTES;
  const uint32_t z1s = base + K::OFF_Z1, pres = base + K::OFF_PRE, bars = base + K::OFF_BARS;
  auto bar = [&](int i) -> uint32_t { return bars + 8u * (uint32_t)i; };
  const uint32_t tmem_slot = bars + 8u * B_COUNT;

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int n_mine = ((int)a.total_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int NB = n_mine * a.nblk;                      // blocks this CTA streams through

  if (tid == 0) {
    for (int i = 0; i < B_COUNT; ++i) {
      uint32_t cnt = 1;
      if (i == B_H1READY || i == B_H1READY + 1 || i == B_H2READY || i == B_H2READY + 1) cnt = 256;
      if (i == B_ACC3EMPTY || i == B_ACC3EMPTY + 1 || (i >= B_BARW && i < B_BARR + NG)) cnt = 128;
      mbar_init(bar(i), cnt);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == W_ISSUE_A) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"((uint32_t)TM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  tmem_base = __shfl_sync(0xffffffffu, tmem_base, 0);

  if (warp == W_LOAD) {
    // ===================== loader: resident weights once, then per block the z1 halo tile and the pre-activation tile =====================
    if (elect_one()) {
      mbar_expect_tx(bar(B_WFULL), K::W_BYTES);
      if (!TAIL) bulk_g2s(w1, a.w, K::W1_BYTES, bar(B_WFULL));
      bulk_g2s(w2, a.w + K::W1_BYTES, W2_BYTES, bar(B_WFULL));
      bulk_g2s(w3, a.w + K::W1_BYTES + W2_BYTES, K::W3_BYTES, bar(B_WFULL));
      bulk_g2s(wid, a.w + K::W1_BYTES + W2_BYTES + K::W3_BYTES, ID_BYTES, bar(B_WFULL));
    }
    __syncwarp();
    TR_DECL(tr_in = 0); TR_T(tr_start);
    BlkIter bi; bi.init(a, 0);
    for (int b = 0; b < NB; ++b, bi.advance(a, 1)) {
      const Blk& k = bi.k;
      const int p = b & 1, j = b >> 1;
      TR_T(tr0);
      // M1 (TAIL: the 1x1) of the block that used stage p two blocks ago has retired
      mbar_wait_relaxed(bar((TAIL ? B_ACC2FULL : B_ACC1FULL) + p), (uint32_t)((j & 1) ^ 1));
      TR_ADD(tr_in, tr0);
      if (!TAIL && elect_one()) {
        mbar_expect_tx(bar(B_INFULL + p), Z1_BYTES);
        tma_load_5d(z1s + p * Z1_BYTES, &a.tm_z1, bar(B_INFULL + p), 0, k.x0 - 2, k.yb - 1, k.n, 0);
      }
      __syncwarp();
      const int ps = b % K::NPRE;
      if (K::NPRE == 1) mbar_wait_relaxed(bar(B_PREEMPTY), (uint32_t)((b & 1) ^ 1));   // single stage: the identity MMAs of block b - 1 have retired
      if (elect_one()) {
        mbar_expect_tx(bar(B_PREFULL + ps), a.fast ? PRE_BYTES / 2 : PRE_BYTES);
#pragma unroll
        for (int c = 0; c < 2; ++c)
#pragma unroll
          for (int pl = 0; pl < 2; ++pl)
            if (!(a.fast && pl)) tma_load_5d(pres + ps * PRE_BYTES + (c * 2 + pl) * PLANE, &a.tm_pre, bar(B_PREFULL + ps), a.pre_coff + 32 * c, k.x0 - 1, k.yb, k.n, pl);
      }
      __syncwarp();
    }
#ifdef BFSR_TC_TRACE
    if (blockIdx.x == 0 && lane == 0) printf("[cf trace] loader: total %lld wait stage free %lld (blocks %d)\n", clock64() - tr_start, tr_in, NB);
#endif
  } else if (warp == W_ISSUE_A) {
    // ===================== MMA issuer A: M1(b) = pre-activation + conv3x3(z1) -> acc1[p], up to two blocks ahead of the epilogues ==========
    // (one thread sustains a tcgen05.mma per ~40-50 clk and every barrier poll / commit costs it ~100 clk more: with a single issuer the
    // 50 MMAs + 10 barrier operations of a block took 3700 clk and bounded the kernel)
    const uint64_t dsc = make_desc(0, 8 * ROWB);        // K-major SWIZZLE_64B, 8-row groups 512 B apart (dense rows)
    auto D = [&](uint32_t addr) -> uint64_t { return dsc | (uint64_t)((addr & 0x3FFFF) >> 4); };
    const uint32_t id32 = make_idesc(32), id64 = make_idesc(64);
    mbar_wait(bar(B_WFULL), 0);
    TR_DECL(tr_in = 0, tr_h1 = 0); TR_T(tr_start);
    for (int b = 0; TAIL && b < NB; ++b) {             // TAIL: this issuer runs the 1x1 straight from the TMA-loaded input tile (SS mode)
      const int p = b & 1, j = b >> 1;
      mbar_wait(bar(B_PREFULL + p), (uint32_t)(j & 1));
      // acc2[p] still holds h2 of block b - 2 until its head MMAs (both halves; the second is committed last) have retired
      if (b >= 2) mbar_wait(bar(B_ACC3FULL + 2 * ((b - 2) % NG) + K::NSUB - 1), (uint32_t)(((b - 2) / NG) & 1));
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (elect_one()) {
        const uint32_t acc = tmem_base + TM_ACC2 + 64 * p;
#pragma unroll 1
        for (int ck = 0; ck < 4; ++ck) {
          const int c = ck >> 1, ks = ck & 1;
          const uint64_t Bh = D(w2 + c * 128 * ROWB + ks * 32), Bl = Bh + ((64 * ROWB) >> 4);
          const uint64_t Sh = D(pres + p * PRE_BYTES + c * 2 * PLANE + ks * 32), Sl = Sh + (PLANE >> 4);
          umma_f16(acc, Sh, Bh, id64, ck ? 1u : 0u);
          if ((a.dbg & 4) || a.fast) continue;
          umma_f16(acc, Sh, Bl, id64, 1u);
          umma_f16(acc, Sl, Bh, id64, 1u);
        }
        umma_commit(bar(B_ACC2FULL + p));                 // releases E2 and, two blocks later, the loader's input stage p
      }
      __syncwarp();
    }
    for (int b = 0; b < (TAIL ? 0 : NB); ++b) {
      const int p = b & 1, j = b >> 1;
      const uint32_t acc = tmem_base + TM_ACC1 + 64 * p;
      const int ps = b % K::NPRE;
      TR_T(tr1);
      mbar_wait(bar(B_ACC2FULL + p), (uint32_t)((j & 1) ^ 1));            // M2 of block b - 2 has consumed h1, which lives in the columns of acc1[p]
      TR_ADD(tr_h1, tr1); TR_T(tr0);
      mbar_wait(bar(B_PREFULL + ps), (uint32_t)((b / K::NPRE) & 1));
      TR_ADD(tr_in, tr0);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      // (the issue loops are rolled: the kernel's hot code must stay inside the instruction caches -- the fully unrolled first version
      // spent half of its issue slots of every role waiting for instruction fetches)
      if (elect_one()) {
#pragma unroll 1
        for (int c = 0; c < 2; ++c) {
          const uint32_t pt = pres + ps * PRE_BYTES + c * 2 * PLANE;
#pragma unroll
          for (int pl = 0; pl < 2; ++pl)
#pragma unroll
            for (int ks = 0; ks < 2; ++ks)
              if ((!(a.dbg & 2) || (c == 0 && (pl | ks) == 0)) && !(a.fast && pl)) umma_f16(acc + 32 * c, D(pt + pl * PLANE + ks * 32), D(wid + ks * 32), id32, (pl | ks) ? 1u : 0u);
        }
        if (K::NPRE == 1) umma_commit(bar(B_PREEMPTY));
      }
      __syncwarp();
      mbar_wait(bar(B_INFULL + p), (uint32_t)(j & 1));
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (elect_one()) {
        const uint32_t zt = z1s + p * Z1_BYTES;
#pragma unroll 1
        for (int dy = 0; dy < 3; ++dy) {
          if (a.dbg & 1) continue;
#pragma unroll
          for (int dx = 0; dx < 3; ++dx) {
            const int tap = dy * 3 + dx;
            const uint64_t A = D(zt + (uint32_t)(dy * 32 + dx) * ROWB), B1 = D(w1 + (uint32_t)tap * 64 * ROWB);
            if (C == 12) {
              umma_f16(acc, A, B1, id64, 1u);                // [z_hi | z_lo] . [W_hi | W_hi]   (fast mode: z_lo = 0)
              if (!a.fast) umma_f16(acc, A, B1 + 2, id64, 1u);   // + 32 bytes: [z_hi | z_lo] . [W_lo | 0]
            } else {
              umma_f16(acc, A, B1, id64, 1u);                // z_hi . W_hi
              if (!a.fast) {
                umma_f16(acc, A + 2, B1 + 2, id64, 1u);      // z_lo . W_hi
                umma_f16(acc, A, D(w1 + (uint32_t)(9 + (tap >> 1)) * 64 * ROWB) + 2 * (tap & 1), id64, 1u);   // z_hi . W_lo
              }
            }
          }
        }
        umma_commit(bar(B_ACC1FULL + p));                 // releases E1 and, two blocks later, the loader's z1 stage p
      }
      __syncwarp();
    }
#ifdef BFSR_TC_TRACE
    if (blockIdx.x == 0 && lane == 0) printf("[cf trace] issuer A: total %lld wait in_full %lld acc2_full(b-2) %lld (blocks %d)\n", clock64() - tr_start, tr_in, tr_h1, NB);
#endif
  } else if (warp == W_ISSUE_B) {
    // ===================== MMA issuer B: M2(it) = 1x1, then M3(it-1) = tap-folded head (E2 of block it-1 runs under M3(it-2) / M2(it)),
    // both with the A operand in tensor memory =====================
    const uint64_t dsc = make_desc(0, 8 * ROWB);
    auto D = [&](uint32_t addr) -> uint64_t { return dsc | (uint64_t)((addr & 0x3FFFF) >> 4); };
    const uint32_t id64 = make_idesc(64), idn3 = make_idesc(N3);
    mbar_wait(bar(B_WFULL), 0);
    TR_DECL(tr_h2 = 0, tr_a3 = 0, tr_h1 = 0); TR_T(tr_start);
    for (int it = 0; it <= NB; ++it) {
      if (it < NB && !TAIL) {                             // ---- M2(b): 1x1, acc2[p]
        const int b = it, p = b & 1, j = b >> 1;
        TR_T(tr4);
        mbar_wait(bar(B_H1READY + p), (uint32_t)(j & 1));
        TR_ADD(tr_h1, tr4);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (elect_one()) {
          const uint32_t acc = tmem_base + TM_ACC2 + 64 * p, At = tmem_base + TM_ACC1 + 64 * p;   // h1 sits where acc1 was
#pragma unroll 1
          for (int ck = 0; ck < 4; ++ck) {
            const int c = ck >> 1, ks = ck & 1;
            const uint64_t Bh = D(w2 + c * 128 * ROWB + ks * 32), Bl = Bh + ((64 * ROWB) >> 4);
            const uint32_t Ah = At + 16 * ck, Al = Ah + 8;          // 16-channel group: 8 columns of hi pairs, then 8 of lo pairs
            umma_f16_ts(acc, Ah, Bh, id64, ck ? 1u : 0u);
            if ((a.dbg & 4) || a.fast) continue;
            umma_f16_ts(acc, Ah, Bl, id64, 1u);
            umma_f16_ts(acc, Al, Bh, id64, 1u);
          }
          umma_commit(bar(B_ACC2FULL + p));
        }
        __syncwarp();
      }
      if (it >= 1) {                                      // ---- M3(b): acc3[r, tap*12+co]
        const int b = it - 1, p = b & 1, j = b >> 1;
        TR_T(tr2);
        mbar_wait(bar(B_H2READY + p), (uint32_t)(j & 1));
        TR_ADD(tr_h2, tr2);
#pragma unroll 1
        for (int hc = 0; hc < K::NSUB; ++hc) {
          const int slot = TAIL ? hc : (C == 12 ? p : 0);   // accumulator slot: head half (TAIL), block parity (C = 12), the only one (C = 24)
          TR_T(tr3);
          mbar_wait(bar(B_ACC3EMPTY + slot), (uint32_t)(((C == 12 && !TAIL ? j : b) & 1) ^ 1));
          TR_ADD(tr_a3, tr3);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          if (elect_one()) {
            const uint32_t acc = tmem_base + TM_ACC3 + N3 * slot, At = tmem_base + TM_ACC2 + 64 * p;    // h2 sits where acc2 was
#pragma unroll 1
            for (int ck = 0; ck < 4; ++ck) {
              const int c = ck >> 1, ks = ck & 1;
              const uint32_t Ah = At + 16 * ck, Al = Ah + 8;
              const uint64_t Bh = D(w3 + (c * K::NSUB + hc) * 2 * N3 * ROWB + ks * 32), Bl = Bh + ((N3 * ROWB) >> 4);
              umma_f16_ts(acc, Ah, Bh, idn3, ck ? 1u : 0u);
              if ((a.dbg & 8) || a.fast) continue;
              umma_f16_ts(acc, Ah, Bl, idn3, 1u);
              umma_f16_ts(acc, Al, Bh, idn3, 1u);
            }
            umma_commit(bar(B_ACC3FULL + 2 * (b % NG) + hc));
          }
          __syncwarp();
        }
      }
    }
#ifdef BFSR_TC_TRACE
    if (blockIdx.x == 0 && lane == 0) printf("[cf trace] issuer B: total %lld wait h2_ready %lld acc3_empty %lld h1_ready %lld (blocks %d)\n", clock64() - tr_start, tr_h2, tr_a3, tr_h1, NB);
#endif
  } else if (warp < 8) {
    // ===================== E1 / E2: TMEM -> bias, ReLU -> packed (hi, lo) A operand of the next GEMM, back into TMEM =====================
    const int q = warp & 3, half = warp >> 2;             // TMEM lane quarter = raster row of the block; 32-channel half
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
    TR_DECL(tr_a1 = 0, tr_a2 = 0); TR_T(tr_start);
    BlkIter bi; bi.init(a, 0);                          // block it - 1 (E2's coordinates)
    for (int it = 0; it <= NB; ++it) {
      if (it >= 2) bi.advance(a, 1);
#pragma unroll 1
      for (int ph = 0; ph < 2; ++ph) {                    // ph 0: E1 of block it, ph 1: E2 of block it - 1 (one code path: instruction-cache footprint)
        const int b = it - ph;
        if (b < 0 || b >= NB || (TAIL && ph == 0)) continue;
        const int p = b & 1, j = b >> 1;
        bool inside = true;
        if (ph == 0) {
          TR_T(tr0);
          mbar_wait_nap(bar(B_ACC1FULL + p), (uint32_t)(j & 1));
          TR_ADD(tr_a1, tr0);
        } else {
          const Blk& k = bi.k;
          const int y = k.yb + q, x = k.x0 - 1 + lane;
          inside = y >= 0 && y < a.H && x >= 0 && x < a.W;  // h2 outside the image is the head conv's zero padding
          TR_T(tr2);
          mbar_wait_nap(bar(B_ACC2FULL + p), (uint32_t)(j & 1));
          TR_ADD(tr_a2, tr2);
        }
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t src = lane_base + (ph ? TM_ACC2 : TM_ACC1) + 64 * p + 32 * half;   // this warp's two 16-channel groups, converted in place
        float v[32];
        tmem_ld16(src, v); tmem_ld16(src + 16, v + 16);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (ph) {                                          // fAffine.2's bias (fAffine.0's rides in the pre-activation)
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] += a.bias2[32 * half + i];
        }
        const bool all_in = __all_sync(0xffffffffu, inside);
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          uint32_t hi[8], lo[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float x0 = fmaxf(v[16 * g + 2 * i], 0.f), x1 = fmaxf(v[16 * g + 2 * i + 1], 0.f);
            hi[i] = pack_bf16(x0, x1);
            lo[i] = pack_bf16(x0 - __uint_as_float(hi[i] << 16), x1 - __uint_as_float(hi[i] & 0xffff0000u));
          }
          if (!all_in) {
#pragma unroll
            for (int i = 0; i < 8; ++i) { hi[i] = inside ? hi[i] : 0u; lo[i] = inside ? lo[i] : 0u; }
          }
          tmem_st8(src + 16 * g, hi);
          if (!a.fast) tmem_st8(src + 16 * g + 8, lo);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        mbar_arrive(bar((ph ? B_H2READY : B_H1READY) + p));
      }
    }
#ifdef BFSR_TC_TRACE
    if (blockIdx.x == 0 && tid == 0) printf("[cf trace] E12: total %lld wait acc1_full %lld acc2_full %lld\n", clock64() - tr_start, tr_a1, tr_a2);
#endif
  } else {
    // ===================== E3: tap sums, cross-sigmoid, FlowStep (NG groups of 4 warps take blocks in rotation) =====================
    const int q = warp & 3, g = (warp - 8) >> 2;
    constexpr int RING = K::RING;
    float* S = reinterpret_cast<float*>(sgen + K::OFF_EXCH);
    TR_DECL(tr_a3 = 0, tr_bc = 0, tr_ld = 0, tr_ex = 0, tr_fl = 0); TR_T(tr_start);
    BlkIter bi; bi.init(a, g < NB ? g : 0);
    for (int b = g; b < NB; b += NG, bi.advance(a, NG)) {
      const Blk& k = bi.k;
      const int yo = k.yb + q - 1, xo = k.x0 + lane;      // the warp holding raster row y finalises output row y - 1
      const bool valid = lane < OUT_W && xo < a.W && yo >= k.y0 && yo < k.y1;
      const long long pix = ((long long)k.n * a.H + yo) * a.W + xo;
      float4 zq[C / 4];
      const float4* zp = reinterpret_cast<const float4*>((const float*)a.z_in.p + pix * a.z_in.cs + a.z_in.coff);
      if (!TAIL && C == 12 && valid) {                             // flow state of the pixel: in flight while the accumulator is awaited / summed
#pragma unroll
        for (int i = 0; i < C / 4; ++i) zq[i] = __ldg(zp + i);
      }
      float u2[C];
      const int R = 4 * b + q;                            // running raster row of the CTA's block stream; ring slot = row % RING
      constexpr int CH = K::CH;
      // accumulator columns of a slot: 3 CH dy + CH dx + co; the tap sums are taken 12 channels at a time (36 columns in registers)
#pragma unroll
      for (int hc = 0; hc < K::NSUB; ++hc) {
        const int slot = TAIL ? hc : (C == 12 ? (b & 1) : 0);
        TR_T(tr0);
        mbar_wait_nap(bar(B_ACC3FULL + 2 * g + hc), (uint32_t)((b / NG) & 1));
        // small ring (C = 24): the previous block has read its rows before this one overwrites the slots they alias
        if (K::CHAIN && hc == 0 && b > 0) mbar_wait_nap(bar(B_BARR + (b - 1) % NG), (uint32_t)(((b - 1) / NG) & 1));
        TR_ADD(tr_a3, tr0); TR_T(tr1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + TM_ACC3 + N3 * slot;
#pragma unroll 1
        for (int dy = 0; dy < 2; ++dy) {                   // needed by the rows below: y + 1 (dy = 0), y (dy = 1)
          float* Sd = S + ((dy * RING + R % RING) * C + CH * hc) * 32 + lane;
#pragma unroll
          for (int h2 = 0; h2 < CH / 12; ++h2) {
            float t[36];
#pragma unroll
            for (int dx = 0; dx < 3; ++dx) { tmem_ld8(t_row + 3 * CH * dy + CH * dx + 12 * h2, t + 12 * dx); tmem_ld4(t_row + 3 * CH * dy + CH * dx + 12 * h2 + 8, t + 12 * dx + 8); }
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int c = 0; c < 12; ++c)
              Sd[(12 * h2 + c) * 32] = t[c] + __shfl_down_sync(0xffffffffu, t[12 + c], 1) + __shfl_down_sync(0xffffffffu, t[24 + c], 2);
          }
        }
#pragma unroll
        for (int h2 = 0; h2 < CH / 12; ++h2) {
          float t[36];
#pragma unroll
          for (int dx = 0; dx < 3; ++dx) { tmem_ld8(t_row + 6 * CH + CH * dx + 12 * h2, t + 12 * dx); tmem_ld4(t_row + 6 * CH + CH * dx + 12 * h2 + 8, t + 12 * dx + 8); }
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
          for (int c = 0; c < 12; ++c)
            u2[CH * hc + 12 * h2 + c] = t[c] + __shfl_down_sync(0xffffffffu, t[12 + c], 1) + __shfl_down_sync(0xffffffffu, t[24 + c], 2);
        }
        TR_ADD(tr_ld, tr1); TR_T(tr2);
        // the previous block's sums are in the ring (awaited BEFORE the accumulator is released: no E3 group can then run two blocks ahead
        // of a waiter, so every waiter sees every phase of the barriers it polls)
        if (b > 0) mbar_wait_nap(bar(B_BARW + (b - 1) % NG), (uint32_t)(((b - 1) / NG) & 1));
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        mbar_arrive(bar(B_ACC3EMPTY + slot));
        TR_ADD(tr_bc, tr2);
      }
      TR_T(tr2);
      mbar_arrive(bar(B_BARW + g));
      if (!TAIL && C != 12 && valid) {                    // C = 24: no registers for it during the tap sums; in flight across the ring hand-off
#pragma unroll
        for (int i = 0; i < C / 4; ++i) zq[i] = __ldg(zp + i);
      }
      constexpr int NHQ = C == 12 ? 6 : 4;
      float4 hq[NHQ];
      if (!TAIL && valid && a.has_hF) {
        const float4* fp = reinterpret_cast<const float4*>((const float*)a.hF.p + pix * a.hF.cs + a.hF.coff);
#pragma unroll
        for (int i = 0; i < NHQ; ++i) hq[i] = __ldg(fp + i);
      }
      mbar_wait_nap(bar(B_BARW + g), (uint32_t)((b / NG) & 1));
      TR_ADD(tr_bc, tr2); TR_T(tr3);
      float h[C];
      {
        const float* S0 = S + ((R + RING - 2) % RING) * C * 32 + lane;
        const float* S1 = S + (RING + (R + RING - 1) % RING) * C * 32 + lane;
#pragma unroll
        for (int c = 0; c < C; ++c) h[c] = S0[c * 32] + S1[c * 32] + u2[c] + a.bias3[c];
      }
      if (K::CHAIN) mbar_arrive(bar(B_BARR + g));
      TR_ADD(tr_ex, tr3); TR_T(tr4);
#pragma unroll
      for (int c = 1; c < C; c += 2) h[c] = __fdividef(1.f, 1.f + __expf(-(h[c] + 2.f))) + a.eps;
      if (TAIL) {                                         // the level's (shiftF, scaleF) pairs: stored, consumed by the FlowSteps of both directions
        if (valid) {
          float4* dst = reinterpret_cast<float4*>((float*)a.z_out.p + pix * a.z_out.cs + a.z_out.coff);
#pragma unroll
          for (int i = 0; i < C / 4; ++i) dst[i] = make_float4(h[4 * i], h[4 * i + 1], h[4 * i + 2], h[4 * i + 3]);
        }
      } else if (valid && !(a.dbg & 16)) flow_apply<C, NHQ>(a, h, pix, zq, hq);
      TR_ADD(tr_fl, tr4);
    }
#ifdef BFSR_TC_TRACE
    if (blockIdx.x == 0 && lane == 0 && q == 0) printf("[cf trace] E3 g%d: total %lld wait acc3_full %lld ld+shuffle %lld wait ring barriers %lld exchange %lld sigmoid+flow %lld\n", g, clock64() - tr_start, tr_a3, tr_ld, tr_bc, tr_ex, tr_fl);
#endif
  }
  __syncthreads();
  if (warp == W_ISSUE_A) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TM_COLS) : "memory");
  }
}

// z (fp32, first C/2 channels) -> z1 operand plane [hi(ZP) | lo(ZP)] bf16 per pixel
template <int C>
__global__ void z1_pack_kernel(const View z, __nv_bfloat16* out, long long npix, int fast) {
  constexpr int ZP = cf::Cfg<C>::ZP;
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= npix) return;
  const float* s = (const float*)z.p + p * z.cs + z.coff;
  uint32_t hi[ZP / 2], lo[ZP / 2];
#pragma unroll
  for (int e = 0; e < ZP / 2; ++e) {
    const float x0 = 2 * e < C / 2 ? s[2 * e] : 0.f, x1 = 2 * e + 1 < C / 2 ? s[2 * e + 1] : 0.f;
    hi[e] = pack_bf16(x0, x1);
    lo[e] = fast ? 0u : pack_bf16(x0 - __uint_as_float(hi[e] << 16), x1 - __uint_as_float(hi[e] & 0xffff0000u));
  }
  uint4* d = reinterpret_cast<uint4*>(out + p * (2 * ZP));
#pragma unroll
  for (int e = 0; e < ZP / 8; ++e) {
    d[e] = make_uint4(hi[4 * e], hi[4 * e + 1], hi[4 * e + 2], hi[4 * e + 3]);
    d[ZP / 8 + e] = make_uint4(lo[4 * e], lo[4 * e + 1], lo[4 * e + 2], lo[4 * e + 3]);
  }
}

void z1_pack(const View& z, void* z1p, int C, cudaStream_t s) {
  BFSR_CHECK(z.fmt == F32 && (C == 12 || C == 24) && z.C >= C / 2, "z1_pack: fp32 view with the C/2 conditioning channels expected (C = 12 or 24)");
  const long long n = z.npix();
  if (n == 0) return;
  snprintf(g_prof_tag, sizeof g_prof_tag, "z1_pack C%d %dx%d", C, z.H, z.W);
  ProfScope prof(PK_OTHER, (double)n * (2 * C + (C == 12 ? 32 : 64)), s);
  const int fast = g_conv_mode == 1;
  if (C == 12) z1_pack_kernel<12><<<cdiv(n, 256), 256, 0, s>>>(z, (__nv_bfloat16*)z1p, n, fast);
  else z1_pack_kernel<24><<<cdiv(n, 256), 256, 0, s>>>(z, (__nv_bfloat16*)z1p, n, fast);
  CUDA_OK(cudaGetLastError());
  count_launch();
}

// ------------------------------------------------------------------ host side
static void put_bf(std::vector<unsigned short>& img, size_t row0, int r, int k, unsigned short v) {   // element k (0..31) of 64-byte row r, SWIZZLE_64B
  img[(row0 + r) * 32 + (size_t)((((k >> 3) ^ ((r >> 1) & 3)) << 3) + (k & 7))] = v;
}

bool coupling_fused_enabled() {
  static const bool off = getenv("BFSR_FUSE_CPL") && atoi(getenv("BFSR_FUSE_CPL")) == 0;
  return !off;
}

// Builds the resident weight image from the three packed convs of a coupling (fp32 [tap][cin_pad][cout_pad] device arrays of pack_conv).
template <int C>
static void pack_fused_t(FusedCouplingW& fw, const ConvW& fA0z, const ConvW& fA2, const ConvW& fA4) {
  using namespace cf;
  using K = Cfg<C>;
  constexpr int N3 = K::N3, ZP = K::ZP;
  auto fetch = [](const ConvW& c, std::vector<float>& w, std::vector<float>& b) {
    w.resize((size_t)c.ks * c.ks * c.cin_pad * c.cout_pad); b.resize(c.cout_pad);
    CUDA_OK(cudaMemcpy(w.data(), c.w, w.size() * 4, cudaMemcpyDeviceToHost));
    CUDA_OK(cudaMemcpy(b.data(), c.bias, b.size() * 4, cudaMemcpyDeviceToHost));
  };
  std::vector<float> wa, ba, wb, bb, wc, bc;
  fetch(fA0z, wa, ba); fetch(fA2, wb, bb); fetch(fA4, wc, bc);
  std::vector<unsigned short> img(K::W_BYTES / 2, 0);
  auto split = [](float w, unsigned short& hi, unsigned short& lo) { hi = f2bf(w); lo = f2bf(w - bf2f(hi)); };
  // W1 (see Cfg): the A operand of a tap is the z1 row [hi(ZP) | lo(ZP)]
  for (int t = 0; t < 9; ++t)
    for (int n = 0; n < 64; ++n)
      for (int ci = 0; ci < fA0z.cin && ci < ZP; ++ci) {
        unsigned short hi, lo; split(wa[((size_t)t * fA0z.cin_pad + ci) * fA0z.cout_pad + n], hi, lo);
        put_bf(img, (size_t)t * 64, n, ci, hi); put_bf(img, (size_t)t * 64, n, ZP + ci, hi);
        if (C == 12) put_bf(img, (size_t)t * 64, n, 16 + ci, lo);
        else put_bf(img, (size_t)(9 + (t >> 1)) * 64, n, 16 * (t & 1) + ci, lo);
      }
  // W2: chunk c, rows 0..63 hi, 64..127 lo
  const size_t r2 = K::W1_BYTES / ROWB;
  for (int c = 0; c < 2; ++c)
    for (int n = 0; n < 64; ++n)
      for (int k = 0; k < 32; ++k) {
        unsigned short hi, lo; split(wb[((size_t)(c * 32 + k)) * fA2.cout_pad + n], hi, lo);
        put_bf(img, r2 + (size_t)c * 128, n, k, hi); put_bf(img, r2 + (size_t)c * 128, 64 + n, k, lo);
      }
  // W3: chunk c, rows tap*C + co (hi), N3 + tap*C + co (lo)
  const size_t r3 = r2 + W2_BYTES / ROWB;
  for (int c = 0; c < 2; ++c)
    for (int t = 0; t < 9; ++t)
      for (int co = 0; co < C; ++co)
        for (int k = 0; k < 32; ++k) {
          unsigned short hi, lo; split(wc[((size_t)t * fA4.cin_pad + c * 32 + k) * fA4.cout_pad + co], hi, lo);
          put_bf(img, r3 + (size_t)c * 2 * N3, t * C + co, k, hi); put_bf(img, r3 + (size_t)c * 2 * N3, N3 + t * C + co, k, lo);
        }
  const size_t r4 = r3 + K::W3_BYTES / ROWB;
  for (int r = 0; r < 32; ++r) put_bf(img, r4, r, r, 0x3F80);
  CUDA_OK(cudaMalloc(&fw.w, K::W_BYTES));
  CUDA_OK(cudaMemcpy(fw.w, img.data(), K::W_BYTES, cudaMemcpyHostToDevice));
  fw.C = C;
  for (int i = 0; i < 64; ++i) { fw.bias1[i] = ba[i]; fw.bias2[i] = bb[i]; }
  for (int i = 0; i < 32; ++i) fw.bias3[i] = i < C ? bc[i] : 0.f;
}
void pack_fused_coupling(FusedCouplingW& fw, const ConvW& fA0z, const ConvW& fA2, const ConvW& fA4, int C) {
  fw = FusedCouplingW();
  static const int max_c = getenv("BFSR_FUSE_CPL_MAXC") ? atoi(getenv("BFSR_FUSE_CPL_MAXC")) : 24;   // 12: keep the C = 24 level on the three-launch chain
  if ((C != 12 && C != 24) || C > max_c || fA0z.cout != 64 || fA2.cin != 64 || fA2.cout != 64 || fA4.cin != 64 || fA4.cout != C ||
      fA0z.cin > cf::Cfg<24>::ZP || (C == 12 && fA0z.cin > 8) || fA0z.ks != 3 || fA2.ks != 1 || fA4.ks != 3) return;
  if (C == 12) pack_fused_t<12>(fw, fA0z, fA2, fA4); else pack_fused_t<24>(fw, fA0z, fA2, fA4);
}
// Feature-only tail of a coupling: fFeatures.2 (1x1 64 -> 64, ReLU) and 24 output channels [co0, co0 + 24) of fFeatures.4 (3x3, cross-sigmoid):
// the whole tail of a C = 12 level, half of it for C = 24 (two launches, each recomputing the cheap 1x1)
void pack_fused_tail(FusedCouplingW& fw, const ConvW& fF2, const ConvW& fF4, int co0) {
  using namespace cf;
  using K = Cfg<24, true>;
  constexpr int N3 = K::N3;
  fw = FusedCouplingW();
  static const bool off = getenv("BFSR_FUSE_TAIL") && atoi(getenv("BFSR_FUSE_TAIL")) == 0;
  if (off || fF2.ks != 1 || fF2.cin != 64 || fF2.cout != 64 || fF4.ks != 3 || fF4.cin != 64 || co0 % 24 != 0 || co0 + 24 > fF4.cout) return;
  auto fetch = [](const ConvW& c, std::vector<float>& w, std::vector<float>& b) {
    w.resize((size_t)c.ks * c.ks * c.cin_pad * c.cout_pad); b.resize(c.cout_pad);
    CUDA_OK(cudaMemcpy(w.data(), c.w, w.size() * 4, cudaMemcpyDeviceToHost));
    CUDA_OK(cudaMemcpy(b.data(), c.bias, b.size() * 4, cudaMemcpyDeviceToHost));
  };
  std::vector<float> wb, bb, wc, bc;
  fetch(fF2, wb, bb); fetch(fF4, wc, bc);
  std::vector<unsigned short> img(K::W_BYTES / 2, 0);
  auto split = [](float w, unsigned short& hi, unsigned short& lo) { hi = f2bf(w); lo = f2bf(w - bf2f(hi)); };
  for (int c = 0; c < 2; ++c)
    for (int n = 0; n < 64; ++n)
      for (int k = 0; k < 32; ++k) {
        unsigned short hi, lo; split(wb[((size_t)(c * 32 + k)) * fF2.cout_pad + n], hi, lo);
        put_bf(img, (size_t)c * 128, n, k, hi); put_bf(img, (size_t)c * 128, 64 + n, k, lo);
      }
  const size_t r3 = W2_BYTES / ROWB;
  for (int c = 0; c < 2; ++c)
    for (int t = 0; t < 9; ++t)
      for (int co = 0; co < 24; ++co)
        for (int k = 0; k < 32; ++k) {
          unsigned short hi, lo; split(wc[((size_t)t * fF4.cin_pad + c * 32 + k) * fF4.cout_pad + co0 + co], hi, lo);
          const size_t base = r3 + (size_t)(c * K::NSUB + co / K::CH) * 2 * N3;
          put_bf(img, base, t * K::CH + co % K::CH, k, hi); put_bf(img, base, N3 + t * K::CH + co % K::CH, k, lo);
        }
  CUDA_OK(cudaMalloc(&fw.w, K::W_BYTES));
  CUDA_OK(cudaMemcpy(fw.w, img.data(), K::W_BYTES, cudaMemcpyHostToDevice));
  fw.C = 24;
  for (int i = 0; i < 64; ++i) { fw.bias1[i] = 0.f; fw.bias2[i] = bb[i]; }
  for (int i = 0; i < 32; ++i) fw.bias3[i] = i < 24 ? bc[co0 + i] : 0.f;
}
void free_fused_coupling(FusedCouplingW& fw) { if (fw.w) cudaFree(fw.w); fw.w = nullptr; }

static int g_cf_sms = 0;

// strips x vertical segments of the tiles -> work items
static void cf_decompose(CfArgs& a, int N, int H, int W) {
  using namespace cf;
  a.strips = cdiv(W, OUT_W);
  // vertical segmentation: the cheapest number of segments per strip under one-CTA-per-SM wave quantisation (each segment recomputes
  // two halo rows; 32 tiles of 320x320: 5 segments of 64 rows = 13 rounds x 17 blocks instead of 3 x 81)
  {
    double best = 1e300; int best_ns = 1;
    for (int ns = 1; ns <= 64 && (ns == 1 || cdiv(H, ns) >= 8); ++ns) {
      const int rows = cdiv(H, ns), nsr = cdiv(H, rows);
      const long long items = (long long)N * a.strips * nsr;
      const double cost = (double)((items + g_cf_sms - 1) / g_cf_sms) * (cdiv(rows + 2, 4) + 0.5);
      if (cost < best * (1.0 - 1e-9)) { best = cost; best_ns = ns; }
    }
    static const int ns_env = getenv("BFSR_CF_SEGS") ? atoi(getenv("BFSR_CF_SEGS")) : 0;
    if (ns_env > 0) best_ns = ns_env;
    a.seg_rows = cdiv(H, best_ns); a.segs = cdiv(H, a.seg_rows);
  }
  a.nblk = cdiv(a.seg_rows + 2, 4);
  const long long items = (long long)N * a.strips * a.segs;
  BFSR_CHECK(items < (1 << 30), "coupling_fused: too many work items");
  a.total_items = (int)items;
}

// z1p_in / z1p_out: [N,H,W,2 ZP] bf16 planes ([hi(ZP) | lo(ZP)] of z1); pre: the BF16X2 64-channel pre-activation slice; f as for the conv
// epilogue (z1op.p != null requests the z1 operand of the next step in z1p_out); hM / hcvec: host copies of f.M / f.cvec
void coupling_fused(const FusedCouplingW& fw, const void* z1p_in, void* z1p_out, const View& pre, const FlowEpi& f, const float* hM,
                    const float* hcvec, float eps, cudaStream_t s) {
  using namespace cf;
  const int C = f.C;
  BFSR_CHECK(fw.w && (C == 12 || C == 24) && fw.C == C, "coupling_fused: weights not packed for C = %d", C);
  const int ZP = C == 12 ? 8 : 16;
  const View& z = f.z_in;
  BFSR_CHECK(pre.fmt == BF16X2 && pre.C == 64 && pre.cs % 8 == 0 && pre.coff % 8 == 0 && pre.plane % 8 == 0 && ((uintptr_t)pre.p % 16) == 0 &&
             pre.N == z.N && pre.H == z.H && pre.W == z.W, "coupling_fused: pre-activation view");
  auto v4 = [](const View& v) { return v.fmt == F32 && v.cs % 4 == 0 && v.coff % 4 == 0 && ((uintptr_t)v.p % 16) == 0; };
  BFSR_CHECK(v4(z) && v4(f.z_out) && z.C == C && f.z_out.C == C && f.z_out.npix() == z.npix() && (!f.hF.p || (v4(f.hF) && f.hF.C == 2 * C)) &&
             (!f.has_mix || (hM && hcvec)) && ((uintptr_t)z1p_in % 16) == 0 && ((uintptr_t)z1p_out % 16) == 0 && (!f.z1op.p || z1p_out),
             "coupling_fused: flow-state views");
  if (z.npix() == 0) return;
  if (!g_cf_sms) { int dev = 0; CUDA_OK(cudaGetDevice(&dev)); CUDA_OK(cudaDeviceGetAttribute(&g_cf_sms, cudaDevAttrMultiProcessorCount, dev)); }
  CfArgs a = CfArgs();
  a.w = (const unsigned char*)fw.w; a.pre_coff = pre.coff;
  a.H = z.H; a.W = z.W; a.N = z.N;
  cf_decompose(a, z.N, z.H, z.W);
  static const int dbg_env = getenv("BFSR_CF_DBG") ? atoi(getenv("BFSR_CF_DBG")) : 0;
  a.dbg = dbg_env;
  a.fast = g_conv_mode == 1 ? 1 : 0;
  a.eps = eps; a.inv = f.inv; a.has_mix = f.has_mix; a.has_hF = f.hF.p ? 1 : 0;
  a.z_in = z; a.z_out = f.z_out; a.hF = f.hF;
  a.z1_out = f.z1op.p ? (__nv_bfloat16*)z1p_out : nullptr;
  for (int i = 0; i < 64; ++i) BFSR_CHECK(fw.bias1[i] == 0.f, "coupling_fused: the z part of fAffine.0 must be packed without a bias");
  memcpy(a.bias2, fw.bias2, sizeof a.bias2); memcpy(a.bias3, fw.bias3, sizeof a.bias3);
  if (f.has_mix) { memcpy(a.M, hM, (size_t)C * C * 4); memcpy(a.cvec, hcvec, (size_t)C * 4); }
  {
    const cuuint64_t px = (cuuint64_t)ZP * 4;      // bytes per pixel of the z1 plane
    const cuuint64_t dims[5] = {(cuuint64_t)(2 * ZP), (cuuint64_t)z.W, (cuuint64_t)z.H, (cuuint64_t)z.N, 1};
    const cuuint64_t strides[4] = {px, (cuuint64_t)z.W * px, (cuuint64_t)z.H * z.W * px, (cuuint64_t)z.N * z.H * z.W * px};
    const cuuint32_t box[5] = {32, 32, 6, 1, 1};
    const cuuint32_t es[5] = {1, 1, 1, 1, 1};
    const CUresult r = encode_tiled()(&a.tm_z1, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(z1p_in), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                      CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    BFSR_CHECK(r == CUDA_SUCCESS, "coupling_fused: cuTensorMapEncodeTiled(z1) failed (%d)", (int)r);
  }
  make_tmap(&a.tm_pre, pre, 32, 4);
  const int grid = a.total_items < g_cf_sms ? a.total_items : g_cf_sms;
  snprintf(g_prof_tag, sizeof g_prof_tag, "cpl-fused C%d %dx%d", C, z.H, z.W);
  ProfScope prof(PK_CONV_TC, 2.0 * (double)z.npix() * (ZP * 9.0 * 64 + 64.0 * 64 + 64.0 * 9 * C), s);
  if (C == 12) {
    CUDA_OK(cudaFuncSetAttribute(coupling_fused_kernel<12, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<12>::SMEM_BYTES));
    coupling_fused_kernel<12, false><<<grid, NTHREADS, Cfg<12>::SMEM_BYTES, s>>>(a);
  } else {
    CUDA_OK(cudaFuncSetAttribute(coupling_fused_kernel<24, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<24>::SMEM_BYTES));
    coupling_fused_kernel<24, false><<<grid, NTHREADS, Cfg<24>::SMEM_BYTES, s>>>(a);
  }
  CUDA_OK(cudaGetLastError());
  count_launch();
}

// hF = cross_sigmoid(conv3x3(relu(conv1x1(in)))) for one step of a C = 12 level: in = BF16X2 64-channel slice of the level's fFeatures.0
// outputs, out = fp32 24-channel (shiftF, scaleF) pairs
void tail_fused(const FusedCouplingW& fw, const View& in, const View& out, float eps, cudaStream_t s) {
  using namespace cf;
  using K = Cfg<24, true>;
  BFSR_CHECK(fw.w && fw.C == 24, "tail_fused: weights not packed");
  BFSR_CHECK(in.fmt == BF16X2 && in.C == 64 && in.cs % 8 == 0 && in.coff % 8 == 0 && in.plane % 8 == 0 && ((uintptr_t)in.p % 16) == 0 &&
             out.fmt == F32 && out.C == 24 && out.cs % 4 == 0 && out.coff % 4 == 0 && ((uintptr_t)out.p % 16) == 0 && in.N == out.N &&
             in.H == out.H && in.W == out.W, "tail_fused: operand views");
  if (in.npix() == 0) return;
  if (!g_cf_sms) { int dev = 0; CUDA_OK(cudaGetDevice(&dev)); CUDA_OK(cudaDeviceGetAttribute(&g_cf_sms, cudaDevAttrMultiProcessorCount, dev)); }
  CfArgs a = CfArgs();
  a.w = (const unsigned char*)fw.w; a.pre_coff = in.coff;
  a.H = in.H; a.W = in.W; a.N = in.N;
  cf_decompose(a, in.N, in.H, in.W);
  static const int dbg_env = getenv("BFSR_CF_DBG") ? atoi(getenv("BFSR_CF_DBG")) : 0;
  a.dbg = dbg_env;
  a.fast = g_conv_mode == 1 ? 1 : 0;
  a.eps = eps;
  a.z_in = out; a.z_out = out;
  memcpy(a.bias2, fw.bias2, sizeof a.bias2); memcpy(a.bias3, fw.bias3, sizeof a.bias3);
  make_tmap(&a.tm_pre, in, 32, 4);
  const int grid = a.total_items < g_cf_sms ? a.total_items : g_cf_sms;
  snprintf(g_prof_tag, sizeof g_prof_tag, "tail-fused 64->64->24 %dx%d", in.H, in.W);
  ProfScope prof(PK_CONV_TC, 2.0 * (double)in.npix() * (64.0 * 64 + 64.0 * 9 * 24), s);
  CUDA_OK(cudaFuncSetAttribute(coupling_fused_kernel<24, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, K::SMEM_BYTES));
  coupling_fused_kernel<24, true><<<grid, NTHREADS, K::SMEM_BYTES, s>>>(a);
  CUDA_OK(cudaGetLastError());
  count_launch();
}

}  // namespace bfsr
