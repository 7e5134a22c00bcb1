// bfsr_b200 — shared device/host plumbing for the sm_100a kernels.
//
// Data layout in HBM: every activation is NHWC ("pixel-major"),
// addressed through a `View` = base pointer + channel stride + channel offset, so a
// consumer can read any channel prefix of a dense-block buffer and a producer can
// write its outputs straight into a channel slice (the reference's torch.cat calls,
// RRDBNet_arch.py:39-45, SRFlowNet_arch.py:136, unet.py:96, never materialise).
// Two element formats:
//   F32   : one fp32 plane                      (flow state z, latents, boundary tensors)
//   BF16X2: two bf16 planes hi, lo with x ≈ hi+lo (operands of the split-bf16 tcgen05 convs)
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <stdexcept>
#include <map>
#include <vector>
#include <functional>

namespace bfsr {

enum Fmt : int { F32 = 0, BF16X2 = 1 };

struct View {
  void* p = nullptr;     // base of plane 0
  int N = 0, H = 0, W = 0;
  int C = 0;             // channels visible through this view
  int cs = 0;            // channel stride of the underlying buffer (elements per pixel)
  int coff = 0;          // first channel of the view inside the buffer
  int fmt = F32;
  long long plane = 0;   // elements between plane hi and plane lo (BF16X2 only)

  __host__ __device__ long long npix() const { return (long long)N * H * W; }
  View slice(int c0, int c) const { View v = *this; v.coff = coff + c0; v.C = c; return v; }
};

__device__ __forceinline__ float ld(const View& v, long long pix, int c) {
  long long i = pix * v.cs + v.coff + c;
  if (v.fmt == F32) return ((const float*)v.p)[i];
  const __nv_bfloat16* b = (const __nv_bfloat16*)v.p;
  return __bfloat162float(b[i]) + __bfloat162float(b[i + v.plane]);
}
__device__ __forceinline__ void st(const View& v, long long pix, int c, float x) {
  long long i = pix * v.cs + v.coff + c;
  if (v.fmt == F32) { ((float*)v.p)[i] = x; return; }
  __nv_bfloat16* b = (__nv_bfloat16*)v.p;
  __nv_bfloat16 hi = __float2bfloat16_rn(x);
  b[i] = hi;
  b[i + v.plane] = __float2bfloat16_rn(x - __bfloat162float(hi));
}

// ---------------------------------------------------------------- errors
struct Error : std::runtime_error { using std::runtime_error::runtime_error; };

#define BFSR_CHECK(cond, ...)                                                        \
  do { if (!(cond)) { char _b[512]; snprintf(_b, sizeof _b, __VA_ARGS__);            \
       throw ::bfsr::Error(std::string(__FILE__) + ":" + std::to_string(__LINE__) + ": " + _b); } } while (0)

#define CUDA_OK(expr)                                                                \
  do { cudaError_t _e = (expr); if (_e != cudaSuccess)                               \
       throw ::bfsr::Error(std::string(__FILE__) + ":" + std::to_string(__LINE__) + ": " #expr ": " + \
                           cudaGetErrorString(_e)); } while (0)

// ---------------------------------------------------------------- workspace arena
// Grow-only bump allocator; one per engine handle.  `plan` mode only measures.
struct Arena {
  char* base = nullptr;
  size_t cap = 0, off = 0, peak = 0;
  bool plan = false;
  void reset() { off = 0; }
  void* alloc(size_t bytes) {
    size_t a = (off + 1023) & ~size_t(1023);
    off = a + bytes;
    if (off > peak) peak = off;
    if (plan) return (void*)(uintptr_t)(0x1000 + a);   // never dereferenced
    BFSR_CHECK(off <= cap, "workspace arena overflow (%zu > %zu)", off, cap);
    return base + a;
  }
  void reserve(size_t bytes) {
    if (bytes <= cap) return;
    if (base) CUDA_OK(cudaFree(base));
    base = nullptr; cap = 0;
    CUDA_OK(cudaMalloc((void**)&base, bytes));
    cap = bytes;
  }
  ~Arena() { if (base) cudaFree(base); }
};

inline View make_view(Arena& a, int N, int H, int W, int C, int fmt = F32) {
  View v; v.N = N; v.H = H; v.W = W; v.C = C; v.cs = C; v.coff = 0; v.fmt = fmt;
  size_t n = (size_t)N * H * W * C;
  if (fmt == F32) v.p = a.alloc(n * 4);
  else { v.plane = (long long)n; v.p = a.alloc(n * 4); }
  return v;
}

// launch counter (bench.py reports `gpu_launches`)
extern thread_local long long g_launches;
inline void count_launch(int n = 1) { g_launches += n; }

inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

// ---------------------------------------------------------------- per-kernel-class device timing (bench.py roofline)
// When enabled, launchers bracket each launch with CUDA events on the launching stream; bfsr_prof_summary()
// returns the summed device time and algorithmic work (flops for convs, bytes for flow steps) per class.
enum ProfKind : int { PK_CONV_FP32 = 0, PK_CONV_TC = 1, PK_FLOWSTEP = 2, PK_OTHER = 3, PK_COUNT = 4 };
extern thread_local char g_prof_tag[96];   // optional shape tag consumed by the next prof_begin (bfsr_prof_dump groups by it)
void prof_begin(int kind, double work, cudaStream_t s);
void prof_end(cudaStream_t s);
struct ProfScope {
  cudaStream_t s;
  ProfScope(int kind, double work, cudaStream_t st) : s(st) { prof_begin(kind, work, st); }
  ~ProfScope() { prof_end(s); }
};

bool prof_enabled();

// ---------------------------------------------------------------- CUDA-graph replay of a fixed launch sequence
// The engines' plans are static per (shapes, buffers): the second call with the same key captures the launch sequence on an
// internal stream and later calls replay the instantiated graph on the caller's stream -- one cudaGraphLaunch instead of hundreds
// of launches, each with its tensor-map encodes (LINF config 3: ~100 launches for ~2 ms of device work; SRFlow at 4 tiles per GPU:
// ~860 launches per 40 ms step).  BFSR_GRAPH=0 disables it; it is bypassed while the per-launch profiler is on.  A cache belongs
// to one engine handle (not re-entrant, like the handle) and must be cleared when the handle's arena moves.
struct GraphCache {
  struct Entry { cudaGraphExec_t exec = nullptr; long long launches = 0; int seen = 0; };
  std::map<std::vector<long long>, Entry> m;
  cudaStream_t cap = nullptr;
  void clear();
  ~GraphCache();
};
bool graphs_enabled();
// body(stream) must enqueue everything on the given stream and must not synchronise, allocate or free
void run_graphed(GraphCache& gc, const std::vector<long long>& key, cudaStream_t s, const std::function<void(cudaStream_t)>& body);

}  // namespace bfsr
