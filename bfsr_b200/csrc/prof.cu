// Optional per-launch device timing with CUDA events (used by bench.py for the roofline numbers; off by default).
#include "common.cuh"
#include <vector>
#include <map>
#include <string>
#include <cstring>
#include <cstdlib>

namespace bfsr {

struct ProfRec { cudaEvent_t a, b; int kind; double work; std::string tag; };
thread_local char g_prof_tag[96] = "";
static thread_local bool g_prof_on = false;
static thread_local std::vector<ProfRec> g_recs;
static thread_local std::vector<cudaEvent_t> g_pool;

static cudaEvent_t get_event() {
  if (!g_pool.empty()) { cudaEvent_t e = g_pool.back(); g_pool.pop_back(); return e; }
  cudaEvent_t e; CUDA_OK(cudaEventCreate(&e)); return e;
}

void prof_begin(int kind, double work, cudaStream_t s) {
  if (!g_prof_on) return;
  ProfRec r; r.a = get_event(); r.b = get_event(); r.kind = kind; r.work = work; r.tag = g_prof_tag; g_prof_tag[0] = 0;
  CUDA_OK(cudaEventRecord(r.a, s));
  g_recs.push_back(r);
}
void prof_end(cudaStream_t s) {
  if (!g_prof_on || g_recs.empty()) return;
  cudaEventRecord(g_recs.back().b, s);
}

bool prof_enabled() { return g_prof_on; }

// ------------------------------------------------------------------ CUDA-graph replay (see common.cuh)
bool graphs_enabled() {
  static const bool on = !(getenv("BFSR_GRAPH") && atoi(getenv("BFSR_GRAPH")) == 0);
  return on && !g_prof_on;
}
void GraphCache::clear() {
  for (auto& kv : m) if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
  m.clear();
}
GraphCache::~GraphCache() { clear(); if (cap) cudaStreamDestroy(cap); }

void run_graphed(GraphCache& gc, const std::vector<long long>& key, cudaStream_t s, const std::function<void(cudaStream_t)>& body) {
  if (!graphs_enabled()) { body(s); return; }
  if (gc.m.size() > 16 && !gc.m.count(key)) gc.clear();           // bounded: callers cycling through many shapes just re-capture
  GraphCache::Entry& e = gc.m[key];
  if (!e.exec) {
    if (e.seen < 0 || e.seen++ == 0) { body(s); return; }         // first sight of a key: plain launches (one-off calls never capture)
    if (!gc.cap) CUDA_OK(cudaStreamCreateWithFlags(&gc.cap, cudaStreamNonBlocking));
    const long long l0 = g_launches;
    CUDA_OK(cudaStreamBeginCapture(gc.cap, cudaStreamCaptureModeThreadLocal));
    cudaGraph_t g = nullptr;
    bool ok = true;
    static const bool dbg = getenv("BFSR_GRAPH_DEBUG") && atoi(getenv("BFSR_GRAPH_DEBUG"));
    try { body(gc.cap); } catch (const std::exception& ex) { ok = false; if (dbg) fprintf(stderr, "[bfsr graph] capture body failed: %s\n", ex.what()); }
    const cudaError_t ce = cudaStreamEndCapture(gc.cap, &g);
    if (ce != cudaSuccess || !g) { ok = false; if (dbg) fprintf(stderr, "[bfsr graph] end capture: %s\n", cudaGetErrorString(ce)); }
    else if (dbg) fprintf(stderr, "[bfsr graph] captured %lld launches\n", g_launches - l0);
    e.launches = g_launches - l0;
    g_launches = l0;
    if (ok && cudaGraphInstantiate(&e.exec, g, 0) != cudaSuccess) { e.exec = nullptr; ok = false; }
    if (g) cudaGraphDestroy(g);
    if (!ok) {                                                     // not capturable: this key keeps the plain launches (errors resurface there)
      cudaGetLastError();
      e.seen = -1;
      body(s);
      return;
    }
  }
  CUDA_OK(cudaGraphLaunch(e.exec, s));
  g_launches += e.launches;
}

}  // namespace bfsr

using namespace bfsr;

extern "C" {
int bfsr_prof_enable(int on) {
  g_prof_on = on != 0;
  for (auto& r : g_recs) { g_pool.push_back(r.a); g_pool.push_back(r.b); }
  g_recs.clear();
  return 0;
}
// Synchronises the device, then reports the sum over recorded launches of class `kind`.
int bfsr_prof_summary(int kind, double* total_ms, double* total_work, int64_t* count) {
  if (cudaDeviceSynchronize() != cudaSuccess) return -1;
  double ms = 0, work = 0; int64_t n = 0;
  for (auto& r : g_recs) {
    if (r.kind != kind) continue;
    float t = 0.f;
    if (cudaEventElapsedTime(&t, r.a, r.b) != cudaSuccess) return -1;
    ms += t; work += r.work; ++n;
  }
  if (total_ms) *total_ms = ms;
  if (total_work) *total_work = work;
  if (count) *count = n;
  return 0;
}
// Per-shape breakdown of the recorded launches: "tag<TAB>launches<TAB>ms<TAB>work" lines into buf (truncated to cap).
int bfsr_prof_dump(char* buf, int cap) {
  if (cudaDeviceSynchronize() != cudaSuccess || !buf || cap <= 0) return -1;
  struct Agg { double ms = 0, work = 0; long long n = 0; };
  std::map<std::string, Agg> m;
  for (auto& r : g_recs) {
    float t = 0.f;
    if (cudaEventElapsedTime(&t, r.a, r.b) != cudaSuccess) return -1;
    Agg& a = m[r.tag.empty() ? std::string("kind") + std::to_string(r.kind) : r.tag];
    a.ms += t; a.work += r.work; ++a.n;
  }
  std::string out;
  for (auto& kv : m) {
    char line[256];
    snprintf(line, sizeof line, "%s\t%lld\t%.4f\t%.6g\n", kv.first.c_str(), kv.second.n, kv.second.ms, kv.second.work);
    out += line;
  }
  strncpy(buf, out.c_str(), cap - 1); buf[cap - 1] = 0;
  return (int)out.size();
}
}
