// extern "C" surface of libbfsr_b200.so (declared in include/bfsr_b200.h).
#include "engine.cuh"
#include <cstring>
#include <memory>
#include <atomic>

using namespace bfsr;

namespace bfsr {
void srflow_build(bfsr_srflow* e, const bfsr_tensor_t* weights, int n);
void srflow_run(bfsr_srflow* e, bfsr_unet* prior, int mode, const float* lr, const float* gt, float* const* lat_out,
                const float* const* lat_in, float* sr, int B, int h, int w, cudaStream_t s);
void unet_build(bfsr_unet* u, const bfsr_tensor_t* weights, int n);
std::vector<View> run_unet_srflow(bfsr_unet* u, Arena& A, const std::vector<View>& lat, cudaStream_t s);
View run_unet_linf(bfsr_unet* u, Arena& A, const View& x, const float* lr_nchw, int h, int w, cudaStream_t s);
void op_flowstep(const bfsr_tensor_t* weights, int n, const char* prefix, int C, bool coupling, bool reverse, const float* z_nchw,
                 const float* ft_nchw, int B, int H, int W, float* out_nchw, int reps, cudaStream_t s);
void op_split2d(const bfsr_tensor_t* weights, int n, const char* prefix, int C, bool reverse, const float* z_nchw, const float* eps_nchw,
                int B, int H, int W, float* out_z, float* out_eps, cudaStream_t s);
void linf_build(bfsr_linf* e, const bfsr_tensor_t* weights, int n);
void linf_gen_feat(bfsr_linf* e, const float* inp, int B, int h, int w, float* feat_out, cudaStream_t s);
void linf_query(bfsr_linf* e, const float* feat_nchw, int B, int h, int w, const float* coord, const float* cell, int qh,
                int qw, int mode, const float* zin, float* out, cudaStream_t s);
void linf_affine(bfsr_linf* e, const float* feat_nchw, int B, int h, int w, const float* coord, const float* cell, int qh, int qw,
                 float* aff_out, cudaStream_t s);
void linf_flow_apply(bfsr_linf* e, const float* aff_nhwc, const float* zin, int B, int qh, int qw, int mode, float* out, cudaStream_t s);
void op_linf_flow(const bfsr_tensor_t* weights, int n, int n_layers, bool inverse, const float* x_dev, const float* aff_dev, long long N,
                  float* out_dev, cudaStream_t s);
void linf_lp_sr(bfsr_linf* e, bfsr_unet* prior, const float* inp, int B, int h, int w, const float* coord, const float* cell,
                const float* gt, int qh, int qw, int OH, int OW, float* pred, cudaStream_t s);
}

static thread_local std::string g_err;
namespace bfsr { void set_last_error(const std::string& m) { g_err = m; } }   // shared by the other translation units' entry points

#define API_BEGIN try {
#define API_END                                                      \
  } catch (const std::exception& ex) { g_err = ex.what(); return -1; } \
    catch (...) { g_err = "unknown error"; return -2; }              \
  return 0;

extern "C" {

const char* bfsr_last_error(void) { return g_err.c_str(); }
const char* bfsr_version(void) { return "bfsr_b200 0.1 (sm_100a)"; }
int64_t bfsr_launch_count(int reset) { int64_t v = g_launches; if (reset) g_launches = 0; return v; }

int bfsr_srflow_create(bfsr_srflow_t** out, const bfsr_srflow_desc_t* desc, const bfsr_tensor_t* weights,
                       int32_t n_weights, int32_t device) {
  API_BEGIN
  BFSR_CHECK(out && desc && weights, "null argument");
  int ndev = 0;
  CUDA_OK(cudaGetDeviceCount(&ndev));
  BFSR_CHECK(device >= 0 && device < ndev, "device %d not available (%d CUDA devices)", device, ndev);
  CUDA_OK(cudaSetDevice(device));
  std::unique_ptr<bfsr_srflow> e(new bfsr_srflow());
  e->d = *desc; e->device = device;
  srflow_build(e.get(), weights, n_weights);
  *out = e.release();
  API_END
}
void bfsr_srflow_destroy(bfsr_srflow_t* h) { if (h) { cudaSetDevice(h->device); delete h; } }

int bfsr_srflow_num_latents(const bfsr_srflow_t* h) { return h ? (int)h->latent_C.size() : -1; }
int bfsr_srflow_latent_shape(const bfsr_srflow_t* h, int32_t i, int32_t lr_h, int32_t lr_w, int32_t* C, int32_t* H,
                             int32_t* W) {
  API_BEGIN
  BFSR_CHECK(h && i >= 0 && i < (int)h->latent_C.size(), "latent index out of range");
  *C = h->latent_C[i];
  *H = (lr_h * h->d.scale) >> h->latent_level[i];
  *W = (lr_w * h->d.scale) >> h->latent_level[i];
  API_END
}

int bfsr_srflow_encode(bfsr_srflow_t* h, const float* lr_dev, const float* gt_dev, int32_t B, int32_t lr_h,
                       int32_t lr_w, float* const* latents_dev, void* stream) {
  API_BEGIN
  BFSR_CHECK(h && latents_dev && (B == 0 || (lr_dev && gt_dev)), "null argument");
  g_conv_mode = h->d.precision;
  srflow_run(h, nullptr, 0, lr_dev, gt_dev, latents_dev, nullptr, nullptr, B, lr_h, lr_w, (cudaStream_t)stream);
  API_END
}
int bfsr_srflow_decode(bfsr_srflow_t* h, const float* lr_dev, const float* const* latents_dev, int32_t B,
                       int32_t lr_h, int32_t lr_w, float* sr_dev, void* stream) {
  API_BEGIN
  BFSR_CHECK(h && latents_dev && (B == 0 || (lr_dev && sr_dev)), "null argument");
  g_conv_mode = h->d.precision;
  srflow_run(h, nullptr, 1, lr_dev, nullptr, nullptr, latents_dev, sr_dev, B, lr_h, lr_w, (cudaStream_t)stream);
  API_END
}
int bfsr_srflow_lp_sr(bfsr_srflow_t* h, bfsr_unet_t* prior, const float* lr_dev, int32_t B, int32_t lr_h,
                      int32_t lr_w, float* sr_dev, void* stream) {
  API_BEGIN
  BFSR_CHECK(h && prior && (B == 0 || (lr_dev && sr_dev)), "null argument");
  BFSR_CHECK(prior->device == h->device, "prior and generator live on different devices");
  g_conv_mode = h->d.precision;
  srflow_run(h, prior, 2, lr_dev, nullptr, nullptr, nullptr, sr_dev, B, lr_h, lr_w, (cudaStream_t)stream);
  API_END
}
int bfsr_srflow_lp_sr_host(bfsr_srflow_t* h, bfsr_unet_t* prior, const float* lr_host, int32_t B, int32_t lr_h,
                           int32_t lr_w, float* sr_host, void* stream) {
  API_BEGIN
  BFSR_CHECK(h && prior && (B == 0 || (lr_host && sr_host)), "null argument");
  BFSR_CHECK(prior->device == h->device, "prior and generator live on different devices");
  CUDA_OK(cudaSetDevice(h->device));
  cudaStream_t s = (cudaStream_t)stream;
  const size_t nin = (size_t)B * 3 * lr_h * lr_w * 4;
  const size_t nout = nin * h->d.scale * h->d.scale;
  if (nin > h->stage_in_sz) { if (h->stage_in) cudaFree(h->stage_in); h->stage_in = nullptr; h->stage_in_sz = 0;
                              CUDA_OK(cudaMalloc((void**)&h->stage_in, nin)); h->stage_in_sz = nin; }
  if (nout > h->stage_out_sz) { if (h->stage_out) cudaFree(h->stage_out); h->stage_out = nullptr; h->stage_out_sz = 0;
                                CUDA_OK(cudaMalloc((void**)&h->stage_out, nout)); h->stage_out_sz = nout; }
  if (B > 0) CUDA_OK(cudaMemcpyAsync(h->stage_in, lr_host, nin, cudaMemcpyHostToDevice, s));
  g_conv_mode = h->d.precision;
  srflow_run(h, prior, 2, h->stage_in, nullptr, nullptr, nullptr, h->stage_out, B, lr_h, lr_w, s);
  if (B > 0) CUDA_OK(cudaMemcpyAsync(sr_host, h->stage_out, nout, cudaMemcpyDeviceToHost, s));
  CUDA_OK(cudaStreamSynchronize(s));
  API_END
}
int64_t bfsr_srflow_workspace_bytes(const bfsr_srflow_t* h) { return h ? (int64_t)h->arena.cap : -1; }

int bfsr_unet_create(bfsr_unet_t** out, const bfsr_unet_desc_t* desc, const bfsr_tensor_t* weights, int32_t n_weights,
                     int32_t device) {
  API_BEGIN
  BFSR_CHECK(out && desc && weights, "null argument");
  int ndev = 0;
  CUDA_OK(cudaGetDeviceCount(&ndev));
  BFSR_CHECK(device >= 0 && device < ndev, "device %d not available (%d CUDA devices)", device, ndev);
  CUDA_OK(cudaSetDevice(device));
  std::unique_ptr<bfsr_unet> u(new bfsr_unet());
  static std::atomic<long long> next_serial{1};
  u->d = *desc; u->device = device; u->serial = next_serial.fetch_add(1);
  unet_build(u.get(), weights, n_weights);
  *out = u.release();
  API_END
}
void bfsr_unet_destroy(bfsr_unet_t* h) { if (h) { cudaSetDevice(h->device); delete h; } }

int bfsr_unet_forward_srflow(bfsr_unet_t* h, const float* const* latents_dev, const int32_t* H, const int32_t* W,
                             int32_t B, float* const* out_dev, void* stream) {
  API_BEGIN
  BFSR_CHECK(h && latents_dev && H && W && out_dev, "null argument");
  BFSR_CHECK(h->d.variant == 0, "not an SRFlow-LP prior");
  if (B == 0) return 0;
  CUDA_OK(cudaSetDevice(h->device));
  cudaStream_t s = (cudaStream_t)stream;
  Arena& A = h->arena;
  g_conv_mode = h->d.precision;
  for (int pass = 0; pass < 2; ++pass) {
    A.reset(); A.plan = pass == 0; if (pass == 0) A.peak = 0;
    std::vector<View> lat;
    for (int i = 0; i < h->d.n_latents; ++i) {
      View v = make_view(A, B, H[i], W[i], h->d.latent_ch[i]);
      if (!A.plan) nchw_to_nhwc(latents_dev[i], v, s);
      lat.push_back(v);
    }
    std::vector<View> out = run_unet_srflow(h, A, lat, s);
    if (!A.plan) for (int i = 0; i < h->d.n_latents; ++i) nhwc_to_nchw(out[i], out_dev[i], s);
    if (pass == 0) { A.plan = false; if (A.peak + (1 << 20) > A.cap) { CUDA_OK(cudaStreamSynchronize(s)); A.reserve(A.peak + (1 << 20)); } }
  }
  CUDA_OK(cudaGetLastError());
  API_END
}

int bfsr_unet_forward_linf(bfsr_unet_t* h, const float* z_dev, const float* inp_dev, int32_t B, int32_t qh, int32_t qw,
                           int32_t lr_h, int32_t lr_w, float* out_dev, void* stream) {
  API_BEGIN
  BFSR_CHECK(h && (B == 0 || (z_dev && inp_dev && out_dev)), "null argument");
  BFSR_CHECK(h->d.variant == 1, "not a LINF-LP prior");
  if (B == 0) return 0;
  CUDA_OK(cudaSetDevice(h->device));
  cudaStream_t s = (cudaStream_t)stream;
  Arena& A = h->arena;
  g_conv_mode = h->d.precision;
  for (int pass = 0; pass < 2; ++pass) {
    A.reset(); A.plan = pass == 0; if (pass == 0) A.peak = 0;
    View x = make_view(A, B, qh, qw, h->d.in_chans);
    if (!A.plan) nchw_to_nhwc(z_dev, x, s);
    View o = run_unet_linf(h, A, x, inp_dev, lr_h, lr_w, s);
    if (!A.plan) nhwc_to_nchw(o, out_dev, s);
    if (pass == 0) { A.plan = false; if (A.peak + (1 << 20) > A.cap) { CUDA_OK(cudaStreamSynchronize(s)); A.reserve(A.peak + (1 << 20)); } }
  }
  CUDA_OK(cudaGetLastError());
  API_END
}

// ------------------------------------------------------------------ LINF
int bfsr_linf_create(bfsr_linf_t** out, const bfsr_linf_desc_t* desc, const bfsr_tensor_t* weights, int32_t n_weights,
                     int32_t device) {
  API_BEGIN
  BFSR_CHECK(out && desc && weights, "null argument");
  int ndev = 0;
  CUDA_OK(cudaGetDeviceCount(&ndev));
  BFSR_CHECK(device >= 0 && device < ndev, "device %d not available (%d CUDA devices)", device, ndev);
  CUDA_OK(cudaSetDevice(device));
  std::unique_ptr<bfsr_linf> e(new bfsr_linf());
  e->d = *desc; e->device = device;
  linf_build(e.get(), weights, n_weights);
  *out = e.release();
  API_END
}
void bfsr_linf_destroy(bfsr_linf_t* h) { if (h) { cudaSetDevice(h->device); delete h; } }

int bfsr_linf_gen_feat(bfsr_linf_t* h, const float* inp_dev, int32_t B, int32_t lr_h, int32_t lr_w, float* feat_dev,
                       void* stream) {
  API_BEGIN
  BFSR_CHECK(h && (B == 0 || (inp_dev && feat_dev)), "null argument");
  BFSR_CHECK(B >= 0 && lr_h > 0 && lr_w > 0, "bad input shape");
  g_conv_mode = h->d.precision;
  linf_gen_feat(h, inp_dev, B, lr_h, lr_w, feat_dev, (cudaStream_t)stream);
  API_END
}
int bfsr_linf_query(bfsr_linf_t* h, const float* feat_dev, int32_t B, int32_t lr_h, int32_t lr_w, const float* coord_dev,
                    const float* cell_dev, int32_t qh, int32_t qw, int32_t mode, const float* zin_dev, float* out_dev,
                    void* stream) {
  API_BEGIN
  BFSR_CHECK(h && (B == 0 || (feat_dev && coord_dev && cell_dev && zin_dev && out_dev)), "null argument");
  BFSR_CHECK(mode == 0 || mode == 1, "mode must be 0 (log_p) or 1 (rgb)");
  BFSR_CHECK(B >= 0 && lr_h > 0 && lr_w > 0 && qh > 0 && qw > 0, "bad shape");
  g_conv_mode = h->d.precision;
  linf_query(h, feat_dev, B, lr_h, lr_w, coord_dev, cell_dev, qh, qw, mode, zin_dev, out_dev, (cudaStream_t)stream);
  API_END
}
int bfsr_linf_affine(bfsr_linf_t* h, const float* feat_dev, int32_t B, int32_t lr_h, int32_t lr_w, const float* coord_dev,
                     const float* cell_dev, int32_t qh, int32_t qw, float* affine_dev, void* stream) {
  API_BEGIN
  BFSR_CHECK(h && (B == 0 || (feat_dev && coord_dev && cell_dev && affine_dev)), "null argument");
  BFSR_CHECK(B >= 0 && lr_h > 0 && lr_w > 0 && qh > 0 && qw > 0, "bad shape");
  g_conv_mode = h->d.precision;
  linf_affine(h, feat_dev, B, lr_h, lr_w, coord_dev, cell_dev, qh, qw, affine_dev, (cudaStream_t)stream);
  API_END
}
int bfsr_linf_flow(bfsr_linf_t* h, const float* affine_dev, const float* zin_dev, int32_t B, int32_t qh, int32_t qw, int32_t mode,
                   float* out_dev, void* stream) {
  API_BEGIN
  BFSR_CHECK(h && (B == 0 || (affine_dev && zin_dev && out_dev)), "null argument");
  BFSR_CHECK(mode == 0 || mode == 1, "mode must be 0 (log_p) or 1 (rgb)");
  BFSR_CHECK(B >= 0 && qh > 0 && qw > 0, "bad shape");
  linf_flow_apply(h, affine_dev, zin_dev, B, qh, qw, mode, out_dev, (cudaStream_t)stream);
  API_END
}
int bfsr_op_linf_flow(const bfsr_tensor_t* weights, int32_t n_weights, int32_t n_layers, int32_t inverse, const float* x_dev,
                      const float* affine_dev, int64_t N, float* out_dev, void* stream) {
  API_BEGIN
  BFSR_CHECK(weights && x_dev && affine_dev && out_dev && n_layers > 0 && N > 0 && N < ((int64_t)1 << 31), "bad argument");
  op_linf_flow(weights, n_weights, n_layers, inverse != 0, x_dev, affine_dev, N, out_dev, (cudaStream_t)stream);
  API_END
}
int bfsr_linf_lp_sr(bfsr_linf_t* h, bfsr_unet_t* prior, const float* inp_dev, int32_t B, int32_t lr_h, int32_t lr_w,
                    const float* coord_dev, const float* cell_dev, const float* gt_lr_up_dev, int32_t qh, int32_t qw,
                    int32_t out_h, int32_t out_w, float* pred_dev, void* stream) {
  API_BEGIN
  BFSR_CHECK(h && prior && (B == 0 || (inp_dev && coord_dev && cell_dev && gt_lr_up_dev && pred_dev)), "null argument");
  BFSR_CHECK(prior->device == h->device, "prior and model live on different devices");
  BFSR_CHECK(B >= 0 && lr_h > 0 && lr_w > 0 && qh > 0 && qw > 0, "bad shape");
  g_conv_mode = h->d.precision;
  linf_lp_sr(h, prior, inp_dev, B, lr_h, lr_w, coord_dev, cell_dev, gt_lr_up_dev, qh, qw, out_h, out_w, pred_dev,
             (cudaStream_t)stream);
  API_END
}
int bfsr_linf_lp_sr_host(bfsr_linf_t* h, bfsr_unet_t* prior, const float* inp_host, int32_t B, int32_t lr_h, int32_t lr_w,
                         const float* coord_host, const float* cell_host, const float* gt_lr_up_host, int32_t qh,
                         int32_t qw, int32_t out_h, int32_t out_w, float* pred_host, void* stream) {
  API_BEGIN
  BFSR_CHECK(h && prior && (B == 0 || (inp_host && coord_host && cell_host && gt_lr_up_host && pred_host)), "null argument");
  BFSR_CHECK(prior->device == h->device, "prior and model live on different devices");
  if (B == 0) return 0;
  CUDA_OK(cudaSetDevice(h->device));
  cudaStream_t s = (cudaStream_t)stream;
  const int D = 3 * h->d.patch_size * h->d.patch_size;
  const size_t n_inp = (size_t)B * 3 * lr_h * lr_w, n_coord = (size_t)B * qh * qw * 2, n_cell = (size_t)B * 2,
               n_gt = (size_t)B * D * qh * qw, n_out = (size_t)B * 3 * out_h * out_w;
  const size_t total = (n_inp + n_coord + n_cell + n_gt + n_out + 64) * 4;
  if (total > h->stage_in_sz) { if (h->stage_in) cudaFree(h->stage_in); h->stage_in = nullptr; h->stage_in_sz = 0;
                                CUDA_OK(cudaMalloc((void**)&h->stage_in, total)); h->stage_in_sz = total; }
  auto al = [](size_t n) { return (n + 15) / 16 * 16; };
  float* d_inp = h->stage_in; float* d_coord = d_inp + al(n_inp); float* d_cell = d_coord + al(n_coord);
  float* d_gt = d_cell + al(n_cell); float* d_out = d_gt + al(n_gt);
  CUDA_OK(cudaMemcpyAsync(d_inp, inp_host, n_inp * 4, cudaMemcpyHostToDevice, s));
  CUDA_OK(cudaMemcpyAsync(d_coord, coord_host, n_coord * 4, cudaMemcpyHostToDevice, s));
  CUDA_OK(cudaMemcpyAsync(d_cell, cell_host, n_cell * 4, cudaMemcpyHostToDevice, s));
  CUDA_OK(cudaMemcpyAsync(d_gt, gt_lr_up_host, n_gt * 4, cudaMemcpyHostToDevice, s));
  g_conv_mode = h->d.precision;
  linf_lp_sr(h, prior, d_inp, B, lr_h, lr_w, d_coord, d_cell, d_gt, qh, qw, out_h, out_w, d_out, s);
  CUDA_OK(cudaMemcpyAsync(pred_host, d_out, n_out * 4, cudaMemcpyDeviceToHost, s));
  CUDA_OK(cudaStreamSynchronize(s));
  API_END
}

int bfsr_linf_build_inputs(const float* lr01_dev, int32_t B, int32_t lr_h, int32_t lr_w, int32_t out_h, int32_t out_w,
                           int32_t patch_size, int32_t always_pad, float* inp_dev, float* coord_dev, float* cell_dev,
                           float* gt_lr_up_dev, int32_t* qh_out, int32_t* qw_out, void* stream) {
  API_BEGIN
  BFSR_CHECK(B >= 0 && lr_h > 0 && lr_w > 0 && out_h > 0 && out_w > 0 && patch_size > 0, "bad shape");
  const int ps = patch_size;
  const int pad_h = always_pad ? ps - out_h % ps : (ps - out_h % ps) % ps, pad_w = always_pad ? ps - out_w % ps : (ps - out_w % ps) % ps;
  const int qh = (out_h + pad_h) / ps, qw = (out_w + pad_w) / ps;
  if (qh_out) *qh_out = qh;
  if (qw_out) *qw_out = qw;
  if (!inp_dev && !coord_dev && !cell_dev && !gt_lr_up_dev) return 0;     // shape query
  BFSR_CHECK(B == 0 || (lr01_dev && inp_dev && coord_dev && cell_dev && gt_lr_up_dev), "null argument");
  if (B == 0) return 0;
  cudaStream_t s = (cudaStream_t)stream;
  float* scratch = nullptr;
  CUDA_OK(cudaMallocAsync((void**)&scratch, (size_t)B * 3 * ((size_t)2 * out_h * out_w + (size_t)lr_h * lr_w) * 4, s));
  try { linf_build_inputs(lr01_dev, B, lr_h, lr_w, out_h, out_w, ps, qh, qw, scratch, inp_dev, coord_dev, cell_dev, gt_lr_up_dev, s); }
  catch (...) { cudaFreeAsync(scratch, s); throw; }
  CUDA_OK(cudaFreeAsync(scratch, s));
  CUDA_OK(cudaGetLastError());
  API_END
}

// ------------------------------------------------------------------ single operators
int bfsr_op_flowstep(const bfsr_tensor_t* weights, int32_t n_weights, const char* prefix, int32_t C, int32_t coupling, int32_t reverse,
                     const float* z_dev, const float* ft_dev, int32_t B, int32_t H, int32_t W, float* out_dev, int32_t precision,
                     int32_t reps, void* stream) {
  API_BEGIN
  BFSR_CHECK(weights && prefix && z_dev && out_dev && (!coupling || ft_dev), "null argument");
  BFSR_CHECK(C > 0 && C % 4 == 0 && B > 0 && H > 0 && W > 0, "bad shape");
  const int saved = g_conv_mode;
  g_conv_mode = precision;
  try { op_flowstep(weights, n_weights, prefix, C, coupling != 0, reverse != 0, z_dev, ft_dev, B, H, W, out_dev, reps, (cudaStream_t)stream); }
  catch (...) { g_conv_mode = saved; throw; }
  g_conv_mode = saved;
  API_END
}
int bfsr_op_split2d(const bfsr_tensor_t* weights, int32_t n_weights, const char* prefix, int32_t C, int32_t reverse, const float* z_dev,
                    const float* eps_dev, int32_t B, int32_t H, int32_t W, float* out_z_dev, float* out_eps_dev, void* stream) {
  API_BEGIN
  BFSR_CHECK(weights && prefix && z_dev && out_z_dev && (reverse ? eps_dev != nullptr : out_eps_dev != nullptr), "null argument");
  BFSR_CHECK(C > 0 && C % 2 == 0 && B > 0 && H > 0 && W > 0, "bad shape");
  op_split2d(weights, n_weights, prefix, C, reverse != 0, z_dev, eps_dev, B, H, W, out_z_dev, out_eps_dev, (cudaStream_t)stream);
  API_END
}
int bfsr_op_conv2d(const float* x_dev, int32_t B, int32_t Cin, int32_t H, int32_t W, const float* w_host,
                   const float* bias_host, int32_t Cout, int32_t ks, int32_t act, int32_t impl, float* y_dev,
                   void* stream) {
  API_BEGIN
  BFSR_CHECK(x_dev && w_host && y_dev, "null argument");
  cudaStream_t s = (cudaStream_t)stream;
  ConvW cw = pack_conv(w_host, Cout, Cin, ks, bias_host, nullptr, {}, impl ? 1 : -1);
  float *xn = nullptr, *yn = nullptr, *xb = nullptr, *yb = nullptr;
  const size_t npix = (size_t)B * H * W;
  try {
    CUDA_OK(cudaMalloc((void**)&xn, (npix * ((Cin + 3) & ~3) + 4) * 4));
    CUDA_OK(cudaMalloc((void**)&yn, (npix * Cout + 4) * 4));
    View x; x.p = xn; x.N = B; x.H = H; x.W = W; x.C = Cin; x.cs = (Cin + 3) & ~3;   // pixel stride padded to 16 bytes
    View y; y.p = yn; y.N = B; y.H = H; y.W = W; y.C = Cout; y.cs = Cout;
    nchw_to_nhwc(x_dev, x, s);
    ConvEpi ep; ep.act = act;
    if (impl == 0) conv2d_fp32(cw, x, y, ep, IN_DIRECT, s);
    else if (impl == 5 || impl == 6 || impl == 7) {   // dx-folded 3x3 (Cout <= 32): 5 = fp32 output, 6 = BF16X2 output (TMA store), 7 = 5 in the bf16 single-pass mode
      const int saved = g_conv_mode;
      g_conv_mode = impl == 7 ? 1 : 0;
      try {
        CUDA_OK(cudaMalloc((void**)&xb, (npix * x.cs + 16) * 4));
        CUDA_OK(cudaMalloc((void**)&yb, (npix * Cout + 16) * 4));
        View xv = x; xv.p = xb; xv.fmt = BF16X2; xv.plane = (long long)npix * x.cs;
        View yv = y; yv.p = yb; yv.fmt = BF16X2; yv.plane = (long long)npix * Cout;
        resample(x, xv, RS_COPY, s);
        g_tc_fold = 2;
        BFSR_CHECK(cw.w_tc_f3, "dx-folded packing needs ks = 3 and Cout <= 32");
        conv2d_tc(cw, xv, impl == 6 ? yv : y, ep, IN_DIRECT, s);
        g_tc_fold = -1;
        if (impl == 6) resample(yv, y, RS_COPY, s);
      } catch (...) { g_conv_mode = saved; g_tc_fold = -1; throw; }
      g_conv_mode = saved;
    }
    else if (impl == 4) {   // as 3, but the output stays fp32 (the layout of the coupling-parameter convs: tap-folded when Cin = 64, Cout <= 24)
      const int saved = g_conv_mode;
      g_conv_mode = 0;
      try {
        CUDA_OK(cudaMalloc((void**)&xb, (npix * x.cs + 16) * 4));
        View xv = x; xv.p = xb; xv.fmt = BF16X2; xv.plane = (long long)npix * x.cs;
        resample(x, xv, RS_COPY, s);
        g_tc_fold = 1;
        conv2d_tc(cw, xv, y, ep, IN_DIRECT, s);
        g_tc_fold = -1;
      } catch (...) { g_conv_mode = saved; g_tc_fold = -1; throw; }
      g_conv_mode = saved;
    } else if (impl == 3) {   // tcgen05 split-bf16 x3 with the operand tensors stored as bf16 (hi, lo) planes: TMA-fed A operand
      const int saved = g_conv_mode;
      g_conv_mode = 0;
      try {
        CUDA_OK(cudaMalloc((void**)&xb, (npix * x.cs + 16) * 4));
        CUDA_OK(cudaMalloc((void**)&yb, (npix * Cout + 16) * 4));
        View xv = x; xv.p = xb; xv.fmt = BF16X2; xv.plane = (long long)npix * x.cs;
        View yv = y; yv.p = yb; yv.fmt = BF16X2; yv.plane = (long long)npix * Cout;
        resample(x, xv, RS_COPY, s);
        g_tc_fold = 0;            // the per-tap evaluation (impl 5 / 6 test the dx-folded one)
        conv2d_tc(cw, xv, yv, ep, IN_DIRECT, s);
        g_tc_fold = -1;
        resample(yv, y, RS_COPY, s);
      } catch (...) { g_conv_mode = saved; g_tc_fold = -1; throw; }
      g_conv_mode = saved;
    } else {   // 1 = tcgen05 split-bf16 x3, 2 = tcgen05 bf16 single pass
      const int saved = g_conv_mode;
      g_conv_mode = impl == 1 ? 0 : 1;
      try { conv2d_tc(cw, x, y, ep, IN_DIRECT, s); } catch (...) { g_conv_mode = saved; throw; }
      g_conv_mode = saved;
    }
    nhwc_to_nchw(y, y_dev, s);
    CUDA_OK(cudaStreamSynchronize(s));
  } catch (...) { cudaFree(xn); cudaFree(yn); cudaFree(xb); cudaFree(yb); free_conv(cw); throw; }
  cudaFree(xn); cudaFree(yn); cudaFree(xb); cudaFree(yb); free_conv(cw);
  API_END
}

int bfsr_op_conv2d_up2(const float* x_dev, int32_t B, int32_t Cin, int32_t H, int32_t W, const float* w_host,
                       const float* bias_host, int32_t Cout, int32_t impl, float* y_dev, void* stream) {
  API_BEGIN
  BFSR_CHECK(x_dev && w_host && y_dev, "null argument");
  cudaStream_t s = (cudaStream_t)stream;
  ConvW cw = pack_conv(w_host, Cout, Cin, 3, bias_host, nullptr, {});
  ConvW pw;
  float *xn = nullptr, *yn = nullptr;
  const size_t npix = (size_t)B * H * W;
  const int saved = g_conv_mode;
  try {
    CUDA_OK(cudaMalloc((void**)&xn, (npix * Cin + 4) * 4));
    CUDA_OK(cudaMalloc((void**)&yn, (npix * 4 * Cout + 4) * 4));
    View x; x.p = xn; x.N = B; x.H = H; x.W = W; x.C = Cin; x.cs = Cin;
    View y; y.p = yn; y.N = B; y.H = 2 * H; y.W = 2 * W; y.C = Cout; y.cs = Cout;
    nchw_to_nhwc(x_dev, x, s);
    if (impl == 0) conv2d_fp32(cw, x, y, ConvEpi(), IN_UP2, s);
    else if (impl == 1) { g_conv_mode = 0; conv2d_tc(cw, x, y, ConvEpi(), IN_UP2, s); }
    else {
      g_conv_mode = 0;
      pw = pack_conv_tc_phase(w_host, Cout, Cin, 0, Cin, nullptr);
      // bias first (as the engine does with the hi-res part of the conditioning tensor), then accumulate the phases
      CUDA_OK(cudaMemsetAsync(yn, 0, npix * 4 * Cout * 4, s));
      ConvEpi ep; ep.pre = &y;
      std::vector<float> hb(Cout, 0.f);
      if (bias_host) for (int i = 0; i < Cout; ++i) hb[i] = bias_host[i];
      CUDA_OK(cudaMemcpyAsync(pw.bias, hb.data(), Cout * 4, cudaMemcpyHostToDevice, s));
      CUDA_OK(cudaStreamSynchronize(s));
      conv2d_tc_up2_phase(pw, x, y, ep, s);
    }
    g_conv_mode = saved;
    nhwc_to_nchw(y, y_dev, s);
    CUDA_OK(cudaStreamSynchronize(s));
  } catch (...) { g_conv_mode = saved; cudaFree(xn); cudaFree(yn); free_conv(cw); free_conv(pw); throw; }
  cudaFree(xn); cudaFree(yn); free_conv(cw); free_conv(pw);
  API_END
}

// y = conv3x3(cat[x_hi, nearest2x(x_lo)]) through the single-pass phase evaluation of the tcgen05 kernel (BF16X2 operands)
int bfsr_op_conv2d_hi_lo(const float* xhi_dev, const float* xlo_dev, int32_t B, int32_t Chi, int32_t Clo, int32_t H, int32_t W,
                         const float* w_host, const float* bias_host, int32_t Cout, int32_t act, float* y_dev, void* stream) {
  API_BEGIN
  BFSR_CHECK(xhi_dev && xlo_dev && w_host && y_dev, "null argument");
  cudaStream_t s = (cudaStream_t)stream;
  ConvW cw = pack_conv_tc_phase1(w_host, Cout, Chi + Clo, 0, Chi, Chi, Clo, nullptr, bias_host);
  float *tmp = nullptr, *hb = nullptr, *lb = nullptr, *yb = nullptr;
  const size_t nlo = (size_t)B * H * W, nhi = nlo * 4;
  const int saved = g_conv_mode;
  try {
    CUDA_OK(cudaMalloc((void**)&tmp, (nhi * (Chi > Cout ? Chi : Cout) + nlo * Clo + 16) * 4));
    CUDA_OK(cudaMalloc((void**)&hb, (nhi * Chi + 16) * 4));
    CUDA_OK(cudaMalloc((void**)&lb, (nlo * Clo + 16) * 4));
    CUDA_OK(cudaMalloc((void**)&yb, (nhi * Cout + 16) * 4));
    View xh; xh.p = tmp; xh.N = B; xh.H = 2 * H; xh.W = 2 * W; xh.C = xh.cs = Chi;
    View xhb = xh; xhb.p = hb; xhb.fmt = BF16X2; xhb.plane = (long long)nhi * Chi;
    nchw_to_nhwc(xhi_dev, xh, s); resample(xh, xhb, RS_COPY, s);
    View xl; xl.p = tmp; xl.N = B; xl.H = H; xl.W = W; xl.C = xl.cs = Clo;
    View xlb = xl; xlb.p = lb; xlb.fmt = BF16X2; xlb.plane = (long long)nlo * Clo;
    nchw_to_nhwc(xlo_dev, xl, s); resample(xl, xlb, RS_COPY, s);
    View y; y.p = tmp; y.N = B; y.H = 2 * H; y.W = 2 * W; y.C = y.cs = Cout;
    View ybv = y; ybv.p = yb; ybv.fmt = BF16X2; ybv.plane = (long long)nhi * Cout;
    ConvEpi ep; ep.act = act;
    g_conv_mode = 0;
    conv2d_tc_phase1(cw, xhb, xlb, ybv, ep, s);
    g_conv_mode = saved;
    resample(ybv, y, RS_COPY, s);
    nhwc_to_nchw(y, y_dev, s);
    CUDA_OK(cudaStreamSynchronize(s));
  } catch (...) { g_conv_mode = saved; cudaFree(tmp); cudaFree(hb); cudaFree(lb); cudaFree(yb); free_conv(cw); throw; }
  cudaFree(tmp); cudaFree(hb); cudaFree(lb); cudaFree(yb); free_conv(cw);
  API_END
}

int bfsr_op_squeeze2d(const float* x_dev, int32_t B, int32_t C, int32_t H, int32_t W, int32_t reverse, float* y_dev,
                      void* stream) {
  API_BEGIN
  BFSR_CHECK(x_dev && y_dev, "null argument");
  cudaStream_t s = (cudaStream_t)stream;
  float *xn = nullptr, *yn = nullptr;
  const size_t n = (size_t)B * C * H * W;
  try {
    CUDA_OK(cudaMalloc((void**)&xn, n * 4 + 16));
    CUDA_OK(cudaMalloc((void**)&yn, n * 4 + 16));
    View x; x.p = xn; x.N = B; x.H = H; x.W = W; x.C = C; x.cs = C;
    View y; y.p = yn; y.N = B;
    if (!reverse) { BFSR_CHECK(H % 2 == 0 && W % 2 == 0, "squeeze2d: odd size"); y.H = H / 2; y.W = W / 2; y.C = y.cs = C * 4; }
    else { BFSR_CHECK(C % 4 == 0, "unsqueeze2d: C %% 4 != 0"); y.H = H * 2; y.W = W * 2; y.C = y.cs = C / 4; }
    nchw_to_nhwc(x_dev, x, s);
    if (!reverse) squeeze_copy(x, y, s); else unsqueeze_copy(x, y, s);
    nhwc_to_nchw(y, y_dev, s);
    CUDA_OK(cudaStreamSynchronize(s));
  } catch (...) { cudaFree(xn); cudaFree(yn); throw; }
  cudaFree(xn); cudaFree(yn);
  API_END
}

}  // extern "C"

namespace bfsr {
// conv dispatcher: the tcgen05 implicit-GEMM path takes the shapes it supports, everything else runs on the
// fp32 CUDA-core kernel.
void conv2d(const ConvW& w, const View& in, const View& out, const ConvEpi& epi, int in_mode, cudaStream_t s) {
  if (conv_tc_eligible(w, in, out, epi)) conv2d_tc(w, in, out, epi, in_mode, s);
  else conv2d_fp32(w, in, out, epi, in_mode, s);
}
}  // namespace bfsr
