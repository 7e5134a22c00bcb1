/* bfsr_b200 — C ABI of the Blackwell-native flow-SR inference engine.
 *
 * The reference (liyuantsao/BFSR) has no FFI: its operator API for the inference hot
 * path is two Python call surfaces.  Each entry point below names the
 * reference interface it stands behind; paths are relative to /root/reference.
 *
 * Conventions
 *  - Every function returns 0 on success or a negative code; the message is in the
 *    thread-local bfsr_last_error().  Nothing throws across the ABI.
 *  - Tensors at the boundary are contiguous fp32 NCHW, exactly what the reference's
 *    callers hold.  `*_dev` pointers are device memory on the handle's device,
 *    `*_host` pointers are host memory (pinned for best speed).  The library owns only
 *    its packed weights and one grow-only workspace per handle.
 *  - Work is enqueued on the caller's stream (cudaStream_t passed as void*); device
 *    entry points do not synchronise, `*_host` entry points return after the result
 *    is in host memory.
 *  - A handle is bound to one device and is not re-entrant.
 */
#ifndef BFSR_B200_H
#define BFSR_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* One named fp32 tensor of a state_dict (host memory), in the reference's checkpoint layout
 * (SRFlow-LP/code/models/base_model.py:95-124; LINF-LP/train.py:234-248). */
typedef struct {
  const char* name;
  const float* data;
  int32_t ndim;
  int64_t shape[4];
} bfsr_tensor_t;

const char* bfsr_last_error(void);
const char* bfsr_version(void);
/* kernels launched by this thread since the last call with reset!=0 */
int64_t bfsr_launch_count(int reset);

/* Per-kernel-class device timing for the roofline report (off by default; adds two event records per launch).
 * kind: 0 fp32 conv, 1 tcgen05 conv, 2 fused flow step, 3 other.  work = algorithmic FLOPs (convs) or bytes (flow steps). */
int bfsr_prof_enable(int on);
int bfsr_prof_summary(int kind, double* total_ms, double* total_work, int64_t* count);
/* per-shape breakdown of the recorded launches as "tag\tlaunches\tms\twork\n" text; returns the untruncated length */
int bfsr_prof_dump(char* buf, int cap);

/* ------------------------------------------------------------------ SRFlow generator
 * Stands behind SRFlowNet.forward (SRFlow-LP/code/models/modules/SRFlowNet_arch.py:60-82),
 * i.e. behind SRFlowModel.get_encode_z / get_sr (models/SRFlow_model.py:198-222). */
typedef struct {
  int32_t scale;            /* opt['scale']: 4 or 8 */
  int32_t nf, nb, gc;       /* network_G.nf / nb, growth channels (64, 23, 32) */
  int32_t K, L;             /* network_G.flow.K / L */
  int32_t n_no_affine;      /* network_G.flow.additionalFlowNoAffine */
  int32_t hidden;           /* coupling hidden channels (64) */
  int32_t n_blocks;         /* len(network_G.flow.stackRRDB.blocks) */
  int32_t blocks[8];        /* stackRRDB.blocks */
  int32_t split_enable;     /* network_G.flow.split.enable */
  int32_t tile_chunk;       /* LR tiles processed per pass through the workspace (0 = default) */
  int32_t precision;        /* 0 = fp32-accurate: split-bf16 x3 on tcgen05 (parity mode); 1 = bf16 single pass (fast);
                               2 = every conv on the fp32 CUDA-core kernel */
} bfsr_srflow_desc_t;

typedef struct bfsr_srflow bfsr_srflow_t;
typedef struct bfsr_unet bfsr_unet_t;

int bfsr_srflow_create(bfsr_srflow_t** out, const bfsr_srflow_desc_t* desc, const bfsr_tensor_t* weights,
                       int32_t n_weights, int32_t device);
void bfsr_srflow_destroy(bfsr_srflow_t* h);
/* number of latent tensors an encode returns and the shape of latent i for an LR input of (h, w) */
int bfsr_srflow_num_latents(const bfsr_srflow_t* h);
int bfsr_srflow_latent_shape(const bfsr_srflow_t* h, int32_t i, int32_t lr_h, int32_t lr_w, int32_t* C, int32_t* H,
                             int32_t* W);

/* netG(gt=gt, lr=lr, reverse=False, epses=[], add_gt_noise=False): SRFlowNet.normal_flow
 * (SRFlowNet_arch.py:83-116) + FlowUpsamplerNet.encode (FlowUpsamplerNet.py:217-251).
 * lr: (B,3,h,w); gt: (B,3,s*h,s*w); latents[i]: device buffers in the order the reference appends
 * them (Split2d eps first, final z last). */
int bfsr_srflow_encode(bfsr_srflow_t* h, const float* lr_dev, const float* gt_dev, int32_t B, int32_t lr_h,
                       int32_t lr_w, float* const* latents_dev, void* stream);
/* netG(lr=lr, z=None, eps_std=None, reverse=True, epses=latents): SRFlowNet.reverse_flow
 * (SRFlowNet_arch.py:145-158) + FlowUpsamplerNet.decode (FlowUpsamplerNet.py:267-296). sr: (B,3,s*h,s*w). */
int bfsr_srflow_decode(bfsr_srflow_t* h, const float* lr_dev, const float* const* latents_dev, int32_t B,
                       int32_t lr_h, int32_t lr_w, float* sr_dev, void* stream);
/* The whole LP inference path of SRFlow-LP/code/test.py:135-148 in one call:
 * lr_up = bilinear(lr) -> encode -> per-pixel latent normalisation -> prior -> decode (pre-clamp SR).
 * The LR encoder and every feature-only conv run once and are shared by both flow directions. */
int bfsr_srflow_lp_sr(bfsr_srflow_t* h, bfsr_unet_t* prior, const float* lr_dev, int32_t B, int32_t lr_h,
                      int32_t lr_w, float* sr_dev, void* stream);
/* Same with host buffers: H2D copy of lr, compute, D2H copy of sr, synchronises the stream. */
int bfsr_srflow_lp_sr_host(bfsr_srflow_t* h, bfsr_unet_t* prior, const float* lr_host, int32_t B, int32_t lr_h,
                           int32_t lr_w, float* sr_host, void* stream);
/* bytes of workspace currently reserved by the handle */
int64_t bfsr_srflow_workspace_bytes(const bfsr_srflow_t* h);

/* ------------------------------------------------------------------ learned prior (UNet)
 * variant 0: SRFlow-LP UNet.forward(epses) (SRFlow-LP/code/models/unet.py:109-181), one branch per latent;
 * variant 1: LINF-LP UNet.forward(x, lr)   (LINF-LP/models/unet.py:105-167). */
typedef struct {
  int32_t variant;
  int32_t depth, dim, bilinear;
  int32_t n_latents;        /* variant 0 */
  int32_t latent_ch[4];     /* variant 0: channels of each latent (6, 96) */
  int32_t in_chans;         /* variant 1: 27 */
  int32_t precision;
} bfsr_unet_desc_t;

int bfsr_unet_create(bfsr_unet_t** out, const bfsr_unet_desc_t* desc, const bfsr_tensor_t* weights, int32_t n_weights,
                     int32_t device);
void bfsr_unet_destroy(bfsr_unet_t* h);
/* variant 0: prior_model(epses) -> [z0, z1]; latent i is (B, latent_ch[i], H[i], W[i]) NCHW */
int bfsr_unet_forward_srflow(bfsr_unet_t* h, const float* const* latents_dev, const int32_t* H, const int32_t* W,
                             int32_t B, float* const* out_dev, void* stream);

/* variant 1: prior_model(z, inp) (LINF-LP/test.py:147): z (B,in_chans,qh,qw), inp (B,3,h,w) -> (B,in_chans,qh,qw) */
int bfsr_unet_forward_linf(bfsr_unet_t* h, const float* z_dev, const float* inp_dev, int32_t B, int32_t qh, int32_t qw,
                           int32_t lr_h, int32_t lr_w, float* out_dev, void* stream);

/* ------------------------------------------------------------------ LINF ('linf-patch', LINF-LP/models/linf.py:218-428)
 * Stands behind model(op, ...) as called by batched_predict / batched_predict_log_p (LINF-LP/test.py:20-47). */
typedef struct {
  int32_t encoder;          /* 0 = 'edsr-baseline' (edsr.py), 1 = 'rrdb' (rrdb.py) */
  int32_t nb;               /* rrdb: number of RRDB blocks (23) ; edsr-baseline: number of ResBlocks (16) */
  int32_t hidden;           /* hidden_dim (256) */
  int32_t flow_layers;      /* 10 */
  int32_t patch_size;       /* 3 */
  int32_t tile_chunk;       /* images per pass through the workspace (0 = default) */
  int32_t precision;        /* as bfsr_srflow_desc_t.precision */
} bfsr_linf_desc_t;
typedef struct bfsr_linf bfsr_linf_t;

int bfsr_linf_create(bfsr_linf_t** out, const bfsr_linf_desc_t* desc, const bfsr_tensor_t* weights, int32_t n_weights,
                     int32_t device);
void bfsr_linf_destroy(bfsr_linf_t* h);
/* model("gen_feat", inp): inp (B,3,h,w) in [-1,1] -> feat (B,64,h,w)   (linf.py:244-246) */
int bfsr_linf_gen_feat(bfsr_linf_t* h, const float* inp_dev, int32_t B, int32_t lr_h, int32_t lr_w, float* feat_dev,
                       void* stream);
/* model("query_log_p", feat, coord, cell, gt) -> z (B,27,qh,qw)  [mode 0]   (linf.py:248-322)
 * model("query_rgb", feat, coord, cell, zmap) -> (B,3,3qh,3qw)   [mode 1]   (linf.py:324-407)
 * coord: (B,qh,qw,2) (row,col) in [-1,1]; cell: (B,2); zin: gt or zmap (B,27,qh,qw). */
int bfsr_linf_query(bfsr_linf_t* h, const float* feat_dev, int32_t B, int32_t lr_h, int32_t lr_w, const float* coord_dev,
                    const float* cell_dev, int32_t qh, int32_t qw, int32_t mode, const float* zin_dev, float* out_dev,
                    void* stream);
/* The two halves of a query, for callers that keep the per-query affine parameters between the log_p and the rgb pass (the
 * reference recomputes them, and the coef / freq convs per 256-row chunk: LINF-LP/test.py:22-32,40-45):
 * bfsr_linf_affine: coef/freq conv + local Fourier features + MLP (linf.py:251-321) -> affine (B,qh,qw,2*27*flow_layers), NHWC fp32;
 * bfsr_linf_flow:   Flow.forward on gt (mode 0, -> z (B,27,qh,qw)) / Flow.inverse on zmap + fold (mode 1, -> (B,3,3qh,3qw)). */
int bfsr_linf_affine(bfsr_linf_t* h, const float* feat_dev, int32_t B, int32_t lr_h, int32_t lr_w, const float* coord_dev,
                     const float* cell_dev, int32_t qh, int32_t qw, float* affine_dev, void* stream);
int bfsr_linf_flow(bfsr_linf_t* h, const float* affine_dev, const float* zin_dev, int32_t B, int32_t qh, int32_t qw, int32_t mode,
                   float* out_dev, void* stream);
/* Stand-alone Flow of LINF-LP/models/flow.py:12-63 (registry name 'flow', 3x3 patches: D = 27): x (N,27), affine_info
 * (N, 54*n_layers) row-major device buffers; keys `linears.<i>._weight/.bias`, `last._weight/.bias`; inverse != 0 -> Flow.inverse. */
int bfsr_op_linf_flow(const bfsr_tensor_t* weights, int32_t n_weights, int32_t n_layers, int32_t inverse, const float* x_dev,
                      const float* affine_dev, int64_t N, float* out_dev, void* stream);
/* The whole LP path of LINF-LP/test.py:143-171 (--patch, eval_bsize set) in one call: encoder and per-query affine
 * parameters computed once and shared by the log_p and rgb passes; returns pred cropped to (out_h,out_w) + bilinear(inp). */
int bfsr_linf_lp_sr(bfsr_linf_t* h, bfsr_unet_t* prior, const float* inp_dev, int32_t B, int32_t lr_h, int32_t lr_w,
                    const float* coord_dev, const float* cell_dev, const float* gt_lr_up_dev, int32_t qh, int32_t qw,
                    int32_t out_h, int32_t out_w, float* pred_dev, void* stream);
/* Same with host buffers (H2D of inp/coord/cell/gt_lr_up, compute, D2H of pred, synchronises the stream). */
int bfsr_linf_lp_sr_host(bfsr_linf_t* h, bfsr_unet_t* prior, const float* inp_host, int32_t B, int32_t lr_h, int32_t lr_w,
                         const float* coord_host, const float* cell_host, const float* gt_lr_up_host, int32_t qh,
                         int32_t qw, int32_t out_h, int32_t out_w, float* pred_host, void* stream);
/* Test-time input construction of the LINF dataset wrappers (LINF-LP/datasets/wrappers.py:154-238 paired / always_pad = 1,
 * :516-613 arbitrary scale / always_pad = 0; utils.make_coord utils.py:105-120) for a batch of LR images in [0,1]:
 * inp = (lr-0.5)/0.5 (B,3,h,w); coord = centres of the ps x ps HR patches (B,qh,qw,2), zeros in the padded row/column;
 * cell = (2/out_h, 2/out_w) (B,2); gt_lr_up = unfold_ps(lr_up - up(down(lr_up))) (B,3*ps*ps,qh,qw), all bilinear
 * align_corners=False.  With every output pointer NULL only (qh, qw) are returned. */
int bfsr_linf_build_inputs(const float* lr01_dev, int32_t B, int32_t lr_h, int32_t lr_w, int32_t out_h, int32_t out_w,
                           int32_t patch_size, int32_t always_pad, float* inp_dev, float* coord_dev, float* cell_dev,
                           float* gt_lr_up_dev, int32_t* qh_out, int32_t* qw_out, void* stream);

/* ------------------------------------------------------------------ single operators (parity tests, P1 in SURVEY.md §8c)
 * fp32 NCHW in / out on the device; weights in the reference's per-module layout (host). */
/* nn.Conv2d(ks in {1,3}, stride 1, 'same') + bias + activation (0 none, 1 LeakyReLU(0.2), 2 ReLU);
 * impl: 0 = fp32 CUDA-core kernel, 1 = tcgen05 split-bf16 x3 (fp32 operand views, register producers), 2 = tcgen05 bf16 single
 * pass, 3 = x3 with input and output stored as bf16 (hi, lo) planes (TMA-fed A operand, TMA-store epilogue), 4 = as 3 with an
 * fp32 output and the tap-folded evaluation when Cin = 64 and Cout <= 24 */
int bfsr_op_conv2d(const float* x_dev, int32_t B, int32_t Cin, int32_t H, int32_t W, const float* w_host,
                   const float* bias_host, int32_t Cout, int32_t ks, int32_t act, int32_t impl, float* y_dev,
                   void* stream);
/* conv3x3(F.interpolate(x, scale_factor=2, mode='nearest')) + bias (RRDBNet_arch.py:105; the conditioning path of
 * SRFlowNet_arch.py:136): y is (B,Cout,2H,2W).  impl: 0 = fp32 kernel, upsample folded into the loader; 1 = tcgen05,
 * folded loader; 2 = tcgen05, four 2x2 phase convs on the low-res grid with pre-summed weights (16/36 of the MACs). */
int bfsr_op_conv2d_up2(const float* x_dev, int32_t B, int32_t Cin, int32_t H, int32_t W, const float* w_host,
                       const float* bias_host, int32_t Cout, int32_t impl, float* y_dev, void* stream);
/* ---- evaluation metrics on the device (the reference computes them on the host with numpy / cv2) ----
 * calc_psnr (LINF-LP/utils.py:132-151) over a whole (B,C,H,W) tensor: mode 0 = plain, 1 = 'benchmark' (luma of the difference,
 * border of `scale` pixels shaved), 2 = 'div2k' (shave only).  fp64 accumulation; result on the host. */
int bfsr_metric_psnr(const float* sr_dev, const float* hr_dev, int32_t B, int32_t C, int32_t H, int32_t W, int32_t mode,
                     int32_t scale, float rgb_range, double* psnr_out, void* stream);
/* calculate_ssim (LINF-LP/utils.py:154-193): (C,H,W) fp32 images, each multiplied by `mul` first (255 for [0,1] inputs); 11x11
 * Gaussian window (sigma 1.5), valid region, fp64, mean over channels. */
int bfsr_metric_ssim(const float* img1_dev, const float* img2_dev, int32_t C, int32_t H, int32_t W, float mul, double* ssim_out,
                     void* stream);
/* imresize(img, scale) of LINF-LP/imresize.py:136-175 (MATLAB-compatible antialiased bicubic, float input, fp64 arithmetic,
 * mirror padding) on a (C,H,W) image: the LR-consistency metric of test.py:183-200 is calc_psnr(imresize(sr, 1/s), lr).
 * out_dev == NULL returns only the output size ceil(scale * H), ceil(scale * W). */
int bfsr_imresize_bicubic(const float* img_dev, int32_t C, int32_t H, int32_t W, double scale, float* out_dev, int32_t* out_h,
                          int32_t* out_w, void* stream);
/* imresize on a uint8 image (values 0..255 held in fp32): as above, but every pass ends with round-half-even(clip(., 0, 255))
 * exactly as imresize.py:108-110,122-124 does for uint8 input -- the LR-consistency metric of SRFlow-LP/code/test.py:159-160
 * (psnr(lq_orig, imresize(sr, 1 / scale))). */
int bfsr_imresize_bicubic_u8(const float* img_dev, int32_t C, int32_t H, int32_t W, double scale, float* out_dev, int32_t* out_h,
                             int32_t* out_w, void* stream);
/* SSIM as SRFlow-LP/code/Measure.py:46-49 takes it from scikit-image (structural_similarity defaults, multichannel): win_size x
 * win_size UNIFORM window (7), sample covariance (x NP/(NP-1)) when sample_cov != 0, data range 255, K1 = 0.01, K2 = 0.03, mean
 * over the region the window fits in and over channels.  (C,H,W) fp32 images, each multiplied by `mul` first. */
int bfsr_metric_ssim_uniform(const float* img1_dev, const float* img2_dev, int32_t C, int32_t H, int32_t W, float mul,
                             int32_t win_size, int32_t sample_cov, double* ssim_out, void* stream);

/* conv3x3(cat[x_hi (B,Chi,2H,2W), nearest2x(x_lo (B,Clo,H,W))]) + bias + activation: the level-1 coupling conditioning of
 * SRFlowNet_arch.py:118-138 evaluated in ONE pass per output phase (low-res channels: four pre-summed 2x2 taps; hi-res
 * channels: stride-2 parity planes); Chi, Clo multiples of 32; operands stored as bf16 (hi, lo) planes, split-bf16 x3. */
int bfsr_op_conv2d_hi_lo(const float* xhi_dev, const float* xlo_dev, int32_t B, int32_t Chi, int32_t Clo, int32_t H, int32_t W,
                         const float* w_host, const float* bias_host, int32_t Cout, int32_t act, float* y_dev, void* stream);
/* One FlowStep of the reference (FlowStep.normal_flow / reverse_flow, FlowStep.py:88-129: ActNorm, InvertibleConv1x1,
 * CondAffineSeparatedAndCond) through the kernels the engine runs for that step.  `weights` is a state_dict table holding the
 * keys `<prefix>.actnorm.*`, `<prefix>.invconv.weight` and, for a coupling step, `<prefix>.affine.*`; z (B,C,H,W), ft
 * (B,320,H,W, ignored without coupling) and out (B,C,H,W) are NCHW device buffers.  reps > 1 repeats the step (profiling). */
int bfsr_op_flowstep(const bfsr_tensor_t* weights, int32_t n_weights, const char* prefix, int32_t C, int32_t coupling, int32_t reverse,
                     const float* z_dev, const float* ft_dev, int32_t B, int32_t H, int32_t W, float* out_dev, int32_t precision,
                     int32_t reps, void* stream);
/* Split2d (Split.py:49-77, ft = None).  forward: z (B,C,H,W) -> out_z = z1 (B,C-C/2,..), out_eps (B,C/2,..);
 * reverse: z = z1, eps -> out_z (B,C,H,W).  Keys `<prefix>.conv.{weight,bias,logs}`. */
int bfsr_op_split2d(const bfsr_tensor_t* weights, int32_t n_weights, const char* prefix, int32_t C, int32_t reverse, const float* z_dev,
                    const float* eps_dev, int32_t B, int32_t H, int32_t W, float* out_z_dev, float* out_eps_dev, void* stream);
/* flow.squeeze2d / unsqueeze2d (flow.py:122-152) */
int bfsr_op_squeeze2d(const float* x_dev, int32_t B, int32_t C, int32_t H, int32_t W, int32_t reverse, float* y_dev,
                      void* stream);

#ifdef __cplusplus
}
#endif
#endif /* BFSR_B200_H */
