// Kernel launchers shared by the engines (SRFlow-LP, LINF-LP) and the per-op C ABI.
#pragma once
#include "common.cuh"
#include <vector>

namespace bfsr {

// ------------------------------------------------------------------ convolution
enum Act : int { ACT_NONE = 0, ACT_LRELU = 1, ACT_RELU = 2,
                 ACT_CROSS_SIGMOID = 3 };  // odd channels -> sigmoid(v+2)+eps (thops 'cross' split, FlowAffineCouplingsAblation.py:108-119)
enum InMode : int { IN_DIRECT = 0, IN_UP2 = 1 /* nearest x2 folded into the loader */ };

// Packed conv weights (device).  fp32 layout: [tap][cin_pad][cout_pad], tap = ky*ks+kx.
struct ConvW {
  int ks = 3, cin = 0, cout = 0, cin_pad = 0, cout_pad = 0, co_tile = 64;
  float* w = nullptr;      // fp32 packed
  float* bias = nullptr;   // [cout_pad] (zeros if the conv has none)
  // split-bf16 packing for the tcgen05 path (built lazily by conv_tc.cu)
  void* w_tc = nullptr;
  int tc_kchunks = 0, tc_npad = 0, tc_phase = 0, tc_n_id = 0;   // tc_n_id: identity tap images appended (pre-activation as K chunks)
  void* w_tc_fold = nullptr;   // tap-folded image (3x3, Cin = 64, Cout <= 24): per 32-channel chunk [9*Cout rows hi ; lo], N padded to tc_fold_np
  int tc_fold_np = 0;
  void* w_tc_f3 = nullptr;     // dx-folded image (3x3, Cout <= 32): per (32-channel chunk, dy) [3 blocks of tc_f3_cb rows hi ; lo], row = dx*cb + co
  int tc_f3_cb = 0;            // column stride of a dx block (16 or 32); N = 3*cb
};

// FlowStep fused into the epilogue of a coupling's last conv (whose output h = (shift, scale) pairs never reaches HBM):
//   inverse: z2 = z2/scale - shift ; z = z/scaleF - shiftF (hF) ; z_out = M z - cvec        (FlowStep.reverse_flow, FlowStep.py:113-129)
//   forward: z2 = (z2 + shift)*scale ; [z_out = M z + cvec ; z_out = (z_out + shiftF)*scaleF]  (coupling of step k, then actnorm /
//            invconv / ft-affine of step k+1; has_mix = 0 at the end of a level)
struct FlowEpi {
  int inv = 0, C = 0, has_mix = 1;
  const float* M = nullptr;     // [C][C] row-major (out, in): Mi of this step (inverse) or Mf of the next step (forward)
  const float* cvec = nullptr;  // [C]
  const float* hM = nullptr;    // host copies of M / cvec (fused coupling kernel: they travel as kernel parameters)
  const float* hcvec = nullptr;
  View z_in, z_out;             // fp32, C channels, same resolution as the conv output
  View hF;                      // (shiftF, scaleF) pairs, 2C channels; p == nullptr: none
  View z1op;                    // optional BF16X2 operand copy of the first C/2 output channels (padded to a multiple of 8)
};

struct ConvEpi {
  const FlowEpi* flow = nullptr;   // tcgen05 path only, C in {12, 24}
  int act = ACT_NONE;
  float eps = 1e-4f;            // ACT_CROSS_SIGMOID epsilon
  const View* pre = nullptr;    // added before the activation
  float alpha = 1.f;            // out = act(acc+bias+pre)*alpha + beta1*res1 + beta2*res2
  const View* res1 = nullptr; float beta1 = 0.f;
  const View* res2 = nullptr; float beta2 = 0.f;
  const View* out2 = nullptr;   // second copy of the result (e.g. fp32 residual stream + BF16X2 operand copy)
};

// out(N,H,W,Cout) = conv_ks(in) ; `in` has spatial dims (H,W) or (H/2,W/2) for IN_UP2.
void conv2d(const ConvW& w, const View& in, const View& out, const ConvEpi& epi, int in_mode, cudaStream_t s);
void conv2d_fp32(const ConvW& w, const View& in, const View& out, const ConvEpi& epi, int in_mode, cudaStream_t s);
// tcgen05 path (conv_tc.cu).  g_conv_mode: 0 = split-bf16 x3 (fp32-accurate), 1 = bf16 single pass (fast),
// 2 = never use the tensor-core path (all convs on the fp32 CUDA-core kernel).
extern thread_local int g_conv_mode;
extern thread_local int g_tc_fold;
void pack_conv_tc(ConvW& c, const std::vector<float>& host_packed, int min_cin = -1);
bool conv_tc_eligible(const ConvW& w, const View& in, const View& out, const ConvEpi& epi);
void conv2d_tc(const ConvW& w, const View& in, const View& out, const ConvEpi& epi, int in_mode, cudaStream_t s);
// 3x3 conv over nearest2x(in_lowres) as four 2x2 phase convs on the low-res grid (16/36 of the MACs); out is (N,2H,2W)
ConvW pack_conv_tc_phase(const float* w_oihw, int cout, int cin_src, int c0, int cn, const float* out_scale);
void conv2d_tc_up2_phase(const ConvW& w, const View& in_lowres, const View& out, const ConvEpi& epi, cudaStream_t s);
// single pass over [in_hi (N,2H,2W,hi_cn) | nearest2x(in_lo (N,H,W,lo_cn))] (both BF16X2): out (N,2H,2W,cout); 1600/2880 of the MACs
// of the plain 3x3 for the level-1 conditioning tensor (64 hi-res + 256 upsampled channels)
ConvW pack_conv_tc_phase1(const float* w_oihw, int cout, int cin_src, int hi_c0, int hi_cn, int lo_c0, int lo_cn,
                          const float* out_scale, const float* bias);
void conv2d_tc_phase1(const ConvW& w, const View& in_hi, const View& in_lo, const View& out, const ConvEpi& epi, cudaStream_t s);

// Host-side packing: src is OIHW fp32 [cout][cin_src][ks][ks]; `out_scale` (optional) multiplies the weights per
// output channel, `bias` is the FINAL bias (already scaled); the input-channel gather map makes packed input channel
// i read source channel map[i] (-1 = zero) so callers can split / pad / reorder concatenated inputs.
ConvW pack_conv(const float* w_oihw, int cout, int cin_src, int ks, const float* bias, const float* out_scale,
                const std::vector<int>& cin_map, int tc_min_cin = -1 /* smallest Cin packed for the tcgen05 path; -1 = default (32) */);
void free_conv(ConvW& w);

// ------------------------------------------------------------------ fused coupling step (coupling_fused.cu), C = 12 / 24 levels
// fAffine.0 (z part) -> ReLU -> fAffine.2 -> ReLU -> fAffine.4 -> cross-sigmoid -> FlowStep in ONE launch, hidden maps on chip
struct FusedCouplingW {
  void* w = nullptr;                       // resident weight image (split-bf16, pre-swizzled); null: shape not eligible
  int C = 0;                               // 12 or 24
  float bias1[64], bias2[64], bias3[32];
};
bool coupling_fused_enabled();             // BFSR_FUSE_CPL=0 keeps the three-launch chain
void pack_fused_coupling(FusedCouplingW& fw, const ConvW& fA0z, const ConvW& fA2, const ConvW& fA4, int C);
void free_fused_coupling(FusedCouplingW& fw);
// z1p_*: [N,H,W,2 ZP] bf16 = [hi(ZP) | lo(ZP)] of the conditioning half z1 (first C/2 channels of z, ZP = 8 / 16); pre: BF16X2 64-channel slice of the
// feature-only pre-activations; f: as for the conv epilogue (f.z1op.p != null asks for the next step's z1 operand in z1p_out);
// hM / hcvec: HOST copies of f.M / f.cvec (they travel as kernel parameters)
void coupling_fused(const FusedCouplingW& fw, const void* z1p_in, void* z1p_out, const View& pre, const FlowEpi& f, const float* hM,
                    const float* hcvec, float eps, cudaStream_t s);
void z1_pack(const View& z, void* z1p, int C, cudaStream_t s);
// feature-only tail of a coupling, 24 output channels per launch: hF = cross_sigmoid(conv3x3(relu(conv1x1(in)))) (fFeatures.2 / .4)
void pack_fused_tail(FusedCouplingW& fw, const ConvW& fF2, const ConvW& fF4, int co0 = 0);   // output channels [co0, co0 + 24) of fF4
void tail_fused(const FusedCouplingW& fw, const View& in, const View& out, float eps, cudaStream_t s);

// ------------------------------------------------------------------ layout / resampling
void nchw_to_nhwc(const float* src, const View& dst, cudaStream_t s);
void nhwc_to_nchw(const View& src, float* dst, cudaStream_t s);
enum Resample : int { RS_COPY = 0, RS_NEAREST_UP2, RS_NEAREST_DOWN2, RS_AVG_DOWN2, RS_MAXPOOL2,
                      RS_BILINEAR_UP2_AC /* align_corners=True x2 + zero pad to dst size (unet.py:80-91) */ };
void resample(const View& src, const View& dst, int mode, cudaStream_t s);
// F.interpolate(x, scale_factor=s, mode='bilinear', align_corners=False) on NCHW input -> NHWC view
void bilinear_up_nchw(const float* src, int N, int C, int h, int w, int scale, const View& dst, cudaStream_t s);

// ------------------------------------------------------------------ flow
struct StepW {      // one FlowStep (device pointers, fp32)
  int C = 0; bool coupling = false;
  float* Mf = nullptr;  float* cf = nullptr;   // forward: y = Mf z + cf   (W diag(e^logs), W (b*e^logs))
  float* Mi = nullptr;  float* ci = nullptr;   // inverse: z = Mi y - ci   (diag(e^-logs) W^-1, bias)
  float* MfT = nullptr; float* MiT = nullptr;  // the same matrices transposed ([in][out]) for the shared-memory tile kernels
  std::vector<float> hMf, hcf, hMi, hci;       // host copies (the fused coupling kernel takes them as kernel parameters)
};
// encode half-step: [finish previous coupling with h_prev] -> actnorm -> invconv -> [ft-affine with hF]
//   squeeze_in: z_in is the un-squeezed tensor (N,2H,2W,C/4) read through the Squeeze2d index map (flow.py:122-134)
//   z1op (optional, both directions): BF16X2 view that receives the operand copy of the first C/2 output channels (zero-padded
//   to a multiple of 8) for the next coupling's z-dependent conv, written by the same kernel
void flowstep_fwd(const StepW& w, const View& z_in, bool squeeze_in, const View* h_prev, const View* hF,
                  const View& z_out, cudaStream_t s, const View* z1op = nullptr);
// z2 = (z2 + shift) * scale with (shift,scale) pairs in h   (FlowAffineCouplingsAblation.py:72-76)
void coupling_finish(const View& z, const View& h, const View& z_out, cudaStream_t s);
// decode step: coupling^-1 (h) -> ft-affine^-1 (hF) -> invconv^-1 -> actnorm^-1 ; unsqueeze_out folds Unsqueeze2d
void flowstep_inv(const StepW& w, const View& z_in, const View* h, const View* hF, const View& z_out,
                  bool unsqueeze_out, cudaStream_t s, const View* z1op = nullptr);
// Split2d (Split.py:49-77): h holds (mean,logs) pairs
void split_fwd(const View& z, const View& h, const View& z1_out, const View& eps_out, cudaStream_t s);
void split_inv(const View& z1, const View& h, const View& eps, const View& z_out, cudaStream_t s);
// per-pixel channel normalisation (test.py:141-145)
void normalise_latent(const View& e, const View& out, cudaStream_t s);
// squeeze / unsqueeze as plain copies (used only where they cannot be folded)
void squeeze_copy(const View& src, const View& dst, cudaStream_t s);
void unsqueeze_copy(const View& src, const View& dst, cudaStream_t s);

// ------------------------------------------------------------------ LINF query side (linf_kernels.cu)
// local Fourier features of every query: cf = [coef | freq] maps (B,h,w,2*hid) -> out (B,qh,qw,4*hid)   (linf.py:251-309)
void linf_features(const View& cf, const float* coord, const float* cell, const float* phase, const View& out, int qh, int qw,
                   cudaStream_t s);
// 27-dim conditional flow; forward: zin (B,27,qh,qw) NCHW -> out same shape; inverse: -> out (B,3,OH,OW) NCHW with the
// 3x3 fold, crop to (OH,OW) and optional `+ bilinear(inp)` fused   (flow.py:44-63, linf.py:401-406, test.py:168-171)
void linf_flow(bool inverse, const float* M, const float* bias, int n_layers, const View& aff, const float* zin, int B,
               int qh, int qw, float* out, int OH, int OW, const float* inp, int h, int w, int ps, cudaStream_t s);
void conv3x3_s3_lrelu(const float* x_nchw, int B, int Cin, int h, int w, const float* w_oihw_dev, const float* bias_dev,
                      const View& out, cudaStream_t s);
void bilinear_resize(const View& src, const View& dst, cudaStream_t s);
// test-time inputs of the LINF wrappers (datasets/wrappers.py:154-238, 516-613) for a batch of LR images in [0,1];
// scratch: B*3*(2*H*W + h*w) floats
void linf_build_inputs(const float* lr01, int B, int h, int w, int H, int W, int ps, int qh, int qw, float* scratch, float* inp,
                       float* coord, float* cell, float* gt, cudaStream_t s);

}  // namespace bfsr
