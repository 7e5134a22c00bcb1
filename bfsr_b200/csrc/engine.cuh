// Engine-side structures: packed weights and the per-handle state of the SRFlow-LP path.
#pragma once
#include "ops.cuh"
#include "../../include/bfsr_b200.h"
#include <map>
#include <string>
#include <vector>

namespace bfsr {

// name -> tensor lookup over the caller's state_dict table
struct Weights {
  std::map<std::string, const bfsr_tensor_t*> m;
  Weights(const bfsr_tensor_t* t, int n) { for (int i = 0; i < n; ++i) m[t[i].name] = &t[i]; }
  const bfsr_tensor_t& get(const std::string& k) const {
    auto it = m.find(k);
    BFSR_CHECK(it != m.end(), "state_dict is missing key '%s'", k.c_str());
    return *it->second;
  }
  const float* data(const std::string& k, std::initializer_list<int64_t> shape) const {
    const bfsr_tensor_t& t = get(k);
    BFSR_CHECK((size_t)t.ndim == shape.size(), "'%s': rank %d, expected %zu", k.c_str(), t.ndim, shape.size());
    int i = 0;
    for (int64_t s : shape) { BFSR_CHECK(t.shape[i] == s, "'%s': dim %d is %lld, expected %lld", k.c_str(), i,
                                         (long long)t.shape[i], (long long)s); ++i; }
    return t.data;
  }
  bool has(const std::string& k) const { return m.count(k) != 0; }
};

struct CouplingW {
  ConvW fF2, fF4;          // fFeatures tail (per step): 1x1 64->64 (+ReLU), 3x3 64->2C (cross-sigmoid)
  ConvW fA0z, fA2, fA4;    // fAffine: z-part of the first conv, 1x1, 3x3 64->C (cross-sigmoid)
  FusedCouplingW fz;       // the same three convs packed for the one-launch coupling step (C = 12 / 24 levels)
  FusedCouplingW ftail, ftail2;   // fF2 + 24 output channels of fF4 packed for the one-launch feature-only tail (C = 12: all of hF; C = 24: two halves)
};
struct LayerW {
  int kind = 0;            // 0 squeeze, 1 nocoupling, 2 coupling, 3 split
  int C = 0, level = 0;
  int k_in_level = -1;     // index of a coupling step inside its level
  StepW step;
  CouplingW cp;
  ConvW split_conv;
};
struct LevelW {
  int C = 0, n_coupling = 0;
  ConvW fF0_all;           // ft -> n_coupling*64, ReLU            (fFeatures.0 of every step, ActNorm folded)
  ConvW fA0ft_all;         // ft -> n_coupling*64, pre-activation  (ft slice of fAffine.0 of every step)
  // level fed by [hi-res base 64 | nearest2x(taps)] (SRFlowNet_arch.py:136): the same two convs packed for the single-pass
  // phase evaluation (low-res taps: four pre-summed 2x2 taps per output phase; hi-res part: parity planes)
  bool has_phase = false;
  ConvW fF0_1p, fA0ft_1p;
};

struct RRDBW {
  ConvW conv_first, trunk_conv;
  std::vector<ConvW> rdb;     // nb*3*5
  std::vector<ConvW> upconv;  // upconv1.. as needed by the top level
};

struct UNetBranchW {
  int nf = 0, nf_pad = 0, gc = 64;
  ConvW dense[5];
  ConvW lr_dense[5];          // LINF variant: lr_proj.2
  float* lr_w = nullptr; float* lr_b = nullptr;   // LINF variant: lr_proj.0 (3 -> in_chans, stride 3), raw OIHW
  ConvW inc[2];
  std::vector<ConvW> down;   // 2 per level
  std::vector<ConvW> up;     // 2 per level
  ConvW outc;
};

std::vector<double> invert_f64(const float* W, int n);   // fp64 Gauss-Jordan inverse (srflow_engine.cu)
float* to_device(const std::vector<float>& v);
// UNet pieces shared by both prior variants (unet_engine.cu)
void pack_dense5(const Weights& W, const std::string& p, int nf, int gc, int out_dim, ConvW* out5, int* nf_pad_out);
View run_dense5(const ConvW* dense, int nf, int nf_pad, int gc, Arena& A, const View& x, cudaStream_t s);
View run_unet_body(const UNetBranchW& B, int depth, int dim, Arena& A, const View& x, cudaStream_t s);

}  // namespace bfsr

struct bfsr_unet {
  bfsr_unet_desc_t d;
  int device = 0;
  long long serial = 0;       // unique per created handle: captured launch plans of the engines key on it (an address can be reused)
  std::vector<bfsr::UNetBranchW> br;
  bfsr::Arena arena;          // used only by the standalone forward entry point
  ~bfsr_unet();
};

struct bfsr_linf {
  bfsr_linf_desc_t d;
  int device = 0;
  bfsr::ConvW head;                     // EDSR head / RRDB conv_first (3 -> 64)
  std::vector<bfsr::ConvW> body;        // EDSR: 2*16 + 1 ; RRDB: nb*15 + trunk_conv
  bfsr::ConvW cf;                       // coef | freq (64 -> 2*hidden)
  bfsr::ConvW mlp[4];                   // 1x1: 4*hidden -> hidden -> hidden -> hidden -> 2*D*flow_layers
  float* phase = nullptr;               // (hidden/2, 2)
  float* Mf = nullptr; float* Mi = nullptr; float* fbias = nullptr;   // flow: W_i, W_i^-1, b_i  (index flow_layers = last)
  bfsr::Arena arena;
  bfsr::GraphCache graphs;
  float* stage_in = nullptr; size_t stage_in_sz = 0;
  ~bfsr_linf();
};

struct bfsr_srflow {
  bfsr_srflow_desc_t d;
  int device = 0;
  int n_cond = 0;
  bfsr::RRDBW rrdb;
  std::vector<bfsr::LayerW> layers;
  std::vector<bfsr::LevelW> levels;   // index 1..L
  std::vector<int> latent_C, latent_level;
  bfsr::Arena arena;
  bfsr::GraphCache graphs;
  float* stage_in = nullptr; float* stage_out = nullptr; size_t stage_in_sz = 0, stage_out_sz = 0;
  ~bfsr_srflow();
};
