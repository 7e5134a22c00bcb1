// SRFlow-LP inference engine: weight packing and the execution plan of
// encoder -> feature-only coupling convs (shared by both flow directions) -> flow encode ->
// latent normalisation -> learned prior -> flow decode.
//
// Restructurings relative to the reference graph (all exact in real arithmetic):
//  * ActNorm after flow.Conv2d, the exp(3*logs) of Conv2dZeros and eval-mode BatchNorm are folded
//    into the conv weights at load time (flow.py:41-83, unet.py:38-56);
//  * fAffine's first conv is split by linearity into a z-part and an ft-part
//    (FlowAffineCouplingsAblation.py:114-116): the ft-part and all of fFeatures depend only on the LR
//    encoder output, so they are computed ONCE per level, batched over the level's K steps, and reused
//    by encode and decode (the reference recomputes them, and the encoder, in each pass);
//  * inverse(W) of every InvertibleConv1x1 is computed once in fp64 (Permutations.py:41 does it per call);
//  * dead encoder heads (upconv2/HRconv/conv_last for 4x, RRDBNet_arch.py:108-125) and the logdet chain
//    (discarded by get_sr/get_encode_z, SRFlow_model.py:199,204) are not computed;
//  * torch.cat never materialises: producers write channel slices of preallocated NHWC buffers.
#include "engine.cuh"
#include <cmath>
#include <cstring>
#include <cstdlib>

using namespace bfsr;

bfsr_srflow::~bfsr_srflow() {
  free_conv(rrdb.conv_first); free_conv(rrdb.trunk_conv);
  for (auto& c : rrdb.rdb) free_conv(c);
  for (auto& c : rrdb.upconv) free_conv(c);
  for (auto& l : layers) {
    if (l.step.Mf) cudaFree(l.step.Mf);
    if (l.step.cf) cudaFree(l.step.cf);
    if (l.step.Mi) cudaFree(l.step.Mi);
    if (l.step.ci) cudaFree(l.step.ci);
    if (l.step.MfT) cudaFree(l.step.MfT);
    if (l.step.MiT) cudaFree(l.step.MiT);
    free_conv(l.cp.fF2); free_conv(l.cp.fF4); free_conv(l.cp.fA0z); free_conv(l.cp.fA2); free_conv(l.cp.fA4);
    free_fused_coupling(l.cp.fz); free_fused_coupling(l.cp.ftail); free_fused_coupling(l.cp.ftail2);
    free_conv(l.split_conv);
  }
  for (auto& l : levels) {
    free_conv(l.fF0_all); free_conv(l.fA0ft_all);
    free_conv(l.fF0_1p); free_conv(l.fA0ft_1p);
  }
  if (stage_in) cudaFree(stage_in);
  if (stage_out) cudaFree(stage_out);
}

namespace bfsr {

float* to_device(const std::vector<float>& v) {
  float* d = nullptr;
  CUDA_OK(cudaMalloc((void**)&d, v.size() * 4));
  CUDA_OK(cudaMemcpy(d, v.data(), v.size() * 4, cudaMemcpyHostToDevice));
  return d;
}

// fp64 Gauss-Jordan inverse with partial pivoting (stands in for torch.inverse(W.double()), Permutations.py:41)
std::vector<double> invert_f64(const float* W, int n) {
  std::vector<double> a((size_t)n * 2 * n, 0.0);
  for (int i = 0; i < n; ++i) { for (int j = 0; j < n; ++j) a[(size_t)i * 2 * n + j] = W[i * n + j]; a[(size_t)i * 2 * n + n + i] = 1.0; }
  for (int c = 0; c < n; ++c) {
    int piv = c; double best = std::fabs(a[(size_t)c * 2 * n + c]);
    for (int r = c + 1; r < n; ++r) { double v = std::fabs(a[(size_t)r * 2 * n + c]); if (v > best) { best = v; piv = r; } }
    BFSR_CHECK(best > 1e-300, "InvertibleConv1x1 weight is singular");
    if (piv != c) for (int j = 0; j < 2 * n; ++j) std::swap(a[(size_t)c * 2 * n + j], a[(size_t)piv * 2 * n + j]);
    const double d = 1.0 / a[(size_t)c * 2 * n + c];
    for (int j = 0; j < 2 * n; ++j) a[(size_t)c * 2 * n + j] *= d;
    for (int r = 0; r < n; ++r) if (r != c) {
      const double f = a[(size_t)r * 2 * n + c];
      if (f != 0.0) for (int j = 0; j < 2 * n; ++j) a[(size_t)r * 2 * n + j] -= f * a[(size_t)c * 2 * n + j];
    }
  }
  std::vector<double> inv((size_t)n * n);
  for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) inv[(size_t)i * n + j] = a[(size_t)i * 2 * n + n + j];
  return inv;
}

static StepW pack_step(const Weights& W, const std::string& p, int C, bool coupling) {
  StepW s; s.C = C; s.coupling = coupling;
  const float* b = W.data(p + ".actnorm.bias", {1, C, 1, 1});
  const float* l = W.data(p + ".actnorm.logs", {1, C, 1, 1});
  const float* w = W.data(p + ".invconv.weight", {C, C});
  std::vector<float> Mf((size_t)C * C), cf(C), Mi((size_t)C * C), ci(C);
  for (int o = 0; o < C; ++o) {
    double acc = 0.0;
    for (int i = 0; i < C; ++i) {
      const double e = std::exp((double)l[i]);
      Mf[(size_t)o * C + i] = (float)((double)w[o * C + i] * e);
      acc += (double)w[o * C + i] * (double)b[i] * e;
    }
    cf[o] = (float)acc;
  }
  std::vector<double> inv = invert_f64(w, C);
  for (int o = 0; o < C; ++o) {
    const double e = std::exp(-(double)l[o]);
    for (int i = 0; i < C; ++i) Mi[(size_t)o * C + i] = (float)(inv[(size_t)o * C + i] * e);
    ci[o] = b[o];
  }
  s.Mf = to_device(Mf); s.cf = to_device(cf); s.Mi = to_device(Mi); s.ci = to_device(ci);
  std::vector<float> MfT((size_t)C * C), MiT((size_t)C * C);
  for (int o = 0; o < C; ++o) for (int i = 0; i < C; ++i) { MfT[(size_t)i * C + o] = Mf[(size_t)o * C + i]; MiT[(size_t)i * C + o] = Mi[(size_t)o * C + i]; }
  s.MfT = to_device(MfT); s.MiT = to_device(MiT);
  s.hMf = Mf; s.hcf = cf; s.hMi = Mi; s.hci = ci;
  return s;
}

// flow.Conv2d (bias-free conv + ActNorm): W' = W*e^logs, b' = b_an*e^logs   (flow.py:41-65)
static ConvW pack_conv_actnorm(const Weights& W, const std::string& p, int cout, int cin_src, int ks,
                               const std::vector<int>& cin_map, bool with_bias, int tc_min_cin = -1) {
  const float* w = W.data(p + ".weight", {cout, cin_src, ks, ks});
  const float* ab = W.data(p + ".actnorm.bias", {1, cout, 1, 1});
  const float* al = W.data(p + ".actnorm.logs", {1, cout, 1, 1});
  std::vector<float> sc(cout), bi(cout);
  for (int i = 0; i < cout; ++i) { sc[i] = std::exp(al[i]); bi[i] = with_bias ? ab[i] * sc[i] : 0.f; }
  return pack_conv(w, cout, cin_src, ks, bi.data(), sc.data(), cin_map, tc_min_cin);
}
// flow.Conv2dZeros: (conv + b) * exp(3*logs)   (flow.py:68-83)
static ConvW pack_conv_zeros(const Weights& W, const std::string& p, int cout, int cin) {
  const float* w = W.data(p + ".weight", {cout, cin, 3, 3});
  const float* b = W.data(p + ".bias", {cout});
  const float* l = W.data(p + ".logs", {cout, 1, 1});
  std::vector<float> sc(cout), bi(cout);
  for (int i = 0; i < cout; ++i) { sc[i] = std::exp(3.f * l[i]); bi[i] = b[i] * sc[i]; }
  return pack_conv(w, cout, cin, 3, bi.data(), sc.data(), {});
}
static ConvW pack_plain(const Weights& W, const std::string& p, int cout, int cin, int ks) {
  return pack_conv(W.data(p + ".weight", {cout, cin, ks, ks}), cout, cin, ks, W.data(p + ".bias", {cout}), nullptr, {});
}

static void build_srflow(bfsr_srflow* e, const Weights& W) {
  const bfsr_srflow_desc_t& d = e->d;
  BFSR_CHECK(d.scale == 4 || d.scale == 8, "scale %d unsupported (4 or 8)", d.scale);
  BFSR_CHECK(d.nf == 64 && d.gc == 32, "nf=%d gc=%d unsupported", d.nf, d.gc);
  BFSR_CHECK(d.n_blocks >= 0 && d.n_blocks <= 8, "too many stackRRDB blocks");
  const int nf = d.nf, gc = d.gc, Hd = d.hidden;
  e->n_cond = (d.n_blocks + 1) * nf;
  BFSR_CHECK(e->n_cond == 320, "conditioning width %d: CondAffineSeparatedAndCond hard-codes 320 "
             "(FlowAffineCouplingsAblation.py:30)", e->n_cond);
  // ---- encoder
  e->rrdb.conv_first = pack_plain(W, "RRDB.conv_first", nf, 3, 3);
  for (int i = 0; i < d.nb; ++i)
    for (int r = 1; r <= 3; ++r)
      for (int c = 1; c <= 5; ++c) {
        const std::string p = "RRDB.RRDB_trunk." + std::to_string(i) + ".RDB" + std::to_string(r) + ".conv" + std::to_string(c);
        e->rrdb.rdb.push_back(pack_plain(W, p, c < 5 ? gc : nf, nf + (c - 1) * gc, 3));
      }
  e->rrdb.trunk_conv = pack_plain(W, "RRDB.trunk_conv", nf, nf, 3);
  int log2s = d.scale == 4 ? 2 : 3;
  const int n_up = log2s - 1;   // level 1 runs at scale/2: upconv1 (.. upconv2 for 8x)
  for (int u = 1; u <= n_up; ++u) e->rrdb.upconv.push_back(pack_plain(W, "RRDB.upconv" + std::to_string(u), nf, nf, 3));
  // ---- flow topology (FlowUpsamplerNet.py:94-115)
  e->levels.resize(d.L + 1);
  int C = 3, idx = 0;
  for (int level = 1; level <= d.L; ++level) {
    C *= 4;
    { LayerW l; l.kind = 0; l.C = C; l.level = level; e->layers.push_back(l); ++idx; }
    for (int k = 0; k < d.n_no_affine; ++k) {
      LayerW l; l.kind = 1; l.C = C; l.level = level;
      l.step = pack_step(W, "flowUpsamplerNet.layers." + std::to_string(idx), C, false);
      e->layers.push_back(l); ++idx;
    }
    LevelW& lv = e->levels[level];
    lv.C = C; lv.n_coupling = d.K;
    std::vector<float> wF((size_t)d.K * Hd * 320 * 9), sF(d.K * Hd), bF(d.K * Hd);
    std::vector<float> wA((size_t)d.K * Hd * 320 * 9), sA(d.K * Hd), bA(d.K * Hd);
    for (int k = 0; k < d.K; ++k) {
      LayerW l; l.kind = 2; l.C = C; l.level = level; l.k_in_level = k;
      const std::string p = "flowUpsamplerNet.layers." + std::to_string(idx);
      l.step = pack_step(W, p, C, true);
      const int Cn = C / 2, Cc = C - Cn;
      // fFeatures.0 (ft only) -> batched
      {
        const float* w = W.data(p + ".affine.fFeatures.0.weight", {Hd, 320, 3, 3});
        const float* ab = W.data(p + ".affine.fFeatures.0.actnorm.bias", {1, Hd, 1, 1});
        const float* al = W.data(p + ".affine.fFeatures.0.actnorm.logs", {1, Hd, 1, 1});
        memcpy(&wF[(size_t)k * Hd * 320 * 9], w, (size_t)Hd * 320 * 9 * 4);
        for (int i = 0; i < Hd; ++i) { sF[k * Hd + i] = std::exp(al[i]); bF[k * Hd + i] = ab[i] * sF[k * Hd + i]; }
      }
      // fAffine.0: input = [z1 (Cn) | ft (320)]  (FlowAffineCouplingsAblation.py:115)
      {
        const float* w = W.data(p + ".affine.fAffine.0.weight", {Hd, Cn + 320, 3, 3});
        const float* ab = W.data(p + ".affine.fAffine.0.actnorm.bias", {1, Hd, 1, 1});
        const float* al = W.data(p + ".affine.fAffine.0.actnorm.logs", {1, Hd, 1, 1});
        for (int o = 0; o < Hd; ++o) {
          memcpy(&wA[((size_t)(k * Hd + o) * 320) * 9], &w[((size_t)o * (Cn + 320) + Cn) * 9], (size_t)320 * 9 * 4);
          sA[k * Hd + o] = std::exp(al[o]); bA[k * Hd + o] = ab[o] * sA[k * Hd + o];
        }
        // z-part input channels padded to a multiple of 8 (zero weights): the operand copy of z1 is a BF16X2 tensor whose
        // pixel stride must be 16-byte aligned for TMA
        const int Cnp = (Cn + 7) & ~7;
        std::vector<int> zmap(Cnp, -1); for (int i = 0; i < Cn; ++i) zmap[i] = i;
        l.cp.fA0z = pack_conv_actnorm(W, p + ".affine.fAffine.0", Hd, Cn + 320, 3, zmap, /*with_bias=*/false, /*tc_min_cin=*/1);
      }
      l.cp.fF2 = pack_conv_actnorm(W, p + ".affine.fFeatures.2", Hd, Hd, 1, {}, true);
      l.cp.fF4 = pack_conv_zeros(W, p + ".affine.fFeatures.4", 2 * C, Hd);
      l.cp.fA2 = pack_conv_actnorm(W, p + ".affine.fAffine.2", Hd, Hd, 1, {}, true);
      l.cp.fA4 = pack_conv_zeros(W, p + ".affine.fAffine.4", 2 * Cc, Hd);
      if (Hd == 64) { pack_fused_coupling(l.cp.fz, l.cp.fA0z, l.cp.fA2, l.cp.fA4, C); if (C == 12 || C == 24) pack_fused_tail(l.cp.ftail, l.cp.fF2, l.cp.fF4, 0); if (C == 24) pack_fused_tail(l.cp.ftail2, l.cp.fF2, l.cp.fF4, 24); }
      e->layers.push_back(l); ++idx;
    }
    lv.fF0_all = pack_conv(wF.data(), d.K * Hd, 320, 3, bF.data(), sF.data(), {});
    lv.fA0ft_all = pack_conv(wA.data(), d.K * Hd, 320, 3, bA.data(), sA.data(), {});
    // Level fed by [hi-res base 64 | nearest2x(taps of the next level)]: single-pass phase evaluation (1600/2880 of the MACs,
    // and the upsampled copy of the 256 tap channels is never materialised).  BFSR_PHASE=0 keeps the plain 3x3.
    static const bool use_phase = !(getenv("BFSR_PHASE") && atoi(getenv("BFSR_PHASE")) == 0);
    if (use_phase && level == log2s - 1) {   // conditioning = [upconv output | nearest2x(taps of the LR-resolution level)]
      lv.has_phase = true;
      lv.fF0_1p = pack_conv_tc_phase1(wF.data(), d.K * Hd, 320, 0, nf, nf, 320 - nf, sF.data(), bF.data());
      lv.fA0ft_1p = pack_conv_tc_phase1(wA.data(), d.K * Hd, 320, 0, nf, nf, 320 - nf, sA.data(), bA.data());
    }
    if (d.split_enable && level < d.L - 1) {
      LayerW l; l.kind = 3; l.C = C; l.level = level;
      const int cons = (int)std::lround(C * 0.5), pass = C - cons;
      l.split_conv = pack_conv_zeros(W, "flowUpsamplerNet.layers." + std::to_string(idx) + ".conv", 2 * cons, pass);
      e->layers.push_back(l); ++idx;
      e->latent_C.push_back(cons); e->latent_level.push_back(level);
      C = pass;
    }
  }
  e->latent_C.push_back(C); e->latent_level.push_back(d.L);
}

// ===================================================================== execution
struct Run {
  bfsr_srflow* e; cudaStream_t s; Arena& A;
  int B, h, w;
  std::vector<View> ft;                    // per level (index 1..L), 320 channels
  std::vector<View> bufA;                  // per level: n_coupling*64 pre-activations of fAffine.0 (ft part)
  std::vector<std::vector<View>> hF;       // per level, per step: (shiftFt, scaleFt) pairs, 2C channels
  bool plan() const { return A.plan; }
  int opfmt() const { return g_conv_mode == 2 ? (int)F32 : (int)BF16X2; }   // format of conv-operand-only tensors
  int lvH(int level) const { return (h * e->d.scale) >> level; }
  int lvW(int level) const { return (w * e->d.scale) >> level; }
};
#define K_(...) do { if (!r.A.plan) { __VA_ARGS__; } } while (0)

// RRDBNet.forward(get_steps=True) + SRFlowNet.rrdbPreprocessing (RRDBNet_arch.py:89-148, SRFlowNet_arch.py:118-138)
static void run_encoder(Run& r, const View& x) {
  bfsr_srflow* e = r.e; const auto& d = e->d;
  const int B = r.B, h = r.h, w = r.w, nf = d.nf, gc = d.gc, L = d.L;
  const int log2s = d.scale == 4 ? 2 : 3;
  // Tensors that are only ever consumed as conv operands live in HBM already split into bf16 (hi, lo) planes (BF16X2:
  // same bytes as fp32, bit-identical MMA operands) so the tensor-core convs fetch them by TMA with no conversion pass.
  const int fmt = r.opfmt();
  const bool split = fmt == BF16X2;
  r.ft.assign(L + 1, View());
  for (int lv = 1; lv <= L; ++lv) r.ft[lv] = make_view(r.A, B, r.lvH(lv), r.lvW(lv), e->n_cond, fmt);
  const int lv0 = log2s;                           // level whose features live at LR resolution ('fea_up1')
  BFSR_CHECK(lv0 <= L, "L=%d too small for scale %d", L, d.scale);
  View& ft0 = r.ft[lv0];
  const size_t mark = r.A.off;
  // Dense-block buffers: D[i] = [x (64) | conv1..4 outputs (4 x 32)] in operand format; the residual stream x also keeps
  // an fp32 copy X[i] (the 69 chained `x5*0.2 + x` updates must not be re-rounded to 16 mantissa bits each time).
  View D[3], X[3];
  for (int i = 0; i < 3; ++i) {
    D[i] = make_view(r.A, B, h, w, nf + 4 * gc, fmt);
    X[i] = split ? make_view(r.A, B, h, w, nf) : D[i].slice(0, nf);
  }
  ConvEpi lrelu; lrelu.act = ACT_LRELU;
  {
    ConvEpi ep; View op0 = D[0].slice(0, nf); if (split) ep.out2 = &op0;
    K_(conv2d(e->rrdb.conv_first, x, X[0], ep, IN_DIRECT, r.s));
  }
  int tap = 0;
  for (int i = 0; i < d.nb; ++i) {
    for (int rb = 0; rb < 3; ++rb) {
      View& cur = D[rb]; const int nx = (rb + 1) % 3;
      const ConvW* cw = &e->rrdb.rdb[(size_t)(i * 3 + rb) * 5];
      for (int c = 0; c < 4; ++c)
        K_(conv2d(cw[c], cur.slice(0, nf + c * gc), cur.slice(nf + c * gc, gc), lrelu, IN_DIRECT, r.s));
      ConvEpi ep;
      View op_next = D[nx].slice(0, nf); if (split) ep.out2 = &op_next;
      if (rb < 2) { ep.alpha = 0.2f; ep.res1 = &X[rb]; ep.beta1 = 1.f; }                      // x5*0.2 + x
      else { ep.alpha = 0.04f; ep.res1 = &X[rb]; ep.beta1 = 0.2f; ep.res2 = &X[0]; ep.beta2 = 1.f; }  // (x5*0.2+x)*0.2 + x_rrdb
      K_(conv2d(cw[4], cur.slice(0, nf + 4 * gc), X[nx], ep, IN_DIRECT, r.s));
    }
    for (int b = 0; b < d.n_blocks; ++b)
      if (d.blocks[b] == i) {   // block_{i} tap -> its slot in the conditioning tensor (order of stackRRDB.blocks)
        K_(resample(X[0], ft0.slice(nf * (1 + b), nf), RS_COPY, r.s));
        ++tap;
      }
  }
  BFSR_CHECK(tap == d.n_blocks, "stackRRDB.blocks reference RRDB indices outside [0, nb)");
  {  // last_lr_fea = fea + trunk_conv(fea)
    ConvEpi ep; ep.res1 = &X[0]; ep.beta1 = 1.f;
    K_(conv2d(e->rrdb.trunk_conv, D[0].slice(0, nf), ft0.slice(0, nf), ep, IN_DIRECT, r.s));
  }
  r.A.off = mark;
  // finer levels: fea_up2 = lrelu(upconv1(nearest2x(last_lr_fea))) etc. (post-activation: in-place LeakyReLU aliasing)
  for (int lv = lv0 - 1, u = 0; lv >= 1; --lv, ++u) {
    K_(conv2d(e->rrdb.upconv[u], r.ft[lv + 1].slice(0, nf), r.ft[lv].slice(0, nf), lrelu, IN_UP2, r.s));
    // the phase path reads the low-res taps directly, so the upsampled copy is only needed by the plain path -- or by a
    // still finer level (8x: level 1 upsamples level 2's taps once more)
    if (!(e->levels[lv].has_phase && g_conv_mode != 2) || lv > 1)
      K_(resample(r.ft[lv + 1].slice(nf, e->n_cond - nf), r.ft[lv].slice(nf, e->n_cond - nf), RS_NEAREST_UP2, r.s));
  }
  // coarser level: fea_up0 = bilinear x0.5 (== 2x2 mean), taps nearest x0.5
  for (int lv = lv0 + 1; lv <= L; ++lv) {
    BFSR_CHECK(lv == lv0 + 1, "levels below fea_up0 are not supported");
    K_(resample(ft0.slice(0, nf), r.ft[lv].slice(0, nf), RS_AVG_DOWN2, r.s));
    K_(resample(ft0.slice(nf, e->n_cond - nf), r.ft[lv].slice(nf, e->n_cond - nf), RS_NEAREST_DOWN2, r.s));
  }
}

static bool fp32_z() { static const bool v = getenv("BFSR_FP32_Z") && atoi(getenv("BFSR_FP32_Z")); return v; }
// one-launch feature-only tail (coupling_fused.cu, TAIL variant): tensor-core modes, C = 12 levels (C = 24: two launches of 24 outputs)
static bool tail_one(const Run& r, const LayerW& l) {
  return coupling_fused_enabled() && g_conv_mode != 2 && r.opfmt() == BF16X2 && l.cp.ftail.w != nullptr && (l.C == 12 || l.cp.ftail2.w != nullptr);
}
// feature-only halves of every coupling step of every level
static void run_ft_convs(Run& r) {
  bfsr_srflow* e = r.e; const auto& d = e->d;
  const int Hd = d.hidden;
  r.bufA.assign(d.L + 1, View());
  r.hF.assign(d.L + 1, {});
  for (int lv = 1; lv <= d.L; ++lv) {
    const LevelW& L = e->levels[lv];
    const int H = r.lvH(lv), W = r.lvW(lv);
    // pre-activations of fAffine.0 (ft part): consumed as identity K chunks of the z-dependent conv -> operand format
    r.bufA[lv] = make_view(r.A, r.B, H, W, L.n_coupling * Hd, fp32_z() ? (int)F32 : r.opfmt());
    for (int k = 0; k < L.n_coupling; ++k) r.hF[lv].push_back(make_view(r.A, r.B, H, W, 2 * L.C));
  }
  for (int lv = 1; lv <= d.L; ++lv) {
    const LevelW& L = e->levels[lv];
    const int H = r.lvH(lv), W = r.lvW(lv);
    const size_t mark = r.A.off;
    View bufF = make_view(r.A, r.B, H, W, L.n_coupling * Hd, r.opfmt());
    View t = make_view(r.A, r.B, H, W, Hd, r.opfmt());
    ConvEpi relu; relu.act = ACT_RELU;
    ConvEpi cs; cs.act = ACT_CROSS_SIGMOID;
    if (L.has_phase && g_conv_mode != 2) {
      const int nf = d.nf;
      View hi = r.ft[lv].slice(0, nf), lo = r.ft[lv + 1].slice(nf, e->n_cond - nf);
      K_(conv2d_tc_phase1(L.fF0_1p, hi, lo, bufF, relu, r.s));
      K_(conv2d_tc_phase1(L.fA0ft_1p, hi, lo, r.bufA[lv], ConvEpi(), r.s));
    } else {
      K_(conv2d(L.fF0_all, r.ft[lv], bufF, relu, IN_DIRECT, r.s));
      K_(conv2d(L.fA0ft_all, r.ft[lv], r.bufA[lv], ConvEpi(), IN_DIRECT, r.s));
    }
    for (const LayerW& l : e->layers) {
      if (l.kind != 2 || l.level != lv) continue;
      if (tail_one(r, l)) {
        const View in = bufF.slice(l.k_in_level * Hd, Hd), &o = r.hF[lv][l.k_in_level];
        K_(tail_fused(l.cp.ftail, in, o.slice(0, 24), cs.eps, r.s));
        if (l.C == 24) K_(tail_fused(l.cp.ftail2, in, o.slice(24, 24), cs.eps, r.s));
        continue;
      }
      K_(conv2d(l.cp.fF2, bufF.slice(l.k_in_level * Hd, Hd), t, relu, IN_DIRECT, r.s));
      K_(conv2d(l.cp.fF4, t, r.hF[lv][l.k_in_level], cs, IN_DIRECT, r.s));
    }
    r.A.off = mark;
  }
}

// z-dependent half of fAffine: h = (shift, scale) pairs for z2   (FlowAffineCouplingsAblation.py:114-119)
static bool z1_fused(const Run& r) { return !fp32_z() && r.opfmt() == BF16X2; }   // step kernels emit the z1 operand copy
// Levels whose FlowStep is fused into the epilogue of the coupling's last conv (C = 12, 24): h never reaches HBM and the
// step costs no launch of its own.  BFSR_FUSE_FLOW=0 keeps the separate step kernels.
static bool flow_fused(const Run& r, int C) {
  static const bool off = getenv("BFSR_FUSE_FLOW") && atoi(getenv("BFSR_FUSE_FLOW")) == 0;
  return !off && z1_fused(r) && (C == 12 || C == 24);
}
// Per-level scratch: the flow state ping-pongs between two buffers; the affine-net intermediates are reused by
// every step of the level (all work is ordered on one stream).
struct LevelBufs {
  View z[2], z1op, t1, t2, h; int pp = 0;
  void* z1p[2] = {nullptr, nullptr};   // fused coupling steps: z1 operand planes [hi(8) | lo(8)] bf16 per pixel, read with a halo while the
  int z1p_cur = 0;                     // next step's plane is written, hence two
  bool z1p_ready = false;              // z1p[z1p_cur] holds the operand of the current flow state
  void alloc(Run& r, int H, int W, int C) {
    const int Hd = r.e->d.hidden;
    z[0] = make_view(r.A, r.B, H, W, C); z[1] = make_view(r.A, r.B, H, W, C);
    t1 = make_view(r.A, r.B, H, W, Hd, fp32_z() ? (int)F32 : r.opfmt()); t2 = make_view(r.A, r.B, H, W, Hd, r.opfmt());
    h = make_view(r.A, r.B, H, W, (C - C / 2) * 2);
    z1op = make_view(r.A, r.B, H, W, (C / 2 + 7) & ~7, r.opfmt());
    if (C == 12 || C == 24) for (int i = 0; i < 2; ++i) z1p[i] = r.A.alloc((size_t)r.B * H * W * (C == 12 ? 32 : 64));
    pp = 0; z1p_cur = 0; z1p_ready = false;
  }
  View& next() { View& v = z[pp]; pp ^= 1; return v; }
};

// One-launch coupling step (coupling_fused.cu): tensor-core modes (accurate and bf16 single-pass), C = 12 / 24 levels, FlowStep fused (BFSR_FUSE_CPL=0: three launches)
static bool cpl_fused(const Run& r, const LayerW& l) {
  return coupling_fused_enabled() && g_conv_mode != 2 && l.cp.fz.w != nullptr && flow_fused(r, l.C);
}
// Returns true when the fused kernel ran (the z1 operand of the next step, if requested, then lives in lb.z1p, not in lb.z1op).
static bool run_affine_net(Run& r, const LayerW& l, const View& z, LevelBufs& lb, bool z1_ready, const FlowEpi* flow = nullptr) {
  const int Hd = r.e->d.hidden;
  const View &z1op = lb.z1op, &t1 = lb.t1, &t2 = lb.t2, &hout = lb.h;
  View pre = r.bufA[l.level].slice(l.k_in_level * Hd, Hd);
  if (flow && cpl_fused(r, l)) {
    if (!lb.z1p_ready) K_(z1_pack(z.slice(0, l.C / 2), lb.z1p[lb.z1p_cur], l.C, r.s));
    K_(coupling_fused(l.cp.fz, lb.z1p[lb.z1p_cur], lb.z1p[lb.z1p_cur ^ 1], pre, *flow, flow->hM, flow->hcvec, ConvEpi().eps, r.s));
    lb.z1p_ready = flow->z1op.p != nullptr;
    if (lb.z1p_ready) lb.z1p_cur ^= 1;
    return true;
  }
  lb.z1p_ready = false;
  ConvEpi e1; e1.act = ACT_RELU; e1.pre = &pre;
  // z-dependent first conv: split-bf16 x3 on the tensor cores like every other conv (products exact to ~2^-17); BFSR_FP32_Z=1
  // keeps it on the fp32 CUDA-core kernel (the round-1 default) for A/B parity runs
  const int Cnp = l.cp.fA0z.cin;                    // C/2 padded to a multiple of 8 (the padding weights are zero)
  if (fp32_z() || r.opfmt() != BF16X2) K_(conv2d_fp32(l.cp.fA0z, z.slice(0, Cnp), t1, e1, IN_DIRECT, r.s));
  else {
    // operand copy of z1 in bf16 (hi, lo) planes so the conv is TMA-fed like every other one (no register producers)
    if (!z1_ready) K_(resample(z.slice(0, l.C / 2), z1op, RS_COPY, r.s));
    K_(conv2d(l.cp.fA0z, z1op, t1, e1, IN_DIRECT, r.s));
  }
  ConvEpi relu; relu.act = ACT_RELU;
  K_(conv2d(l.cp.fA2, t1, t2, relu, IN_DIRECT, r.s));
  ConvEpi cs; cs.act = ACT_CROSS_SIGMOID; cs.flow = flow;
  if (flow) { View none = hout; none.p = nullptr; K_(conv2d_tc(l.cp.fA4, t2, none, cs, IN_DIRECT, r.s)); }
  else K_(conv2d(l.cp.fA4, t2, hout, cs, IN_DIRECT, r.s));
  return false;
}

// FlowUpsamplerNet.encode (FlowUpsamplerNet.py:217-251).  gt: (B, sh, sw, 3) NHWC.  Returns latent views.
static std::vector<View> run_encode(Run& r, const View& gt) {
  bfsr_srflow* e = r.e;
  std::vector<View> lat;
  View z = gt;                    // current flow state
  bool pending = false;           // coupling whose second half has not been applied to z yet
  bool premixed = false;          // the previous coupling's fused epilogue already applied this step's actnorm/invconv/ft-affine
  LevelBufs lb; int cur_level = 0;
  for (size_t i = 0; i < e->layers.size(); ++i) {
    const LayerW& l = e->layers[i];
    if (l.kind == 0) continue;    // squeeze: folded into the next step's load
    const int H = r.lvH(l.level), W = r.lvW(l.level);
    if (l.kind == 1 || l.kind == 2) {
      if (l.level != cur_level) { lb.alloc(r, H, W, l.C); cur_level = l.level; }
      const bool sq = e->layers[i - 1].kind == 0;
      BFSR_CHECK(!(sq && pending), "internal: pending coupling across a squeeze");
      const View* hF = l.kind == 2 ? &r.hF[l.level][l.k_in_level] : nullptr;
      const bool emit_z1 = l.kind == 2 && z1_fused(r) && !cpl_fused(r, l);     // this step's output feeds its own affine net (the one-launch coupling packs its own operand)
      const bool level_end = (i + 1 == e->layers.size()) || e->layers[i + 1].kind == 0 || e->layers[i + 1].kind == 3;
      if (!premixed) {
        View zo = lb.next();
        K_(flowstep_fwd(l.step, z, sq, pending ? &lb.h : nullptr, hF, zo, r.s, emit_z1 ? &lb.z1op : nullptr));
        pending = false;
        z = zo;
        lb.z1p_ready = false;
      }
      const bool z1_have = emit_z1;      // premixed steps received their z1 copy from the previous coupling's epilogue
      premixed = false;
      if (l.kind == 2 && flow_fused(r, l.C)) {
        // coupling of this step + actnorm / invconv / ft-affine of the next step run in the epilogue of fAffine's last conv
        FlowEpi f; f.inv = 0; f.C = l.C; f.z_in = z; f.z_out = lb.next();
        if (!level_end) {
          const LayerW& nx = e->layers[i + 1];
          BFSR_CHECK(nx.kind == 2 && nx.level == l.level, "internal: coupling followed by a non-coupling step inside a level");
          f.has_mix = 1; f.M = nx.step.Mf; f.cvec = nx.step.cf; f.hF = r.hF[nx.level][nx.k_in_level]; f.z1op = lb.z1op;
          f.hM = nx.step.hMf.data(); f.hcvec = nx.step.hcf.data();
        } else f.has_mix = 0;
        run_affine_net(r, l, z, lb, z1_have, &f);
        z = f.z_out;
        premixed = !level_end;
      } else {
        if (l.kind == 2) { run_affine_net(r, l, z, lb, z1_have); pending = true; }
        if (level_end && pending) { K_(coupling_finish(z, lb.h, z, r.s)); pending = false; }   // in place
      }
    } else {   // Split2d forward (Split.py:49-61)
      const int cons = (int)std::lround(l.C * 0.5), pass = l.C - cons;
      View hs = make_view(r.A, r.B, H, W, 2 * cons);
      View z1 = make_view(r.A, r.B, H, W, pass), eps = make_view(r.A, r.B, H, W, cons);
      K_(conv2d_fp32(l.split_conv, z.slice(0, pass), hs, ConvEpi(), IN_DIRECT, r.s));
      K_(split_fwd(z, hs, z1, eps, r.s));
      lat.push_back(eps);
      z = z1;
    }
  }
  lat.push_back(z);
  return lat;
}

// FlowUpsamplerNet.decode (FlowUpsamplerNet.py:267-296).  Returns SR as NHWC (B, sh, sw, 3).
static View run_decode(Run& r, const std::vector<View>& lat) {
  bfsr_srflow* e = r.e;
  int li = (int)lat.size() - 1;
  View z = lat[li--];
  LevelBufs lb; int cur_level = 0;
  bool z1_ready = false;
  for (int i = (int)e->layers.size() - 1; i >= 0; --i) {
    const LayerW& l = e->layers[i];
    if (l.kind == 0) continue;   // unsqueeze: folded into the previous step's store
    const int H = r.lvH(l.level), W = r.lvW(l.level);
    if (l.kind == 3) {
      const int cons = (int)std::lround(l.C * 0.5);
      BFSR_CHECK(li >= 0, "decode: not enough latents");
      View hs = make_view(r.A, r.B, H, W, 2 * cons);
      View zo = make_view(r.A, r.B, H, W, l.C);
      K_(conv2d_fp32(l.split_conv, z, hs, ConvEpi(), IN_DIRECT, r.s));
      K_(split_inv(z, hs, lat[li], zo, r.s));
      --li; z = zo; z1_ready = false; lb.z1p_ready = false;
      continue;
    }
    if (l.level != cur_level) { lb.alloc(r, H, W, l.C); cur_level = l.level; }
    const bool unsq = e->layers[i - 1].kind == 0;
    View zo = unsq ? make_view(r.A, r.B, 2 * H, 2 * W, l.C / 4) : lb.next();
    // the step's output is the input of the next processed layer: if that is a coupling of the same level, emit its z1 copy
    const bool next_cpl = i > 0 && e->layers[i - 1].kind == 2 && e->layers[i - 1].level == l.level && !unsq && z1_fused(r);
    // a one-launch coupling next packs / receives its operand in lb.z1p: the step kernels need not emit the BF16X2 copy for it
    const bool next_one = next_cpl && cpl_fused(r, e->layers[i - 1]) && !(i > 1 && e->layers[i - 2].kind == 0);
    bool one = false;
    if (l.kind == 2 && flow_fused(r, l.C) && !unsq) {
      FlowEpi f; f.inv = 1; f.C = l.C; f.z_in = z; f.z_out = zo; f.has_mix = 1; f.M = l.step.Mi; f.cvec = l.step.ci;
      f.hM = l.step.hMi.data(); f.hcvec = l.step.hci.data();
      f.hF = r.hF[l.level][l.k_in_level];
      if (next_cpl) f.z1op = lb.z1op;
      one = run_affine_net(r, l, z, lb, z1_ready, &f);
    } else if (l.kind == 2) {
      run_affine_net(r, l, z, lb, z1_ready);
      K_(flowstep_inv(l.step, z, &lb.h, &r.hF[l.level][l.k_in_level], zo, unsq, r.s, next_cpl && !next_one ? &lb.z1op : nullptr));
    } else {
      K_(flowstep_inv(l.step, z, nullptr, nullptr, zo, unsq, r.s, next_cpl && !next_one ? &lb.z1op : nullptr));
      lb.z1p_ready = false;
    }
    z1_ready = next_cpl && !one && !next_one;
    z = zo;
  }
  BFSR_CHECK(z.C == 3, "decode: final tensor has %d channels", z.C);
  return z;
}

}  // namespace bfsr

// ===================================================================== single modules (P1 parity tests, profiling harness)
namespace bfsr {

static void pack_coupling(const Weights& W, const std::string& p, int C, int Hd, LayerW& l, ConvW& fF0, ConvW& fA0ft) {
  const int Cn = C / 2, Cc = C - Cn;
  fF0 = pack_conv_actnorm(W, p + ".affine.fFeatures.0", Hd, 320, 3, {}, true);
  std::vector<int> ftmap(320); for (int i = 0; i < 320; ++i) ftmap[i] = Cn + i;
  fA0ft = pack_conv_actnorm(W, p + ".affine.fAffine.0", Hd, Cn + 320, 3, ftmap, true);
  const int Cnp = (Cn + 7) & ~7;
  std::vector<int> zmap(Cnp, -1); for (int i = 0; i < Cn; ++i) zmap[i] = i;
  l.cp.fA0z = pack_conv_actnorm(W, p + ".affine.fAffine.0", Hd, Cn + 320, 3, zmap, /*with_bias=*/false, /*tc_min_cin=*/1);
  l.cp.fF2 = pack_conv_actnorm(W, p + ".affine.fFeatures.2", Hd, Hd, 1, {}, true);
  l.cp.fF4 = pack_conv_zeros(W, p + ".affine.fFeatures.4", 2 * C, Hd);
  l.cp.fA2 = pack_conv_actnorm(W, p + ".affine.fAffine.2", Hd, Hd, 1, {}, true);
  l.cp.fA4 = pack_conv_zeros(W, p + ".affine.fAffine.4", 2 * Cc, Hd);
  if (Hd == 64) { pack_fused_coupling(l.cp.fz, l.cp.fA0z, l.cp.fA2, l.cp.fA4, C); if (C == 12 || C == 24) pack_fused_tail(l.cp.ftail, l.cp.fF2, l.cp.fF4, 0); if (C == 24) pack_fused_tail(l.cp.ftail2, l.cp.fF2, l.cp.fF4, 24); }
}

// One FlowStep of the reference (FlowStep.py:88-129) through exactly the kernels the engine uses for that step: the feature-only
// convs, the z-dependent affine net, and the step kernel / fused conv epilogue.  `reps` > 1 repeats the z-dependent part (the
// per-step cost inside a level) for profiling.
void op_flowstep(const bfsr_tensor_t* weights, int n, const char* prefix, int C, bool coupling, bool reverse, const float* z_nchw,
                 const float* ft_nchw, int B, int H, int Wd, float* out_nchw, int reps, cudaStream_t s) {
  Weights W(weights, n);
  const std::string p = prefix;
  bfsr_srflow e; memset(&e.d, 0, sizeof e.d); e.d.hidden = 64; e.d.scale = 4; e.d.L = 1;
  const int Hd = 64;
  LayerW l; l.kind = coupling ? 2 : 1; l.C = C; l.level = 1; l.k_in_level = 0;
  l.step = pack_step(W, p, C, coupling);
  ConvW fF0, fA0ft;
  if (coupling) pack_coupling(W, p, C, Hd, l, fF0, fA0ft);
  e.layers.push_back(l);     // owned (and freed) by e
  Arena& A = e.arena;
  for (int pass = 0; pass < 2; ++pass) {
    A.plan = pass == 0; A.reset();
    if (pass == 1) A.reserve(A.peak + (1 << 20));
    Run r{&e, s, A, B, H, Wd};
    View z = make_view(A, B, H, Wd, C), zo = make_view(A, B, H, Wd, C), zo2 = make_view(A, B, H, Wd, C);
    K_(nchw_to_nhwc(z_nchw, z, s));
    View hF, ft;
    r.bufA.assign(2, View()); r.hF.assign(2, {});
    if (coupling) {
      View ftf = make_view(A, B, H, Wd, 320);
      ft = make_view(A, B, H, Wd, 320, r.opfmt());
      K_(nchw_to_nhwc(ft_nchw, ftf, s));
      K_(resample(ftf, ft, RS_COPY, s));
      View bufF = make_view(A, B, H, Wd, Hd, r.opfmt()), t = make_view(A, B, H, Wd, Hd, r.opfmt());
      r.bufA[1] = make_view(A, B, H, Wd, Hd, fp32_z() ? (int)F32 : r.opfmt());
      hF = make_view(A, B, H, Wd, 2 * C);
      ConvEpi relu; relu.act = ACT_RELU;
      ConvEpi cs; cs.act = ACT_CROSS_SIGMOID;
      K_(conv2d(fF0, ft, bufF, relu, IN_DIRECT, s));
      K_(conv2d(fA0ft, ft, r.bufA[1], ConvEpi(), IN_DIRECT, s));
      if (tail_one(r, l)) {
        K_(tail_fused(l.cp.ftail, bufF, hF.slice(0, 24), cs.eps, s));
        if (l.C == 24) K_(tail_fused(l.cp.ftail2, bufF, hF.slice(24, 24), cs.eps, s));
      }
      else {
        K_(conv2d(l.cp.fF2, bufF, t, relu, IN_DIRECT, s));
        K_(conv2d(l.cp.fF4, t, hF, cs, IN_DIRECT, s));
      }
      r.hF[1].push_back(hF);
    }
    LevelBufs lb; lb.alloc(r, H, Wd, C);
    const LayerW& L = e.layers[0];
    for (int it = 0; it < (reps > 0 ? reps : 1); ++it) {
      if (!reverse) {
        const bool emit_z1 = coupling && z1_fused(r);
        K_(flowstep_fwd(L.step, z, false, nullptr, coupling ? &hF : nullptr, zo, s, emit_z1 ? &lb.z1op : nullptr));
        if (coupling && flow_fused(r, C)) {
          FlowEpi f; f.inv = 0; f.C = C; f.z_in = zo; f.z_out = zo2; f.has_mix = 0;
          lb.z1p_ready = false;
          run_affine_net(r, L, zo, lb, emit_z1, &f);
        } else if (coupling) {
          run_affine_net(r, L, zo, lb, emit_z1);
          K_(coupling_finish(zo, lb.h, zo2, s));
        } else K_(resample(zo, zo2, RS_COPY, s));
      } else {
        if (coupling && flow_fused(r, C)) {
          FlowEpi f; f.inv = 1; f.C = C; f.z_in = z; f.z_out = zo2; f.has_mix = 1; f.M = L.step.Mi; f.cvec = L.step.ci; f.hF = hF;
          f.hM = L.step.hMi.data(); f.hcvec = L.step.hci.data();
          lb.z1p_ready = false;
          run_affine_net(r, L, z, lb, false, &f);
        } else if (coupling) {
          run_affine_net(r, L, z, lb, false);
          K_(flowstep_inv(L.step, z, &lb.h, &hF, zo2, false, s, nullptr));
        } else K_(flowstep_inv(L.step, z, nullptr, nullptr, zo2, false, s, nullptr));
      }
    }
    K_(nhwc_to_nchw(zo2, out_nchw, s));
  }
  CUDA_OK(cudaStreamSynchronize(s));
  free_conv(fF0); free_conv(fA0ft);
  CUDA_OK(cudaGetLastError());
}

// Split2d of the reference (Split.py:49-77) through the engine's kernels.  forward: z (C) -> z1 (C - cons), eps (cons);
// reverse: z1, eps -> z.
void op_split2d(const bfsr_tensor_t* weights, int n, const char* prefix, int C, bool reverse, const float* z_nchw, const float* eps_nchw,
                int B, int H, int Wd, float* out_z, float* out_eps, cudaStream_t s) {
  Weights W(weights, n);
  const int cons = (int)std::lround(C * 0.5), pass = C - cons;
  ConvW cw = pack_conv_zeros(W, std::string(prefix) + ".conv", 2 * cons, pass);
  Arena A;
  try {
    for (int ps = 0; ps < 2; ++ps) {
      A.plan = ps == 0; A.reset();
      if (ps == 1) A.reserve(A.peak + (1 << 20));
      View hs = make_view(A, B, H, Wd, 2 * cons);
      if (!reverse) {
        View z = make_view(A, B, H, Wd, C), z1 = make_view(A, B, H, Wd, pass), eps = make_view(A, B, H, Wd, cons);
        if (!A.plan) {
          nchw_to_nhwc(z_nchw, z, s);
          conv2d_fp32(cw, z.slice(0, pass), hs, ConvEpi(), IN_DIRECT, s);
          split_fwd(z, hs, z1, eps, s);
          nhwc_to_nchw(z1, out_z, s); nhwc_to_nchw(eps, out_eps, s);
        }
      } else {
        View z1 = make_view(A, B, H, Wd, pass), eps = make_view(A, B, H, Wd, cons), zo = make_view(A, B, H, Wd, C);
        if (!A.plan) {
          nchw_to_nhwc(z_nchw, z1, s); nchw_to_nhwc(eps_nchw, eps, s);
          conv2d_fp32(cw, z1, hs, ConvEpi(), IN_DIRECT, s);
          split_inv(z1, hs, eps, zo, s);
          nhwc_to_nchw(zo, out_z, s);
        }
      }
    }
    CUDA_OK(cudaStreamSynchronize(s));
  } catch (...) { free_conv(cw); throw; }
  free_conv(cw);
}

}  // namespace bfsr

// ===================================================================== UNet prior (defined in unet_engine.cu)
namespace bfsr {
std::vector<View> run_unet_srflow(bfsr_unet* u, Arena& A, const std::vector<View>& lat, cudaStream_t s);
}

// ===================================================================== entry points used by capi.cu
namespace bfsr {

void srflow_build(bfsr_srflow* e, const bfsr_tensor_t* weights, int n) {
  Weights W(weights, n);
  build_srflow(e, W);
}

static void check_dims(const bfsr_srflow* e, int B, int h, int w) {
  BFSR_CHECK(B >= 0 && h > 0 && w > 0, "bad input shape (%d,3,%d,%d)", B, h, w);
  const int div = 1 << e->d.L;
  BFSR_CHECK((h * e->d.scale) % div == 0 && (w * e->d.scale) % div == 0,
             "LR size %dx%d: scale*size must be divisible by 2^L=%d (the reference pads to even, test.py:126-130)", h, w, div);
}

enum Mode { M_ENCODE, M_DECODE, M_LP };

// One chunk of tiles through the workspace.  Pointers are NCHW device buffers already offset to the chunk.
static void run_chunk(bfsr_srflow* e, bfsr_unet* prior, Mode mode, const float* lr, const float* gt,
                      float* const* lat_out, const float* const* lat_in, float* sr, int B, int h, int w,
                      cudaStream_t s) {
  Arena& A = e->arena;
  A.reset();
  Run r{e, s, A, B, h, w};
  const int S = e->d.scale;
  View x = make_view(A, B, h, w, 3);
  K_(nchw_to_nhwc(lr, x, s));
  run_encoder(r, x);
  run_ft_convs(r);
  std::vector<View> lat;
  const int nl = (int)e->latent_C.size();
  if (mode == M_ENCODE || mode == M_LP) {
    View g = make_view(A, B, S * h, S * w, 3);
    if (mode == M_LP) K_(bilinear_up_nchw(lr, B, 3, h, w, S, g, s));
    else K_(nchw_to_nhwc(gt, g, s));
    lat = run_encode(r, g);
    BFSR_CHECK((int)lat.size() == nl, "internal: latent count");
    if (mode == M_ENCODE) {
      for (int i = 0; i < nl; ++i) K_(nhwc_to_nchw(lat[i], lat_out[i], s));
      return;
    }
    std::vector<View> nrm;
    for (int i = 0; i < nl; ++i) {
      View o = make_view(A, B, lat[i].H, lat[i].W, lat[i].C);
      K_(normalise_latent(lat[i], o, s));
      nrm.push_back(o);
    }
    lat = run_unet_srflow(prior, A, nrm, s);
  } else {
    for (int i = 0; i < nl; ++i) {
      const int lv = e->latent_level[i];
      View v = make_view(A, B, r.lvH(lv), r.lvW(lv), e->latent_C[i]);
      K_(nchw_to_nhwc(lat_in[i], v, s));
      lat.push_back(v);
    }
  }
  View out = run_decode(r, lat);
  K_(nhwc_to_nchw(out, sr, s));
}

void srflow_run(bfsr_srflow* e, bfsr_unet* prior, int mode_i, const float* lr, const float* gt,
                float* const* lat_out, const float* const* lat_in, float* sr, int B, int h, int w, cudaStream_t s) {
  const Mode mode = (Mode)mode_i;
  check_dims(e, B, h, w);
  if (mode == M_LP) BFSR_CHECK(prior != nullptr && prior->d.variant == 0, "lp_sr needs an SRFlow-LP prior handle");
  if (mode == M_LP) {
    BFSR_CHECK(prior->d.n_latents == (int)e->latent_C.size(), "prior expects %d latents, flow produces %zu",
               prior->d.n_latents, e->latent_C.size());
    for (size_t i = 0; i < e->latent_C.size(); ++i)
      BFSR_CHECK(prior->d.latent_ch[i] == e->latent_C[i], "prior latent %zu has %d channels, flow produces %d", i,
                 prior->d.latent_ch[i], e->latent_C[i]);
  }
  if (B == 0) return;
  CUDA_OK(cudaSetDevice(e->device));
  // tiles per pass through the workspace: 32 by default, halved until the planned workspace fits 64 GB
  int chunk = e->d.tile_chunk > 0 ? e->d.tile_chunk : 32;
  const int S = e->d.scale;
  const int nl = (int)e->latent_C.size();
  {
    Arena& A = e->arena;
    const size_t peak0 = A.peak;
    size_t need = 0;
    for (;;) {
      A.plan = true; A.peak = 0;
      run_chunk(e, prior, mode, nullptr, nullptr, nullptr, nullptr, nullptr, B < chunk ? B : chunk, h, w, s);
      A.plan = false;
      need = A.peak + (1 << 20);
      if (e->d.tile_chunk > 0 || chunk == 1 || need <= ((size_t)64 << 30)) break;
      chunk = (chunk + 1) / 2;
    }
    A.peak = peak0 > need ? peak0 : need;
    if (need > A.cap) { CUDA_OK(cudaStreamSynchronize(s)); A.reserve(need); e->graphs.clear(); }
  }
  // fixed launch sequence per (mode, shapes, buffers, arena, precision): replayed as a CUDA graph from the third identical call on
  std::vector<long long> key = {mode_i, B, h, w, chunk, (long long)(uintptr_t)lr, (long long)(uintptr_t)gt, (long long)(uintptr_t)sr,
                                (prior ? prior->serial : 0), (long long)(uintptr_t)e->arena.base, g_conv_mode};
  for (int i = 0; i < nl; ++i) { key.push_back(lat_out ? (long long)(uintptr_t)lat_out[i] : 0); key.push_back(lat_in ? (long long)(uintptr_t)lat_in[i] : 0); }
  run_graphed(e->graphs, key, s, [&](cudaStream_t st) {
    for (int b0 = 0; b0 < B; b0 += chunk) {
      const int nb = B - b0 < chunk ? B - b0 : chunk;
      std::vector<float*> lo(nl, nullptr); std::vector<const float*> li(nl, nullptr);
      for (int i = 0; i < nl; ++i) {
        const int lv = e->latent_level[i];
        const size_t per = (size_t)e->latent_C[i] * ((h * S) >> lv) * ((w * S) >> lv);
        if (lat_out) lo[i] = lat_out[i] + per * b0;
        if (lat_in) li[i] = lat_in[i] + per * b0;
      }
      run_chunk(e, prior, mode, lr + (size_t)b0 * 3 * h * w, gt ? gt + (size_t)b0 * 3 * S * h * S * w : nullptr,
                lo.data(), li.data(), sr ? sr + (size_t)b0 * 3 * S * h * S * w : nullptr, nb, h, w, st);
    }
  });
  CUDA_OK(cudaGetLastError());
}

}  // namespace bfsr
