// Learned-prior latent module (UNet) of SRFlow-LP — SRFlow-LP/code/models/unet.py:109-181.
//
// One independent branch per latent tensor: DenseBlock_5C stem (dense NHWC buffer, conv k writes channel
// slice k), DoubleConv blocks with eval-mode BatchNorm folded into the bias-free convs, MaxPool2d(2) and the
// align_corners=True bilinear x2 as resampling kernels that write straight into the skip-concat buffers
// (torch.cat([x2, x1]) at unet.py:96 never materialises), 1x1 OutConv.
#include "engine.cuh"
#include <cmath>
#include <cstring>

using namespace bfsr;

bfsr_unet::~bfsr_unet() {
  for (auto& b : br) {
    for (auto& c : b.dense) free_conv(c);
    for (auto& c : b.lr_dense) free_conv(c);
    if (b.lr_w) cudaFree(b.lr_w);
    if (b.lr_b) cudaFree(b.lr_b);
    for (auto& c : b.inc) free_conv(c);
    for (auto& c : b.down) free_conv(c);
    for (auto& c : b.up) free_conv(c);
    free_conv(b.outc);
  }
}

namespace bfsr {

// OutConv (1x1, dim -> nf) with nf = 6 / 27: padded with zero output channels to a multiple of 8 so that it is a tcgen05 conv with
// 16-byte addressable output rows (the caller sees the first nf channels of the padded buffer); on the CUDA-core kernel the
// 64 -> 6 conv at 320x320 cost 1.15 ms per step for 0.13 ms of HBM traffic.
static ConvW pack_outc(const float* w, const float* b, int nf, int dim) {
  const int np = (nf + 7) & ~7;
  std::vector<float> wp((size_t)np * dim, 0.f), bp(np, 0.f);
  memcpy(wp.data(), w, (size_t)nf * dim * 4); memcpy(bp.data(), b, nf * 4);
  return pack_conv(wp.data(), np, dim, 1, bp.data(), nullptr, {});
}


// conv (no bias) + BatchNorm2d(eval) : W' = W*g, b' = beta - mean*g, g = gamma/sqrt(var+1e-5)   (unet.py:44-51)
static ConvW pack_conv_bn(const Weights& W, const std::string& conv, const std::string& bn, int cout, int cin) {
  const float* w = W.data(conv + ".weight", {cout, cin, 3, 3});
  const float* g = W.data(bn + ".weight", {cout});
  const float* b = W.data(bn + ".bias", {cout});
  const float* m = W.data(bn + ".running_mean", {cout});
  const float* v = W.data(bn + ".running_var", {cout});
  std::vector<float> sc(cout), bi(cout);
  for (int i = 0; i < cout; ++i) {
    const double s = (double)g[i] / std::sqrt((double)v[i] + 1e-5);
    sc[i] = (float)s; bi[i] = (float)((double)b[i] - (double)m[i] * s);
  }
  return pack_conv(w, cout, cin, 3, bi.data(), sc.data(), {});
}
static void pack_double_conv(const Weights& W, const std::string& p, int cin, int mid, int cout, ConvW* out2) {
  out2[0] = pack_conv_bn(W, p + ".double_conv.0", p + ".double_conv.1", mid, cin);
  out2[1] = pack_conv_bn(W, p + ".double_conv.3", p + ".double_conv.4", cout, mid);
}

// DenseBlock_5C with the nf input channels padded to a multiple of 8 inside the dense buffer (16-byte bf16 pixel rows for TMA)
void pack_dense5(const Weights& W, const std::string& p, int nf, int gc, int out_dim, ConvW* out5, int* nf_pad_out) {
  const int nf_pad = (nf + 7) / 8 * 8;
  *nf_pad_out = nf_pad;
  for (int c = 1; c <= 5; ++c) {
    const int cin_src = nf + (c - 1) * gc, cout = c < 5 ? gc : out_dim;
    std::vector<int> map(nf_pad + (c - 1) * gc, -1);
    for (int i = 0; i < nf; ++i) map[i] = i;
    for (int i = 0; i < (c - 1) * gc; ++i) map[nf_pad + i] = nf + i;
    const std::string q = p + ".conv" + std::to_string(c);
    out5[c - 1] = pack_conv(W.data(q + ".weight", {cout, cin_src, 3, 3}), cout, cin_src, 3, W.data(q + ".bias", {cout}),
                            nullptr, map, /*tc_min_cin=*/1);
  }
}

void unet_build(bfsr_unet* u, const bfsr_tensor_t* weights, int n) {
  Weights W(weights, n);
  const auto& d = u->d;
  BFSR_CHECK(d.bilinear == 1, "UNet: only bilinear=True is supported (the shipped priors, SURVEY.md §5)");
  const int dim = d.dim, depth = d.depth;
  if (d.variant == 1) {   // LINF-LP UNet (LINF-LP/models/unet.py:105-142)
    UNetBranchW B; B.nf = d.in_chans; B.gc = dim / 2;
    pack_dense5(W, "input_proj", B.nf, dim / 2, dim / 2, B.dense, &B.nf_pad);
    int pad2 = 0;
    pack_dense5(W, "lr_proj.2", B.nf, dim / 2, dim / 2, B.lr_dense, &pad2);
    const float* lw = W.data("lr_proj.0.weight", {B.nf, 3, 3, 3});
    const float* lb = W.data("lr_proj.0.bias", {B.nf});
    B.lr_w = to_device(std::vector<float>(lw, lw + (size_t)B.nf * 27));
    B.lr_b = to_device(std::vector<float>(lb, lb + B.nf));
    pack_double_conv(W, "inc", dim, dim, dim, B.inc);
    B.down.resize(2 * depth); B.up.resize(2 * depth);
    for (int i = 0; i < depth; ++i) {
      const int cin = dim << i, cout = (dim << (i + 1)) / (i == depth - 1 ? 2 : 1);
      pack_double_conv(W, "down_layers." + std::to_string(i) + ".maxpool_conv.1", cin, cout, cout, &B.down[2 * i]);
    }
    for (int i = 0; i < depth; ++i) {
      const int cin = dim << (depth - i), cout = (dim << (depth - i - 1)) / (i < depth - 1 ? 2 : 1);
      pack_double_conv(W, "up_layers." + std::to_string(i) + ".conv", cin, cin / 2, cout, &B.up[2 * i]);
    }
    B.outc = pack_outc(W.data("outc.conv.weight", {B.nf, dim, 1, 1}), W.data("outc.conv.bias", {B.nf}), B.nf, dim);
    u->br.push_back(B);
    return;
  }
  BFSR_CHECK(d.variant == 0, "UNet variant %d unknown", d.variant);
  BFSR_CHECK(d.n_latents >= 1 && d.n_latents <= 4, "UNet: bad latent count");
  for (int b = 0; b < d.n_latents; ++b) {
    UNetBranchW B; B.nf = d.latent_ch[b]; B.gc = dim;
    const std::string sb = std::to_string(b);
    pack_dense5(W, "input_proj" + sb, B.nf, dim, dim, B.dense, &B.nf_pad);
    pack_double_conv(W, "inc" + sb, dim, dim, dim, B.inc);
    B.down.resize(2 * depth); B.up.resize(2 * depth);
    for (int i = 0; i < depth; ++i) {
      const int cin = dim << i, cout = (dim << (i + 1)) / (i == depth - 1 ? 2 : 1);
      pack_double_conv(W, "down_layers" + sb + "." + std::to_string(i) + ".maxpool_conv.1", cin, cout, cout, &B.down[2 * i]);
    }
    for (int i = 0; i < depth; ++i) {
      const int cin = dim << (depth - i), cout = (dim << (depth - i - 1)) / (i < depth - 1 ? 2 : 1);
      pack_double_conv(W, "up_layers" + sb + "." + std::to_string(i) + ".conv", cin, cin / 2, cout, &B.up[2 * i]);
    }
    B.outc = pack_outc(W.data("outc" + sb + ".conv.weight", {B.nf, dim, 1, 1}), W.data("outc" + sb + ".conv.bias", {B.nf}), B.nf, dim);
    u->br.push_back(B);
  }
}

#define K_(...) do { if (!A.plan) { __VA_ARGS__; } } while (0)

// UNet body shared by both variants: x (dim channels) -> DoubleConv, depth x Down, depth x Up, 1x1 out
View run_unet_body(const UNetBranchW& B, int depth, int dim, Arena& A, const View& x, cudaStream_t s) {
  const int N = x.N;
  ConvEpi lrelu; lrelu.act = ACT_LRELU;
  std::vector<int> Hs(depth + 1), Ws(depth + 1), Cs(depth + 1);
  Hs[0] = x.H; Ws[0] = x.W; Cs[0] = dim;
  for (int i = 0; i < depth; ++i) { Hs[i + 1] = Hs[i] / 2; Ws[i + 1] = Ws[i] / 2; Cs[i + 1] = B.down[2 * i + 1].cout; }
  BFSR_CHECK(Hs[depth] >= 1 && Ws[depth] >= 1, "UNet: latent %dx%d too small for depth %d", x.H, x.W, depth);
  // skip-concat buffers: cat[i] = [features[i] | upsampled] at resolution i (i < depth)
  // every intermediate is only ever a conv operand: stored as bf16 (hi, lo) planes so the tcgen05 convs are TMA-fed
  const int fmt = g_conv_mode == 2 ? (int)F32 : (int)BF16X2;
  std::vector<View> cat(depth);
  for (int i = 0; i < depth; ++i) cat[i] = make_view(A, N, Hs[i], Ws[i], 2 * Cs[i], fmt);
  View t = make_view(A, N, Hs[0], Ws[0], B.inc[0].cout, fmt);
  K_(conv2d(B.inc[0], x, t, lrelu, IN_DIRECT, s));
  K_(conv2d(B.inc[1], t, cat[0].slice(0, Cs[0]), lrelu, IN_DIRECT, s));
  View cur = cat[0].slice(0, Cs[0]);
  View bottom;
  for (int i = 0; i < depth; ++i) {
    View p = make_view(A, N, Hs[i + 1], Ws[i + 1], Cs[i], fmt);
    K_(resample(cur, p, RS_MAXPOOL2, s));
    View m = make_view(A, N, Hs[i + 1], Ws[i + 1], B.down[2 * i].cout, fmt);
    K_(conv2d(B.down[2 * i], p, m, lrelu, IN_DIRECT, s));
    View o = i + 1 < depth ? cat[i + 1].slice(0, Cs[i + 1]) : make_view(A, N, Hs[i + 1], Ws[i + 1], Cs[i + 1], fmt);
    K_(conv2d(B.down[2 * i + 1], m, o, lrelu, IN_DIRECT, s));
    cur = o;
    if (i + 1 == depth) bottom = o;
  }
  View z = bottom;
  for (int i = 0; i < depth; ++i) {
    const int lvl = depth - 1 - i;
    BFSR_CHECK(z.C == Cs[lvl], "UNet: up path channel mismatch (%d vs %d)", z.C, Cs[lvl]);
    K_(resample(z, cat[lvl].slice(Cs[lvl], Cs[lvl]), RS_BILINEAR_UP2_AC, s));
    View m = make_view(A, N, Hs[lvl], Ws[lvl], B.up[2 * i].cout, fmt);
    K_(conv2d(B.up[2 * i], cat[lvl], m, lrelu, IN_DIRECT, s));
    View o = make_view(A, N, Hs[lvl], Ws[lvl], B.up[2 * i + 1].cout, fmt);
    K_(conv2d(B.up[2 * i + 1], m, o, lrelu, IN_DIRECT, s));
    z = o;
  }
  View out = make_view(A, N, Hs[0], Ws[0], B.outc.cout);      // cout = nf rounded up to a multiple of 8 (zero channels)
  K_(conv2d(B.outc, z, out, ConvEpi(), IN_DIRECT, s));
  return out.slice(0, B.nf);
}

// DenseBlock_5C (unet.py:30-36): x -> dense buffer -> conv5 output (out_dim channels)
View run_dense5(const ConvW* dense, int nf, int nf_pad, int gc, Arena& A, const View& x, cudaStream_t s) {
  ConvEpi lrelu; lrelu.act = ACT_LRELU;
  const int fmt = g_conv_mode == 2 ? (int)F32 : (int)BF16X2;
  View D = make_view(A, x.N, x.H, x.W, nf_pad + 4 * gc, fmt);
  K_(resample(x, D.slice(0, nf_pad), RS_COPY, s));
  for (int c = 0; c < 4; ++c)
    K_(conv2d(dense[c], D.slice(0, nf_pad + c * gc), D.slice(nf_pad + c * gc, gc), lrelu, IN_DIRECT, s));
  View o = make_view(A, x.N, x.H, x.W, dense[4].cout, fmt);
  K_(conv2d(dense[4], D, o, ConvEpi(), IN_DIRECT, s));
  (void)nf;
  return o;
}

// LINF-LP UNet.forward(x, lr) (LINF-LP/models/unet.py:144-167): x NHWC (B,qh,qw,in_chans), lr NCHW device (B,3,h,w)
View run_unet_linf(bfsr_unet* u, Arena& A, const View& x, const float* lr_nchw, int h, int w, cudaStream_t s) {
  BFSR_CHECK(u->d.variant == 1, "not a LINF-LP prior");
  const UNetBranchW& B = u->br[0];
  BFSR_CHECK(x.C == B.nf, "prior: latent has %d channels, expected %d", x.C, B.nf);
  const int dim = u->d.dim, half = dim / 2;
  View cat = make_view(A, x.N, x.H, x.W, dim, g_conv_mode == 2 ? (int)F32 : (int)BF16X2);
  View xa = run_dense5(B.dense, B.nf, B.nf_pad, B.gc, A, x, s);
  K_(resample(xa, cat.slice(0, half), RS_COPY, s));
  const int eh = (h + 2 - 3) / 3 + 1, ew = (w + 2 - 3) / 3 + 1;
  View e0 = make_view(A, x.N, eh, ew, B.nf);
  K_(conv3x3_s3_lrelu(lr_nchw, x.N, 3, h, w, B.lr_w, B.lr_b, e0, s));
  View e1 = run_dense5(B.lr_dense, B.nf, B.nf_pad, B.gc, A, e0, s);
  if (eh == x.H && ew == x.W) K_(resample(e1, cat.slice(half, half), RS_COPY, s));
  else K_(bilinear_resize(e1, cat.slice(half, half), s));
  return run_unet_body(B, u->d.depth, dim, A, cat, s);
}

std::vector<View> run_unet_srflow(bfsr_unet* u, Arena& A, const std::vector<View>& lat, cudaStream_t s) {
  BFSR_CHECK((int)lat.size() == u->d.n_latents, "prior: %zu latents given, %d expected", lat.size(), u->d.n_latents);
  std::vector<View> out;
  for (size_t b = 0; b < lat.size(); ++b) {
    const UNetBranchW& B = u->br[b];
    BFSR_CHECK(lat[b].C == B.nf, "prior: latent %zu has %d channels, expected %d", b, lat[b].C, B.nf);
    View o = make_view(A, lat[b].N, lat[b].H, lat[b].W, B.nf);   // survives the branch's temporaries
    const size_t mark = A.off;
    View x = run_dense5(B.dense, B.nf, B.nf_pad, B.gc, A, lat[b], s);
    View y = run_unet_body(B, u->d.depth, u->d.dim, A, x, s);
    K_(resample(y, o, RS_COPY, s));
    A.off = mark;
    out.push_back(o);
  }
  return out;
}

}  // namespace bfsr
