"""GPU parity tests of the SRFlow-LP path: CUDA engine (through the C ABI) vs the CPU oracle and the golden
fixtures recorded from the unmodified reference.  Tolerances: bit-exact for index ops; fp32 conv/coupling
<= 1e-4 relative (BASELINE.json north_star), in practice ~1e-6."""
import ctypes as C
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from tests.util import SMALL, golden, max_abs, rel_l2

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lib():
    from bfsr_b200 import _lib
    return _lib


def _small():
    from tools import synth
    from bfsr_b200 import models
    t = synth.SRFlowTopo(**SMALL)
    sd = synth.synth_srflow_state_dict(t, seed=11)
    usd = synth.synth_unet_state_dict(synth.unet_srflow_param_shapes(), seed=12)
    net = models.define_Flow(t.opt())
    net.load_state_dict(sd, strict=True)
    prior = models.make({"name": "unet", "args": {"depth": 3, "dim": 64, "bilinear": True}, "sd": usd}, load_sd=True)
    return t, sd, usd, net, prior


@pytest.mark.parametrize("shape", [(2, 3, 8, 6), (1, 12, 4, 4), (3, 96, 2, 10)])
def test_squeeze_bit_exact(lib, shape):
    from oracle import srflow_oracle as O
    x = torch.randn(*shape)
    xd = x.cuda()
    B, Cc, H, W = shape
    y = torch.empty(B, Cc * 4, H // 2, W // 2, device="cuda")
    lib.check(lib.lib().bfsr_op_squeeze2d(xd.data_ptr(), B, Cc, H, W, 0, y.data_ptr(), None))
    assert torch.equal(y.cpu(), O.squeeze2d(x))
    if Cc % 4 == 0:
        u = torch.empty(B, Cc // 4, H * 2, W * 2, device="cuda")
        lib.check(lib.lib().bfsr_op_squeeze2d(xd.data_ptr(), B, Cc, H, W, 1, u.data_ptr(), None))
        assert torch.equal(u.cpu(), O.unsqueeze2d(x))
        assert torch.equal(O.unsqueeze2d(O.squeeze2d(x)), x)


@pytest.mark.parametrize("cin,cout,ks,H,W,act", [
    (3, 64, 3, 17, 23, 0), (6, 64, 3, 33, 40, 2), (12, 64, 3, 16, 16, 1), (16, 64, 3, 5, 50, 0), (64, 32, 3, 16, 16, 1), (70, 64, 3, 9, 31, 1), (6, 12, 3, 20, 12, 0),
    (64, 64, 1, 13, 7, 2), (192, 64, 3, 32, 32, 0), (320, 128, 3, 8, 24, 2), (64, 48, 3, 40, 40, 0),
    (96, 27, 1, 5, 5, 0)])
def test_conv2d_fp32(lib, cin, cout, ks, H, W, act):
    g = torch.Generator().manual_seed(cin * 1000 + cout)
    x = torch.randn(2, cin, H, W, generator=g)
    w = torch.randn(cout, cin, ks, ks, generator=g) / (cin * ks * ks) ** 0.5
    b = torch.randn(cout, generator=g)
    ref = F.conv2d(x.double(), w.double(), b.double(), padding=ks // 2)
    ref = {0: lambda t: t, 1: lambda t: F.leaky_relu(t, 0.2), 2: F.relu}[act](ref)
    y = torch.empty(2, cout, H, W, device="cuda")
    lib.check(lib.lib().bfsr_op_conv2d(x.cuda().data_ptr(), 2, cin, H, W, w.data_ptr(), b.data_ptr(), cout, ks, act, 0,
                                      y.data_ptr(), None))
    assert rel_l2(ref, y) < 2e-6


def test_encode_decode_small_vs_golden():
    from oracle import srflow_oracle as O
    t, sd, usd, net, prior = _small()
    g = golden("srflow_small")
    lr = torch.from_numpy(g["lr"])
    lr_up = F.interpolate(lr, scale_factor=4, mode="bilinear", align_corners=False)
    epses = []
    out, nll, logdet = net(gt=lr_up, lr=lr, reverse=False, epses=epses, add_gt_noise=False)
    assert out is epses and len(epses) == 2            # appended in place, reference order
    for i, e in enumerate(epses):
        assert rel_l2(g[f"eps_lr{i}"], e) < 1e-4, i
    # decode of the reference's learned latents reproduces the reference SR
    learned = [torch.from_numpy(g[f"learned{i}"]) for i in range(2)]
    keep = list(learned)
    sr, _ = net(lr=lr, z=None, eps_std=None, reverse=True, epses=learned, reverse_with_grad=True)
    assert len(learned) == 2 and all(a is b for a, b in zip(keep, learned))   # caller's list untouched
    assert rel_l2(g["sr"], sr) < 1e-4
    # invertibility (P2): decode(encode(x)) == x
    rt, _ = net(lr=lr, reverse=True, epses=epses)
    assert max_abs(lr_up, rt) < 5e-5


def test_prior_small_vs_oracle():
    from oracle import srflow_oracle as O
    t, sd, usd, net, prior = _small()
    g = golden("srflow_small")
    eps = O.normalise_latents([torch.from_numpy(g[f"eps_lr{i}"]) for i in range(2)])
    got = prior(eps)
    for i in range(2):
        assert rel_l2(g[f"learned{i}"], got[i]) < 1e-4


@pytest.mark.parametrize("name,kw", [("srflow_small", SMALL), ("srflow_full40", {})])
def test_lp_sr_vs_golden(name, kw):
    """Whole LP path (test.py:135-148) through bfsr_srflow_lp_sr vs the reference's recorded output (P3)."""
    from tools import synth
    from bfsr_b200 import models
    g = golden(name)
    B, h, w, wseed, iseed = [int(v) for v in g["meta"]]
    t = synth.SRFlowTopo(**kw)
    net = models.define_Flow(t.opt())
    net.load_state_dict(synth.synth_srflow_state_dict(t, seed=wseed), strict=True)
    usd = synth.synth_unet_state_dict(synth.unet_srflow_param_shapes(), seed=wseed + 1)
    prior = models.make({"name": "unet", "args": {"depth": 3, "dim": 64, "bilinear": True}, "sd": usd}, load_sd=True)
    lr = torch.from_numpy(g["lr"])
    sr = net.lp_sr(lr, prior)
    assert torch.isfinite(sr).all()
    assert rel_l2(g["sr"], sr) < 1e-4
    # host-buffer entry point gives the same bits
    sr_h = net.lp_sr_host(lr.contiguous(), prior)
    assert torch.equal(sr_h, sr.cpu())


def test_lp_sr_ragged_batch_chunks():
    """Batch not a multiple of the tile chunk, non-square tiles; chunked result == per-tile result."""
    from tools import synth
    from bfsr_b200 import models
    t, sd, usd, _, prior = _small()
    net = models.define_Flow(t.opt(), tile_chunk=2)
    net.load_state_dict(sd, strict=True)
    lr = synth.img(5, 16, 24, 77)
    sr = net.lp_sr(lr, prior)
    for i in (0, 4):
        one = net.lp_sr(lr[i:i + 1], prior)
        assert torch.equal(one, sr[i:i + 1])
    empty = net.lp_sr(lr[:0], prior)
    assert empty.shape == (0, 3, 64, 96)


def test_rejects_odd_lr_size():
    from bfsr_b200 import BfsrError
    t, sd, usd, net, prior = _small()
    with pytest.raises(BfsrError):
        net.lp_sr(torch.rand(1, 3, 11, 12), prior)


@pytest.mark.parametrize("cin,cout,H,W,act", [
    (64, 32, 16, 8, 0), (64, 64, 32, 24, 1), (96, 32, 17, 23, 1), (192, 64, 40, 40, 0), (160, 32, 33, 9, 2),
    (320, 128, 24, 16, 2), (72, 64, 20, 12, 0), (64, 24, 16, 16, 3), (352, 64, 8, 8, 0), (64, 192, 10, 10, 3)])
@pytest.mark.parametrize("impl", [1, 2, 3])
def test_conv2d_tcgen05(lib, cin, cout, H, W, act, impl):
    """tcgen05 implicit-GEMM conv: split-bf16 x3 (impl 1) must be fp32-accurate; single-pass bf16 (impl 2) is the fast mode;
    impl 3 = x3 with input and output stored as bf16 (hi, lo) planes in HBM (TMA-fed A operand, split-store epilogue)."""
    g = torch.Generator().manual_seed(cin * 1000 + cout + H)
    x = torch.randn(2, cin, H, W, generator=g)
    w = torch.randn(cout, cin, 3, 3, generator=g) / (cin * 9) ** 0.5
    b = torch.randn(cout, generator=g)
    ref = F.conv2d(x.double(), w.double(), b.double(), padding=1)
    if act == 3:   # 'cross' sigmoid on odd channels
        ref = ref.clone(); ref[:, 1::2] = torch.sigmoid(ref[:, 1::2] + 2.0) + 1e-4
    else:
        ref = {0: lambda t: t, 1: lambda t: F.leaky_relu(t, 0.2), 2: F.relu}[act](ref)
    y = torch.empty(2, cout, H, W, device="cuda")
    lib.check(lib.lib().bfsr_op_conv2d(x.cuda().data_ptr(), 2, cin, H, W, w.data_ptr(), b.data_ptr(), cout, 3, act, impl,
                                      y.data_ptr(), None))
    err = rel_l2(ref, y)
    assert err < ({1: 2e-5, 2: 1e-2, 3: 2.5e-5}[impl]), err


@pytest.mark.parametrize("cin,cout,H,W", [(6, 64, 20, 24), (12, 64, 17, 9), (48, 64, 10, 10), (8, 64, 16, 8)])
def test_conv_small_cin_tcgen05(lib, cin, cout, H, W):
    """z-dependent first conv of a coupling (Cin = C/2 = 6, 12, 48): fp32 input through the register producer with a ragged
    channel group, split-bf16 x3."""
    g = torch.Generator().manual_seed(cin + cout)
    x = torch.randn(2, cin, H, W, generator=g)
    w = torch.randn(cout, cin, 3, 3, generator=g) / (cin * 9) ** 0.5
    b = torch.randn(cout, generator=g)
    ref = F.relu(F.conv2d(x.double(), w.double(), b.double(), padding=1))
    y = torch.empty(2, cout, H, W, device="cuda")
    lib.check(lib.lib().bfsr_op_conv2d(x.cuda().data_ptr(), 2, cin, H, W, w.data_ptr(), b.data_ptr(), cout, 3, 2, 1,
                                      y.data_ptr(), None))
    assert rel_l2(ref, y) < 2e-5


@pytest.mark.parametrize("impl", [1, 3])
@pytest.mark.parametrize("cin,cout,H,W,act", [(64, 64, 20, 24, 2), (256, 28, 9, 9, 0), (1024, 256, 33, 5, 2), (96, 540, 8, 8, 0)])
def test_conv1x1_tcgen05(lib, cin, cout, H, W, act, impl):
    """1x1 convs (coupling hidden layers, LINF MLP) on the tcgen05 kernel (single tap, no halo), split-bf16 x3."""
    g = torch.Generator().manual_seed(cin + cout)
    x = torch.randn(2, cin, H, W, generator=g)
    w = torch.randn(cout, cin, 1, 1, generator=g) / cin ** 0.5
    b = torch.randn(cout, generator=g)
    ref = F.conv2d(x.double(), w.double(), b.double())
    ref = {0: lambda t: t, 2: F.relu}[act](ref)
    y = torch.empty(2, cout, H, W, device="cuda")
    lib.check(lib.lib().bfsr_op_conv2d(x.cuda().data_ptr(), 2, cin, H, W, w.data_ptr(), b.data_ptr(), cout, 1, act, impl,
                                      y.data_ptr(), None))
    assert rel_l2(ref, y) < (2e-5 if impl == 1 else 2.5e-5)


@pytest.mark.parametrize("impl,tol", [(0, 2e-6), (1, 2e-5), (2, 2e-5)])
@pytest.mark.parametrize("cin,cout,H,W", [(64, 64, 12, 10), (256, 128, 17, 9), (96, 48, 8, 24)])
def test_conv_over_nearest_up2(lib, cin, cout, H, W, impl, tol):
    """conv3x3(nearest2x(x)): folded loader (fp32 / tcgen05) and the four-phase 2x2 evaluation all equal the reference op."""
    g = torch.Generator().manual_seed(cin + cout + H)
    x = torch.randn(2, cin, H, W, generator=g)
    w = torch.randn(cout, cin, 3, 3, generator=g) / (cin * 9) ** 0.5
    b = torch.randn(cout, generator=g)
    ref = F.conv2d(F.interpolate(x.double(), scale_factor=2, mode="nearest"), w.double(), b.double(), padding=1)
    y = torch.empty(2, cout, 2 * H, 2 * W, device="cuda")
    lib.check(lib.lib().bfsr_op_conv2d_up2(x.cuda().data_ptr(), 2, cin, H, W, w.data_ptr(), b.data_ptr(), cout, impl,
                                          y.data_ptr(), None))
    assert rel_l2(ref, y) < tol


@pytest.mark.parametrize("chi,clo,cout,H,W,act", [(64, 256, 128, 12, 10, 2), (32, 64, 64, 17, 9, 0), (64, 256, 1024, 8, 16, 2),
                                                 (64, 32, 24, 20, 24, 0)])
def test_conv_hi_lo_single_pass_phase(lib, chi, clo, cout, H, W, act):
    """conv3x3(cat[x_hi, nearest2x(x_lo)]) evaluated per output phase in one pass (TMA parity planes + pre-summed taps)
    equals the plain op on the materialised tensor (SRFlowNet_arch.py:136 conditioning of the finest level)."""
    g = torch.Generator().manual_seed(chi + clo + cout + H)
    xh = torch.randn(2, chi, 2 * H, 2 * W, generator=g)
    xl = torch.randn(2, clo, H, W, generator=g)
    w = torch.randn(cout, chi + clo, 3, 3, generator=g) / ((chi + clo) * 9) ** 0.5
    b = torch.randn(cout, generator=g)
    xin = torch.cat([xh, F.interpolate(xl, scale_factor=2, mode="nearest")], 1)
    ref = F.conv2d(xin.double(), w.double(), b.double(), padding=1)
    if act == 2:
        ref = F.relu(ref)
    y = torch.empty(2, cout, 2 * H, 2 * W, device="cuda")
    xh_d, xl_d = xh.cuda(), xl.cuda()    # keep both device tensors alive: two temporaries would share one cached block
    lib.check(lib.lib().bfsr_op_conv2d_hi_lo(xh_d.data_ptr(), xl_d.data_ptr(), 2, chi, clo, H, W, w.data_ptr(),
                                            b.data_ptr(), cout, act, y.data_ptr(), None))
    assert rel_l2(ref, y) < 3e-5


@pytest.mark.parametrize("cout,H,W,act", [(12, 20, 24, 3), (24, 16, 16, 3), (24, 33, 17, 0), (12, 7, 40, 0), (8, 16, 16, 0), (16, 30, 30, 3)])
def test_conv_tap_folded_small_cout(lib, cout, H, W, act):
    """3x3 convs with 64 input and <= 24 output channels (the (shift, scale) heads of every coupling): one GEMM over the halo tile
    with N = 9*Cout columns plus a shift-add epilogue must equal the plain conv (split-bf16 x3)."""
    g = torch.Generator().manual_seed(cout * 100 + H)
    x = torch.randn(2, 64, H, W, generator=g)
    w = torch.randn(cout, 64, 3, 3, generator=g) / (64 * 9) ** 0.5
    b = torch.randn(cout, generator=g)
    ref = F.conv2d(x.double(), w.double(), b.double(), padding=1)
    if act == 3:
        ref = ref.clone(); ref[:, 1::2] = torch.sigmoid(ref[:, 1::2] + 2.0) + 1e-4
    y = torch.empty(2, cout, H, W, device="cuda")
    x_d = x.cuda()
    lib.check(lib.lib().bfsr_op_conv2d(x_d.data_ptr(), 2, 64, H, W, w.data_ptr(), b.data_ptr(), cout, 3, act, 4, y.data_ptr(), None))
    assert rel_l2(ref, y) < 2e-5


@pytest.mark.parametrize("impl", [5, 6, 7])
@pytest.mark.parametrize("cin,cout,H,W,act", [(64, 32, 16, 30, 1), (96, 32, 17, 23, 1), (160, 32, 33, 61, 2), (64, 12, 20, 24, 3),
                                             (64, 24, 16, 16, 3), (64, 24, 9, 95, 0), (128, 32, 40, 40, 1), (64, 8, 5, 31, 0),
                                             (80, 16, 12, 64, 0)])
def test_conv_dx_folded(lib, cin, cout, H, W, act, impl):
    """3x3 convs with <= 32 output channels evaluated with the three dx taps folded into N (one MMA per filter row over a
    raster of pitch 32, shift-add by warp shuffles): fp32 output (5), BF16X2 output through the 30-pixel TMA store box (6),
    bf16 single-pass mode (7).  Sizes cover ragged tiles in both directions, Cin % 32 != 0 and a block stride of 16 / 32."""
    g = torch.Generator().manual_seed(cin * 1000 + cout + H)
    x = torch.randn(2, cin, H, W, generator=g)
    w = torch.randn(cout, cin, 3, 3, generator=g) / (cin * 9) ** 0.5
    b = torch.randn(cout, generator=g)
    ref = F.conv2d(x.double(), w.double(), b.double(), padding=1)
    if act == 3:
        ref = ref.clone(); ref[:, 1::2] = torch.sigmoid(ref[:, 1::2] + 2.0) + 1e-4
    else:
        ref = {0: lambda t: t, 1: lambda t: F.leaky_relu(t, 0.2), 2: F.relu}[act](ref)
    y = torch.empty(2, cout, H, W, device="cuda")
    x_d = x.cuda()
    lib.check(lib.lib().bfsr_op_conv2d(x_d.data_ptr(), 2, cin, H, W, w.data_ptr(), b.data_ptr(), cout, 3, act, impl, y.data_ptr(), None))
    assert rel_l2(ref, y) < {5: 2e-5, 6: 2.5e-5, 7: 1e-2}[impl]


def test_full_size_roundtrip_and_chunk_independence():
    """BASELINE config-2 geometry (shipped topology nb=23, K=16, L=3; 160x160 LR tiles): size-independent properties at full
    size -- decode(encode(x)) == x (P2), results independent of how the batch is chunked (bit-equal), finite LP output."""
    from tools import synth
    from bfsr_b200 import models
    t = synth.SRFlowTopo()
    sd = synth.synth_srflow_state_dict(t, seed=0)
    usd = synth.synth_unet_state_dict(synth.unet_srflow_param_shapes(), seed=1)
    prior = models.make({"name": "unet", "args": {"depth": 3, "dim": 64, "bilinear": True}, "sd": usd}, load_sd=True)
    lr = synth.img(6, 160, 160, 1236)
    net = models.define_Flow(t.opt(), tile_chunk=4)
    net.load_state_dict(sd, strict=True)
    lr_up = F.interpolate(lr, scale_factor=4, mode="bilinear", align_corners=False)
    epses = []
    net(gt=lr_up, lr=lr, reverse=False, epses=epses, add_gt_noise=False)
    assert [tuple(e.shape) for e in epses] == [(6, 6, 320, 320), (6, 96, 80, 80)]
    rt, _ = net(lr=lr, reverse=True, epses=epses)
    # 54 steps deep the fp32 conditioning of the synthetic flow itself limits the round trip: the all-fp32 CUDA-core mode
    # (precision=2) measures rel-L2 2.0e-5 / max-abs 2.0e-4 on these tiles, the tensor-core mode 2.2e-5 / 2.3e-4
    assert rel_l2(lr_up, rt) < 1e-4 and max_abs(lr_up, rt) < 1e-3
    sr = net.lp_sr(lr, prior)                       # chunks of 4 + 2 tiles
    assert torch.isfinite(sr).all() and tuple(sr.shape) == (6, 3, 640, 640)
    net1 = models.define_Flow(t.opt(), tile_chunk=3)
    net1.load_state_dict(sd, strict=True)
    assert torch.equal(net1.lp_sr(lr, prior), sr)   # chunks of 3 + 3 tiles give the same bits


def test_sr_image_driver_odd_size():
    """Image-level body of test.py:121-151 (reflect pad to even, t(), LP path, clamp, truncating uint8 cast, crop) on an odd-sized
    uint8 image vs the same steps around the oracle."""
    import numpy as np
    from oracle import srflow_oracle as O
    t, sd, usd, net, prior = _small()
    rs = np.random.RandomState(3)
    base = rs.randint(0, 256, size=(4, 6, 3)).astype(np.float32)
    lr = np.clip(np.kron(base, np.ones((4, 4, 1), np.float32))[:15, :21] + rs.randn(15, 21, 3) * 4, 0, 255).astype(np.uint8)
    got = net.sr_image(lr, prior)
    assert got.shape == (60, 84, 3) and got.dtype == np.uint8
    pad = np.pad(lr, [(0, 1), (0, 1), (0, 0)], "reflect")
    lr_t = torch.from_numpy(pad.transpose(2, 0, 1)[None].astype(np.float32)) / 255
    ref = O.lp_sr(sd, usd, t, lr_t, literal=False)
    ref = (np.clip(ref[0].numpy().transpose(1, 2, 0), 0, 1) * 255).astype(np.uint8)[:60, :84]
    diff = np.abs(got.astype(np.int32) - ref.astype(np.int32))
    assert diff.max() <= 1 and (diff != 0).mean() < 0.01     # 1e-5-level differences can flip a truncation boundary


def test_x8_topology_encode_decode_vs_reference_golden():
    """8x SRFlow (BASELINE config 4 family: L = 4, latents at 4h / 2h / h/2, two Split2d, conditioning by fea_up4 / fea_up2 /
    fea_up1 / fea_up0) against outputs recorded from the unmodified reference."""
    from tools import synth
    from bfsr_b200 import models
    g = golden("srflow_x8_small")
    B, h, w, wseed, iseed = [int(v) for v in g["meta"]]
    t = synth.SRFlowTopo(scale=8, L=4, nb=2, blocks=(0, 1, 0, 1), K=1)
    net = models.define_Flow(t.opt())
    net.load_state_dict(synth.synth_srflow_state_dict(t, seed=wseed), strict=True)
    lr = torch.from_numpy(g["lr"])
    lr_up = F.interpolate(lr, scale_factor=8, mode="bilinear", align_corners=False)
    epses = []
    net(gt=lr_up, lr=lr, reverse=False, epses=epses, add_gt_noise=False)
    assert [tuple(e.shape) for e in epses] == [(1, 6, 64, 48), (1, 12, 32, 24), (1, 192, 8, 6)]
    for i, e in enumerate(epses):
        assert rel_l2(g[f"eps{i}"], e) < 1e-4, i
    sr, _ = net(lr=lr, reverse=True, epses=[0.5 * torch.from_numpy(g[f"eps{i}"]) for i in range(3)])
    assert rel_l2(g["sr_half"], sr) < 1e-4
    rt, _ = net(lr=lr, reverse=True, epses=epses)
    assert max_abs(lr_up, rt) < 5e-5


@pytest.mark.skipif(os.environ.get("BFSR_TEST_CLUSTER") != "1",
                    reason="2-CTA cluster / multicast weight variant is opt-in until measured (BFSR_TEST_CLUSTER=1)")
def test_cluster_multicast_variant_subprocess():
    """conv_tc_kernel<true, 2> (BFSR_TC_CLUSTER=2: every eligible TMA-fed conv runs as 2-CTA clusters with multicast weight
    stages) must pass the same parity tests as the default kernel.  The knob is read once per process, hence the subprocess."""
    import subprocess
    import sys
    env = dict(os.environ, BFSR_TC_CLUSTER="2", BFSR_TEST_CLUSTER="0")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(root, "tests", "test_gpu_srflow.py"), "-q", "-x", "-m", "gpu",
                        "-k", "not cluster_multicast"], env=env, cwd=root, capture_output=True, text=True, timeout=1500)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]


def test_one_launch_coupling_vs_three_launch_chain(tmp_path):
    """coupling_fused_kernel (C = 12 levels: z-conv -> 1x1 -> tap-folded head -> FlowStep in one launch, hidden maps in tensor memory)
    against the three conv_tc launches it replaces, through bfsr_op_flowstep in both directions: ragged strips / rows, batch > 1,
    several vertical segments per strip (forced with BFSR_CF_SEGS as well).  Same split-bf16 products, different summation order:
    <= 5e-6 rel-L2.  The switches are read once per process, hence the subprocesses (tools/cpl_check.py)."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    from tools import cpl_check
    outs = {}
    for name, env in (("chain", {"BFSR_FUSE_CPL": "0"}), ("fused", {"BFSR_FUSE_CPL": "1"}), ("fused_s3", {"BFSR_FUSE_CPL": "1", "BFSR_CF_SEGS": "3"})):
        path = str(tmp_path / f"{name}.pt")
        r = subprocess.run([sys.executable, os.path.join(root, "tools", "cpl_check.py"), "run", path], env=dict(os.environ, **env), cwd=root,
                           capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        outs[name] = path
    assert cpl_check.cmp(outs["chain"], outs["fused"], tol=5e-6) == 0
    assert cpl_check.cmp(outs["chain"], outs["fused_s3"], tol=5e-6) == 0
    # bf16 single-pass mode (precision 1): same hi x hi products in both forms; a 1e-7 difference of summation order flips the bf16
    # rounding of a hidden value now and then (2^-9 relative at that point), hence the wider gate (the mode itself is a 5e-3 class)
    for name, env in (("chain_fast", {"BFSR_FUSE_CPL": "0", "BFSR_CPL_CHECK_PREC": "1"}), ("fused_fast", {"BFSR_FUSE_CPL": "1", "BFSR_CPL_CHECK_PREC": "1"})):
        path = str(tmp_path / f"{name}.pt")
        r = subprocess.run([sys.executable, os.path.join(root, "tools", "cpl_check.py"), "run", path], env=dict(os.environ, **env), cwd=root,
                           capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        outs[name] = path
    assert cpl_check.cmp(outs["chain_fast"], outs["fused_fast"], tol=1e-4) == 0


# ------------------------------------------------------------------ reference-recorded pins at the BASELINE shapes (round 2)
def test_config2_tile_vs_reference_golden():
    """BASELINE config 2 geometry: tile 0 of the bench batch (160x160 LR, shipped topology) through the engine -- fused LP path,
    and encode / prior / decode step by step -- against the values recorded from the UNMODIFIED reference on the same tile."""
    from tests.test_oracle_golden import check_full160, full160_inputs
    from bfsr_b200 import models
    g, t, sd, usd, lr = full160_inputs()
    net = models.define_Flow(t.opt())
    net.load_state_dict(sd, strict=True)
    prior = models.make({"name": "unet", "args": {"depth": 3, "dim": 64, "bilinear": True}, "sd": usd}, load_sd=True)
    sr = net.lp_sr(lr, prior)
    assert torch.isfinite(sr).all()
    check_full160(g, sr.cpu(), tol=1e-4)
    lr_up = F.interpolate(lr, scale_factor=4, mode="bilinear", align_corners=False)
    epses = []
    net(gt=lr_up, lr=lr, reverse=False, epses=epses, add_gt_noise=False)
    check_full160(g, sr.cpu(), [e.cpu() for e in epses], tol=1e-4)


def test_config4_topology_vs_reference_golden():
    """BASELINE config 4 topology (8x, nb=23, K=16, L=4, two Split2d) on a 40x40 tile: encode, decode(0.9 x latents) and the round
    trip against the reference's recorded outputs."""
    from tests.test_oracle_golden import X8K16
    from tools import synth
    from bfsr_b200 import models
    g = golden("srflow_x8_k16")
    B, h, w, wseed, iseed = [int(v) for v in g["meta"]]
    t = synth.SRFlowTopo(**X8K16)
    net = models.define_Flow(t.opt())
    net.load_state_dict(synth.synth_srflow_state_dict(t, seed=wseed), strict=True)
    lr = torch.from_numpy(g["lr"])
    lr_up = F.interpolate(lr, scale_factor=8, mode="bilinear", align_corners=False)
    epses = []
    net(gt=lr_up, lr=lr, reverse=False, epses=epses, add_gt_noise=False)
    assert [list(e.shape) for e in epses] == g["shapes"].tolist()
    for i, e in enumerate(epses):
        assert rel_l2(g[f"eps{i}"], e) < 1e-4, i
    # decode of 0.9 x latents: off the data manifold the untrained 64-step inverse amplifies rounding (the reference's own round
    # trip on this tile is only good to 2.2e-4); the all-fp32 mode must still meet 1e-4, the split-bf16 x3 mode 5e-4
    lat = [0.9 * torch.from_numpy(g[f"eps{i}"]) for i in range(3)]
    sr, _ = net(lr=lr, reverse=True, epses=lat)
    assert rel_l2(g["sr_s2"], sr[..., ::2, ::2]) < 5e-4 and rel_l2(g["sr_tl"], sr[..., :48, :48]) < 5e-4
    net32 = models.define_Flow(t.opt(), precision=2)
    net32.load_state_dict(synth.synth_srflow_state_dict(t, seed=wseed), strict=True)
    sr32, _ = net32(lr=lr, reverse=True, epses=lat)
    assert rel_l2(g["sr_s2"], sr32[..., ::2, ::2]) < 1e-4 and rel_l2(g["sr_tl"], sr32[..., :48, :48]) < 1e-4
    rt, _ = net(lr=lr, reverse=True, epses=epses)
    assert max_abs(lr_up, rt) < 1e-3          # the reference's own round trip on this tile: 2.2e-4


def test_config4_lp_path_vs_reference_blocks():
    """BASELINE config 4 LP path (8x, L=4, three latents 6 / 12 / 192 channels) through bfsr_srflow_lp_sr with the three-branch prior
    (`make({'name': 'unet', 'args': {..., 'latent_ch': (6, 12, 192)}})`) against the reference's SRFlowNet + a prior assembled from
    the reference's own UNet blocks (tests/golden/srflow_x8_lp.npz)."""
    from tests.test_oracle_golden import x8lp_inputs
    from bfsr_b200 import models
    g, t, sd, usd, lr = x8lp_inputs()
    net = models.define_Flow(t.opt())
    net.load_state_dict(sd, strict=True)
    prior = models.make({"name": "unet", "args": {"depth": 3, "dim": 64, "bilinear": True, "latent_ch": (6, 12, 192)}, "sd": usd}, load_sd=True)
    sr = net.lp_sr(lr, prior)
    assert torch.isfinite(sr).all()
    assert rel_l2(g["sr_s2"], sr[..., ::2, ::2]) < 1e-4 and rel_l2(g["sr_tl"], sr[..., :48, :48]) < 1e-4


def _module_table(lib):
    from tests.test_oracle_golden import module_cases
    g, sd, HW, inputs = module_cases()
    table, keep = lib.tensor_table(sd)
    return g, HW, inputs, table, keep


@pytest.mark.parametrize("precision,tol", [(0, 1e-4), (2, 1e-5)])
def test_flowstep_modules_vs_reference(lib, precision, tol):
    """P1: every FlowStep of the small topology (no-coupling and coupling, C = 12 / 24 / 96), forward and reverse,
    through bfsr_op_flowstep -- the kernels the engine runs for that step, including the fused conv epilogue at C = 12 / 24 --
    against the outputs of the reference's FlowStep modules on the same z / ft (tests/golden/srflow_modules.npz).
    precision 2 (all convs on the fp32 CUDA cores) isolates the flow arithmetic: <= 1e-5."""
    g, HW, inputs, table, keep = _module_table(lib)
    n = 0
    for idx, kind, C in g["layers"].tolist():
        if kind == 2:
            continue
        H, W = HW[C]
        z, ft = inputs(idx, C, H, W)
        zd, fd = z.cuda(), ft.cuda()
        for rev, key in ((0, "fwd"), (1, "inv")):
            out = torch.empty_like(zd)
            lib.check(lib.lib().bfsr_op_flowstep(table, len(table), f"flowUpsamplerNet.layers.{idx}".encode(), C, kind, rev,
                                                zd.data_ptr(), fd.data_ptr(), 2, H, W, out.data_ptr(), precision, 1, None))
            assert rel_l2(g[f"l{idx}_{key}"], out) < tol, (idx, kind, C, key)
            n += 1
    assert n >= 24


def test_split2d_module_vs_reference(lib):
    """P1: Split2d forward (z -> z1, eps) and reverse (z1, eps -> z) vs the reference module (Split.py:49-77)."""
    g, HW, inputs, table, keep = _module_table(lib)
    n = 0
    for idx, kind, C in g["layers"].tolist():
        if kind != 2:
            continue
        z, eps = inputs(idx, C, 12, 10, split=True)
        zd, ed = z.cuda(), eps.cuda()
        z1 = torch.empty(2, C // 2, 12, 10, device="cuda"); e = torch.empty_like(z1)
        p = f"flowUpsamplerNet.layers.{idx}".encode()
        lib.check(lib.lib().bfsr_op_split2d(table, len(table), p, C, 0, zd.data_ptr(), None, 2, 12, 10, z1.data_ptr(), e.data_ptr(), None))
        assert torch.equal(z1.cpu(), z[:, :C // 2]) and rel_l2(g[f"l{idx}_eps"], e) < 1e-5
        zi_in = z[:, :C // 2].contiguous().cuda()
        zo = torch.empty_like(zd)
        lib.check(lib.lib().bfsr_op_split2d(table, len(table), p, C, 1, zi_in.data_ptr(), ed.data_ptr(), 2, 12, 10, zo.data_ptr(), None, None))
        assert rel_l2(g[f"l{idx}_zinv"], zo) < 1e-5
        n += 1
    assert n == 1


# ------------------------------------------------------------------ a15 / f1: the SRFlowModel surface and the sampling mode
def test_srflow_model_surface_runs_the_reference_test_loop():
    """SRFlow-LP/code/test.py:135-148 written against `SRFlowModel` (get_encode_z with an in-place `epses` list, the latent
    normalisation, `prior_model(epses)`, get_sr(lq, epses=...)) on top of the engine reproduces the reference's recorded SR; the
    wrapper leaves the net in train() mode like the reference (SRFlow_model.py:205,221) without disturbing the packed weights."""
    from bfsr_b200 import models
    t, sd, usd, net, prior = _small()
    g = golden("srflow_small")
    model = models.SRFlowModel(t.opt(), netG=net)
    lr_t = torch.from_numpy(g["lr"])
    lr_up = F.interpolate(lr_t, scale_factor=4, mode="bilinear", align_corners=False)
    epses_lr = []
    model.get_encode_z(lr_t, lr_up, epses=epses_lr, add_gt_noise=False)
    assert len(epses_lr) == 2 and net.training
    epses = [e.detach() for e in epses_lr]
    for i in range(len(epses)):
        mean = torch.mean(epses[i], dim=[1], keepdim=True)
        std = torch.std(epses[i], dim=[1], keepdim=True)
        epses[i] = (epses[i] - mean) / (std + 1e-8)
    learned = prior(epses)
    sr_t = model.get_sr(lq=lr_t, epses=learned)
    assert rel_l2(g["sr"], sr_t) < 1e-4 and net.training
    z = model.get_z(0.5, seed=3, batch_size=2, lr_shape=lr_t.shape)
    assert tuple(z.shape) == (2, 96, lr_t.shape[2] // 2, lr_t.shape[3] // 2)


def test_sampling_mode_without_prior():
    """f1: `get_sr(lq, heat=tau)` (SRFlow_model.py:198-237) -- z ~ N(0, tau^2) from get_z, Split2d draws its eps ~ N(0, tau^2)
    (Split.py:66-68).  heat = 0 equals the oracle's decode of all-zero latents; heat > 0 is reproducible under the seed and equals
    the decode of the explicitly drawn (eps, z)."""
    from oracle import srflow_oracle as O
    from bfsr_b200 import models
    t, sd, usd, net, prior = _small()
    model = models.SRFlowModel(t.opt(), netG=net)
    lr = torch.from_numpy(golden("srflow_small")["lr"])
    B, _, h, w = lr.shape
    sr0 = model.get_sr(lr, heat=0)
    ref0 = O.decode(sd, t, lr, [torch.zeros(B, 6, 2 * h, 2 * w), torch.zeros(B, 96, h // 2, w // 2)])
    assert rel_l2(ref0, sr0) < 1e-4
    tau = 0.3
    sr_a = model.get_sr(lr, heat=tau, seed=5)
    sr_b = model.get_sr(lr, heat=tau, seed=5)
    assert torch.isfinite(sr_a).all() and torch.equal(sr_a, sr_b)
    torch.manual_seed(5)                                           # replay the two draws: z on the host (get_z), eps on the device
    z = torch.normal(mean=0, std=tau, size=(B, 96, h // 2, w // 2))
    eps = torch.zeros((B, 6, 2 * h, 2 * w), device="cuda").normal_(0.0, tau)
    sr_c, _ = net(lr=lr, reverse=True, epses=[eps, z])
    assert torch.equal(sr_a, sr_c)
    assert not torch.equal(sr_a, model.get_sr(lr, heat=tau, seed=6))


def test_div2k_size_image_in_one_pass():
    """f3: a full DIV2K-validation-size LR image (339 x 510, odd height -> reflect-padded like test.py:126-130) through
    `sr_image` in ONE pass: the reference never tiles an image (tiling would change the result at the seams -- the RRDB trunk alone
    has a receptive field of ~350 LR pixels), and on a 180 GB part the whole-image workspace (about 10 GB) simply fits.  Checked:
    output geometry, finiteness, the workspace bound, and invertibility at that size."""
    import ctypes as C
    from tools import synth
    from bfsr_b200 import _lib, models
    t = synth.SRFlowTopo()
    net = models.define_Flow(t.opt())
    net.load_state_dict(synth.synth_srflow_state_dict(t, seed=0), strict=True)
    usd = synth.synth_unet_state_dict(synth.unet_srflow_param_shapes(), seed=1)
    prior = models.make({"name": "unet", "args": {"depth": 3, "dim": 64, "bilinear": True}, "sd": usd}, load_sd=True)
    lr = (synth.img(1, 339, 510, 99)[0].permute(1, 2, 0) * 255).round().to(torch.uint8).numpy()
    sr = net.sr_image(lr, prior)
    assert sr.shape == (339 * 4, 510 * 4, 3) and sr.dtype == np.uint8
    assert _lib.lib().bfsr_srflow_workspace_bytes(net.handle()) < 20e9
    lr_t = synth.img(1, 340, 510, 98)
    lr_up = F.interpolate(lr_t, scale_factor=4, mode="bilinear", align_corners=False)
    epses = []
    net(gt=lr_up, lr=lr_t, reverse=False, epses=epses, add_gt_noise=False)
    rt, _ = net(lr=lr_t, reverse=True, epses=epses)
    assert rel_l2(lr_up, rt) < 1e-4 and max_abs(lr_up, rt) < 5e-3


def test_cuda_graph_replay_matches_plain_launches():
    """The engine replays its fixed launch plan as a CUDA graph from the third identical call on (same shapes and buffers).  New
    CONTENTS in the same buffers must give the same bits as plain launches, and a refreshed prior must never be served from a plan
    captured with the old prior's weights."""
    from tools import synth
    from bfsr_b200 import models
    t, sd, usd, net, prior = _small()
    lr = synth.img(3, 16, 16, 501).cuda()
    out = torch.empty(3, 3, 64, 64, device="cuda")
    for _ in range(3):
        net.lp_sr(lr, prior, out=out)                 # third call replays the captured plan
    for seed in (502, 503):
        lr.copy_(synth.img(3, 16, 16, seed))
        net.lp_sr(lr, prior, out=out)                 # replay on new contents
        fresh = models.define_Flow(t.opt())
        fresh.load_state_dict(sd, strict=True)
        assert torch.equal(out, fresh.lp_sr(lr.clone(), prior))    # first call of a new engine: plain launches
    usd2 = synth.synth_unet_state_dict(synth.unet_srflow_param_shapes(), seed=99)
    prior.load_state_dict(usd2)                        # new packed weights (possibly at the old address)
    got = net.lp_sr(lr, prior, out=out).clone()
    prior2 = models.make({"name": "unet", "args": {"depth": 3, "dim": 64, "bilinear": True}, "sd": usd2}, load_sd=True)
    fresh = models.define_Flow(t.opt())
    fresh.load_state_dict(sd, strict=True)
    assert torch.equal(got, fresh.lp_sr(lr.clone(), prior2))
