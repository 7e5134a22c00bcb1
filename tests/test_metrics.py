"""Evaluation metrics: the numpy oracle is pinned against outputs of the unmodified reference functions
(tests/golden/metrics.npz from oracle/make_golden_metrics.py); the device kernels are checked against both."""
import numpy as np
import pytest
import torch

from tests.util import golden


def _cases(g):
    sr, hr = g["sr"], g["hr"]
    return sr, hr, sr[0].transpose(1, 2, 0) * np.float32(255.0), hr[0].transpose(1, 2, 0) * np.float32(255.0)


def test_metrics_oracle_vs_reference_golden():
    from oracle import metrics_oracle as MO
    g = golden("metrics")
    sr, hr, a, b = _cases(g)
    assert abs(MO.calc_psnr(sr, hr) - float(g["psnr_plain"])) < 2e-4            # the reference's mean / log10 are fp32
    assert abs(MO.calc_psnr(sr, hr, "benchmark", 4) - float(g["psnr_benchmark_x4"])) < 2e-4
    assert abs(MO.calc_psnr(sr, hr, "div2k", 3) - float(g["psnr_div2k_x3"])) < 2e-4
    assert abs(MO.calc_psnr(sr * 255, hr * 255, rgb_range=255) - float(g["psnr_range255"])) < 2e-4
    assert abs(MO.calculate_ssim(a, b) - float(g["ssim_rgb"])) < 1e-10
    assert abs(MO.calculate_ssim(a[:, :, 1], b[:, :, 1]) - float(g["ssim_gray"])) < 1e-10


def test_imresize_oracle_vs_reference_golden():
    from oracle import metrics_oracle as MO
    g = golden("metrics")
    sr = g["sr"]
    for key, img, sc in (("lr_x4", sr[0], 1 / 4), ("lr_x3", sr[1], 1 / 3), ("up_x2", sr[0][:, :12, :10], 2)):
        got = MO.imresize(img.transpose(1, 2, 0), sc)
        assert got.shape == g[key].shape and np.abs(got - g[key]).max() < 1e-12, key


@pytest.mark.gpu
def test_device_imresize_vs_reference_golden():
    from bfsr_b200 import metrics as M
    g = golden("metrics")
    sr = torch.from_numpy(g["sr"]).cuda()
    for key, img, sc in (("lr_x4", sr[0], 1 / 4), ("lr_x3", sr[1], 1 / 3), ("up_x2", sr[0][:, :12, :10], 2)):
        got = M.imresize(img, sc).permute(1, 2, 0).cpu().numpy()
        assert got.shape == g[key].shape and np.abs(got - g[key].astype(np.float32)).max() < 1e-6, key


@pytest.mark.gpu
def test_device_metrics_vs_reference_golden():
    from bfsr_b200 import metrics as M
    g = golden("metrics")
    sr, hr = torch.from_numpy(g["sr"]).cuda(), torch.from_numpy(g["hr"]).cuda()
    assert abs(M.calc_psnr(sr, hr) - float(g["psnr_plain"])) < 2e-4
    assert abs(M.calc_psnr(sr, hr, "benchmark", 4) - float(g["psnr_benchmark_x4"])) < 2e-4
    assert abs(M.calc_psnr(sr, hr, "div2k", 3) - float(g["psnr_div2k_x3"])) < 2e-4
    assert abs(M.calc_psnr(sr * 255, hr * 255, rgb_range=255) - float(g["psnr_range255"])) < 2e-4
    assert abs(M.calculate_ssim(sr[0], hr[0], mul=255.0) - float(g["ssim_rgb"])) < 1e-9
    assert abs(M.calculate_ssim(sr[0, 1], hr[0, 1], mul=255.0) - float(g["ssim_gray"])) < 1e-9
    with pytest.raises(NotImplementedError):
        M.calc_psnr(sr, hr, dataset="set5")


@pytest.mark.gpu
def test_device_metrics_full_size_properties():
    """At SR-output size: PSNR of identical images is +inf, SSIM is 1, and both are symmetric / degrade with noise."""
    from bfsr_b200 import metrics as M
    from tools import synth
    x = synth.img(1, 640, 640, 5).cuda()
    y = (x + 0.05 * torch.randn_like(x)).clamp(0, 1)
    assert M.calc_psnr(x, x) == float("inf")
    assert abs(M.calculate_ssim(x[0], x[0], 255.0) - 1.0) < 1e-12
    assert abs(M.calculate_ssim(x[0], y[0], 255.0) - M.calculate_ssim(y[0], x[0], 255.0)) < 1e-12
    assert M.calculate_ssim(x[0], y[0], 255.0) < 0.99 and 20 < M.calc_psnr(x, y) < 40


# ---- SRFlow-LP's Measure.py variants: uint8 imresize (pinned on the reference's own imresize), skimage PSNR / SSIM (restated)
_U8 = (("lr8_x4", 0, 1 / 4, None), ("lr8_x3", 1, 1 / 3, None), ("lr8_x8", 0, 1 / 8, (40, 56)))


def _u8_close(got, want, key):
    # an exact .5 before rounding can fall either way with the summation order: at most 1 level, on a handful of pixels
    assert got.shape == want.shape and got.dtype == np.uint8, key
    d = np.abs(got.astype(np.int32) - want.astype(np.int32))
    assert d.max() <= 1 and int((d > 0).sum()) <= max(1, d.size // 200), (key, int(d.max()), int((d > 0).sum()))


def test_imresize_u8_oracle_vs_reference_golden():
    from oracle import metrics_oracle as MO
    g = golden("metrics")
    for key, i, sc, crop in _U8:
        img = g["sr8"][i] if crop is None else g["sr8"][i][:crop[0], :crop[1]]
        _u8_close(MO.imresize_u8(img, sc), g[key], key)


def test_skimage_restatement_closed_form_cases():
    from oracle import metrics_oracle as MO
    rng = np.random.default_rng(5)
    a = rng.integers(0, 256, size=(24, 31, 3), dtype=np.uint8)
    assert abs(MO.skimage_ssim(a, a) - 1.0) < 1e-12
    # constant images: variances vanish, S = (2 c1 c2 + C1) / (c1^2 + c2^2 + C1) everywhere
    c1, c2 = 90.0, 140.0
    x, y = np.full((16, 20), int(c1), np.uint8), np.full((16, 20), int(c2), np.uint8)
    C1 = (0.01 * 255) ** 2
    assert abs(MO.skimage_ssim(x, y) - (2 * c1 * c2 + C1) / (c1 * c1 + c2 * c2 + C1)) < 1e-12
    # one window exactly (7x7 image): plain sample statistics of the 49 pixels
    p, q = a[:7, :7, 0].astype(np.float64), a[3:10, 5:12, 1].astype(np.float64)
    C2 = (0.03 * 255) ** 2
    cov = ((p - p.mean()) * (q - q.mean())).sum() / 48
    want = ((2 * p.mean() * q.mean() + C1) * (2 * cov + C2)) / ((p.mean() ** 2 + q.mean() ** 2 + C1) * (p.var(ddof=1) + q.var(ddof=1) + C2))
    assert abs(MO.skimage_ssim(p.astype(np.uint8), q.astype(np.uint8)) - want) < 1e-10
    b = np.clip(a.astype(np.int32) + rng.integers(-9, 10, size=a.shape), 0, 255).astype(np.uint8)
    assert abs(MO.skimage_psnr(a, b) - 10 * np.log10(255.0 ** 2 / np.mean((a.astype(np.float64) - b) ** 2))) < 1e-12
    assert abs(MO.skimage_psnr(a, b) - MO.calc_psnr(a[None].astype(np.float64), b[None].astype(np.float64), rgb_range=255)) < 1e-9


@pytest.mark.gpu
def test_device_imresize_u8_vs_reference_golden():
    from bfsr_b200 import metrics as M
    g = golden("metrics")
    for key, i, sc, crop in _U8:
        img = g["sr8"][i] if crop is None else g["sr8"][i][:crop[0], :crop[1]]
        got = M.imresize_u8(torch.from_numpy(np.ascontiguousarray(img.transpose(2, 0, 1))).cuda(), sc)
        assert got.dtype == torch.uint8
        _u8_close(got.permute(1, 2, 0).cpu().numpy(), g[key], key)


@pytest.mark.gpu
def test_device_skimage_ssim_vs_oracle():
    from bfsr_b200 import metrics as M
    from oracle import metrics_oracle as MO
    g = golden("metrics")
    a8 = g["sr8"][0]
    b8 = (np.clip(g["hr"][0].transpose(1, 2, 0), 0, 1) * 255).astype(np.uint8)
    ta = torch.from_numpy(np.ascontiguousarray(a8.transpose(2, 0, 1))).cuda()
    tb = torch.from_numpy(np.ascontiguousarray(b8.transpose(2, 0, 1))).cuda()
    assert abs(M.ssim_skimage(ta, tb) - MO.skimage_ssim(a8, b8)) < 1e-9
    assert abs(M.ssim_skimage(ta[1], tb[1]) - MO.skimage_ssim(a8[:, :, 1], b8[:, :, 1])) < 1e-9
    assert abs(M.ssim_skimage(ta, ta) - 1.0) < 1e-12
    assert abs(M.ssim_skimage(ta, tb, sample_cov=False) - MO.skimage_ssim(a8, b8, sample_cov=False)) < 1e-9
    assert abs(M.calc_psnr(ta[None].float(), tb[None].float(), rgb_range=255) - MO.skimage_psnr(a8, b8)) < 1e-4   # the kernel scales the difference by fp32(1/255)
