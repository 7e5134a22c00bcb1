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
