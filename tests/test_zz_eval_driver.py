"""Host logic of the SRFlow-LP evaluation driver (bfsr_b200/eval.py = SRFlow-LP/code/test.py:84-176): file pairing, PNG output,
CSV layout and resume behaviour, with a stand-in engine and the numpy oracle as the metric backend (CPU); on the GPU the device
`Measure` is checked against the same oracle on the same images."""
import os

import numpy as np
import pytest

from bfsr_b200 import eval as E


class _NearestEngine:
    """stand-in for SRFlowNetEngine.sr_image: nearest-neighbour x`scale` (keeps the test about the driver, not the model)"""
    scale = 4

    def sr_image(self, lr_img, prior, pad_factor=2):
        return np.repeat(np.repeat(lr_img, self.scale, axis=0), self.scale, axis=1)


class _OracleMeasure:
    def measure(self, a, b):
        from oracle import metrics_oracle as MO
        return [MO.skimage_psnr(a, b), MO.skimage_ssim(a, b), float("nan")]

    def lr_consistency_psnr(self, lq, sr, scale):
        from oracle import metrics_oracle as MO
        return MO.skimage_psnr(lq, MO.imresize_u8(sr, 1.0 / scale))


def _write_pairs(tmp_path, n=3):
    rng = np.random.default_rng(11)
    names = ["10.png", "9.png", "100.png"][:n]          # natural order: 9, 10, 100
    for k, nm in enumerate(names):
        h, w = 9 + k, 12 + 2 * k
        lr = rng.integers(0, 256, size=(h, w, 3), dtype=np.uint8)
        hr = np.clip(np.repeat(np.repeat(lr, 4, 0), 4, 1).astype(np.int32) + rng.integers(-6, 7, size=(4 * h, 4 * w, 3)), 0, 255).astype(np.uint8)
        E.imwrite(str(tmp_path / "lr" / nm), lr)
        E.imwrite(str(tmp_path / "hr" / nm), hr)
    return names


def test_natural_sort_and_formatting():
    assert E.natsorted(["10.png", "9.png", "100.png", "a2.png", "a10.png"]) == ["9.png", "10.png", "100.png", "a2.png", "a10.png"]
    assert E.format_measurements({"conf": "c", "name": 3, "PSNR": 28.123456}) == "conf: c, name: 3, PSNR: 28.1235"


def test_eval_driver_csv_png_and_resume(tmp_path):
    import pandas as pd
    _write_pairs(tmp_path)
    out = tmp_path / "results"
    logs = []
    df = E.evaluate_srflow_dir(_NearestEngine(), None, str(tmp_path / "lr"), str(tmp_path / "hr"), str(out), scale=4,
                               conf="SRFlow-LP_DF2K_4X", measure=_OracleMeasure(), log=logs.append)
    assert list(df.columns) == ["conf", "name", "PSNR", "SSIM", "LPIPS", "LRC PSNR"]
    assert sorted(df["name"]) == [0, 1, 2] and list(df["name"]) == [2, 1, 0]          # newest row first (test.py:167)
    assert os.path.isfile(out / "measure_full.csv") and not os.path.exists(out / "measure_full.csv_")
    back = pd.read_csv(out / "measure_full.csv")
    assert np.allclose(back["PSNR"], df["PSNR"]) and back["LPIPS"].isna().all()
    # image 0 is the naturally first file (9.png): its SR is the nearest upsample of that LR, cropped to (4h, 4w)
    lr0 = E.imread(str(tmp_path / "lr" / "9.png"))
    sr0 = E.imread(str(out / "000000.png"))
    assert sr0.shape == (4 * lr0.shape[0], 4 * lr0.shape[1], 3) and np.array_equal(sr0[::4, ::4], lr0)
    assert (df["PSNR"] > 25).all() and (df["SSIM"] > 0.5).all() and (df["LRC PSNR"] > 20).all()
    assert logs[-1].startswith("Mean: PSNR:") or logs[-1].startswith("Mean: name:")
    # a second run finds the finished CSV and stacks its rows on top of it (test.py:104-109,167)
    df2 = E.evaluate_srflow_dir(_NearestEngine(), None, str(tmp_path / "lr"), str(tmp_path / "hr"), str(out), scale=4,
                                measure=_OracleMeasure(), write_png=False, log=lambda s: None)
    assert len(df2) == 6


def test_eval_driver_needs_images(tmp_path):
    with pytest.raises(FileNotFoundError):
        E.evaluate_srflow_dir(_NearestEngine(), None, str(tmp_path / "none"), str(tmp_path / "none"), str(tmp_path / "o"), scale=4,
                              measure=_OracleMeasure(), log=lambda s: None)


@pytest.mark.gpu
def test_device_measure_vs_oracle():
    from oracle import metrics_oracle as MO
    rng = np.random.default_rng(3)
    lr = rng.integers(0, 256, size=(17, 23, 3), dtype=np.uint8)
    sr = np.repeat(np.repeat(lr, 4, 0), 4, 1)
    hr = np.clip(sr.astype(np.int32) + rng.integers(-8, 9, size=sr.shape), 0, 255).astype(np.uint8)
    m = E.Measure()
    p, s, l = m.measure(sr, hr)
    assert abs(p - MO.skimage_psnr(sr, hr)) < 1e-4 and abs(s - MO.skimage_ssim(sr, hr)) < 1e-9 and np.isnan(l)
    want = MO.skimage_psnr(lr, MO.imresize_u8(sr, 0.25))
    got = m.lr_consistency_psnr(lr, sr, 4)
    assert abs(got - want) < 0.05, (got, want)        # an exact-.5 pixel may round the other way (1 level on a handful of pixels)


# ---- LINF-LP eval_psnr (LINF-LP/test.py:50-230): host logic with stand-in model / prior and the oracle as metric backend
class _FakeLINF:
    patch_size = 3

    def __init__(self):
        self.calls = []

    def eval(self):
        return self

    def __call__(self, op, inp=None, feat=None, coord=None, cell=None, gt=None, temperature=0, zmap=None):
        import torch
        self.calls.append(op)
        if op == "gen_feat":
            return inp
        if op in ("query_log_p", "log_p"):
            return None, gt * 2.0
        if op in ("query_rgb", "rgb"):
            B, q1, q2, _ = coord.shape
            out = torch.zeros(B, 3, 3 * q1, 3 * q2)
            if zmap is not None:      # make the prior's output visible in the prediction
                out = out + 0.01 * zmap[:, :3].repeat_interleave(3, 2).repeat_interleave(3, 3)
            if temperature:
                out = out + temperature * torch.randn(out.shape, generator=torch.Generator().manual_seed(len(self.calls)))
            return out
        raise NotImplementedError(op)


class _FakePrior:
    def eval(self):
        return self

    def __call__(self, z, inp):
        return 0.5 * z


class _OracleMetrics:
    @staticmethod
    def calc_psnr(sr, hr, dataset=None, scale=1, rgb_range=1):
        from oracle import metrics_oracle as MO
        return MO.calc_psnr(sr.numpy(), hr.numpy(), dataset, scale, rgb_range)

    @staticmethod
    def calculate_ssim(a, b, mul=1.0):
        from oracle import metrics_oracle as MO
        return MO.calculate_ssim(a.permute(1, 2, 0).numpy() * mul, b.permute(1, 2, 0).numpy() * mul)

    @staticmethod
    def imresize(img, s):
        import torch
        from oracle import metrics_oracle as MO
        return torch.from_numpy(MO.imresize(img.permute(1, 2, 0).numpy(), s).astype(np.float32)).permute(2, 0, 1)


def _linf_batches(n=2, h=12, w=16, s=4):
    import torch
    g = torch.Generator().manual_seed(21)
    out = []
    for k in range(n):
        H, W = s * h, s * w
        q1, q2 = H // 3 + 1, W // 3 + 1                      # the paired wrapper always pads (wrappers.py:218-219)
        out.append({"inp": torch.rand(1, 3, h, w, generator=g), "gt": torch.rand(1, 3, H, W, generator=g),
                    "coord": torch.rand(1, q1, q2, 2, generator=g) * 2 - 1, "cell": torch.full((1, 2), 2.0 / H),
                    "gt_lr_up": 0.1 * torch.randn(1, 27, q1, q2, generator=g)})
    return out


def test_linf_eval_psnr_host_logic():
    import torch
    import torch.nn.functional as F
    from oracle import metrics_oracle as MO
    norm = {"inp": {"sub": [0.5], "div": [0.5]}, "gt": {"sub": [0.5], "div": [0.5]}}
    batches = _linf_batches()
    model = _FakeLINF()
    res = E.eval_psnr([dict(b) for b in batches], model, _FakePrior(), data_norm=norm, eval_type="div2k-4", eval_bsize=300,
                      detail=True, patch=True, device="cpu", metrics=_OracleMetrics())
    assert set(res) == {"psnr", "ssim", "lpips", "LR recon"} and np.isnan(res["lpips"])
    assert model.calls[:3] == ["gen_feat", "query_log_p", "gen_feat"] and model.calls.count("query_rgb") == 2
    # independent recomputation: pred = 0.01 * up3(0.5 * 2 * gt_lr_up)[:3] cropped + bilinear(inp_norm), denormalised and clamped
    want_psnr, want_ssim, want_lr = [], [], []
    for b in batches:
        inp = (b["inp"] - 0.5) / 0.5
        H, W = b["gt"].shape[-2:]
        pred = (0.01 * b["gt_lr_up"][:, :3].repeat_interleave(3, 2).repeat_interleave(3, 3))[..., :H, :W]
        pred = pred + F.interpolate(inp, (H, W), mode="bilinear", align_corners=False)
        p01 = torch.clamp(pred * 0.5 + 0.5, 0, 1)
        want_psnr.append(MO.calc_psnr(p01.numpy(), b["gt"].numpy(), "div2k", 4))
        want_ssim.append(MO.calculate_ssim(p01[0].permute(1, 2, 0).numpy() * 255.0, b["gt"][0].permute(1, 2, 0).numpy() * 255.0))
        lr = MO.imresize(p01[0].permute(1, 2, 0).numpy(), 0.25).astype(np.float32).transpose(2, 0, 1)[None]
        want_lr.append(MO.calc_psnr(lr, b["inp"].numpy(), "div2k", 4))
    assert abs(res["psnr"] - np.mean(want_psnr)) < 1e-9 and abs(res["ssim"] - np.mean(want_ssim)) < 1e-9
    assert abs(res["LR recon"] - np.mean(want_lr)) < 1e-9
    # plain PSNR return, no prior, un-chunked path of patch mode (evaluation during training, test.py:118-141): the batch holds
    # one gt pixel per sampled coordinate and only the centre pixel of every 3x3 patch is scored
    tb = []
    for b in batches:
        q1, q2 = b["coord"].shape[1:3]
        tb.append(dict(b, gt=torch.rand(1, 3, q1, q2, generator=torch.Generator().manual_seed(q1))))
    v = E.eval_psnr([dict(b) for b in tb], _FakeLINF(), None, data_norm=norm, eval_type=None, eval_bsize=None, patch=True,
                    device="cpu", metrics=_OracleMetrics())
    want = []
    for b in tb:
        inp = (b["inp"] - 0.5) / 0.5
        pred = F.grid_sample(inp, b["coord"].flip(-1), mode="bilinear", padding_mode="border", align_corners=False)
        want.append(MO.calc_psnr(torch.clamp(pred * 0.5 + 0.5, 0, 1).numpy(), b["gt"].numpy()))
    assert isinstance(v, float) and abs(v - np.mean(want)) < 1e-9


def test_linf_eval_psnr_randomness_and_errors(tmp_path):
    norm = {"inp": {"sub": [0.5], "div": [0.5]}, "gt": {"sub": [0.5], "div": [0.5]}}
    batches = _linf_batches(n=1)
    res = E.eval_psnr([dict(b) for b in batches], _FakeLINF(), _FakePrior(), data_norm=norm, eval_type="benchmark-4", eval_bsize=300,
                      detail=True, randomness=True, temperature=0.2, patch=True, sample=1, save_path=str(tmp_path), device="cpu",
                      metrics=_OracleMetrics())
    assert res["diversity"] > 0 and os.path.isfile(tmp_path / "801x4.png")
    with pytest.raises(NotImplementedError):
        E.eval_psnr([], _FakeLINF(), None, window_size=8, device="cpu", metrics=_OracleMetrics())
    with pytest.raises(NotImplementedError):
        E.eval_psnr([], _FakeLINF(), None, eval_type="set5", device="cpu", metrics=_OracleMetrics())
