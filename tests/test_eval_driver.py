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
