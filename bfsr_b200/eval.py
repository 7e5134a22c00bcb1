"""Evaluation driver of SRFlow-LP (SRFlow-LP/code/test.py:84-176) around the engine: per image pad -> LP path -> crop -> PNG,
PSNR / SSIM / LPIPS (Measure.py:31-53), LR-consistency PSNR (test.py:159-160), running CSV with the reference's resume
behaviour (`measure_full.csv_` while running, renamed to `measure_full.csv` at the end).

The metrics run on the device through the C ABI (`bfsr_b200.metrics`); there is no CPU fallback.  LPIPS needs the `lpips`
package and its AlexNet weights, neither of which is available offline: the column is written as NaN unless the caller passes
an `lpips_fn(imgA_uint8_hwc, imgB_uint8_hwc) -> float`.
"""
from __future__ import annotations

import glob
import os
import re
from collections import OrderedDict

import numpy as np
import torch

from . import metrics as M


def natsorted(seq):
    """natsort.natsorted for plain file names (test.py:38-39 sorts the PNG lists this way): digit runs compare as numbers."""
    def key(s):
        return [(0, int(t), "") if t.isdigit() else (1, 0, t) for t in re.split(r"(\d+)", s) if t != ""]
    return sorted(seq, key=key)


def fiFindByWildcard(wildcard):
    """test.py:38-39."""
    return natsorted(glob.glob(wildcard, recursive=True))


def imread(path):
    """test.py:63-64: RGB uint8 HWC."""
    import cv2
    img = cv2.imread(path)
    if img is None:
        raise FileNotFoundError(path)
    return img[:, :, [2, 1, 0]]


def imwrite(path, img):
    """test.py:66-68."""
    import cv2
    os.makedirs(os.path.dirname(path), exist_ok=True)
    cv2.imwrite(path, np.ascontiguousarray(img[:, :, [2, 1, 0]]))


def _dev(img, device):
    assert img.dtype == np.uint8 and img.ndim == 3, "uint8 HWC image expected"
    return torch.from_numpy(np.ascontiguousarray(img.transpose(2, 0, 1))).to(device)


def psnr(imgA, imgB, device="cuda"):
    """Measure.psnr / skimage peak_signal_noise_ratio on uint8 HWC images (data range 255), computed on the device."""
    a, b = _dev(imgA, device).float()[None], _dev(imgB, device).float()[None]
    return M.calc_psnr(a, b, rgb_range=255)


class Measure:
    """Measure.py:31-53 with the arithmetic on the device.  `measure(imgA, imgB)` -> [PSNR, SSIM, LPIPS]."""

    def __init__(self, net="alex", use_gpu=True, device="cuda", lpips_fn=None):
        self.device = device
        self.net = net
        self.lpips_fn = lpips_fn

    def measure(self, imgA, imgB):
        return [float(f(imgA, imgB)) for f in [self.psnr, self.ssim, self.lpips]]

    def lpips(self, imgA, imgB, model=None):
        return float("nan") if self.lpips_fn is None else self.lpips_fn(imgA, imgB)

    def ssim(self, imgA, imgB):
        return M.ssim_skimage(_dev(imgA, self.device), _dev(imgB, self.device))

    def psnr(self, imgA, imgB):
        return psnr(imgA, imgB, self.device)

    def lr_consistency_psnr(self, lq_orig, sr, scale):
        """test.py:159-160: psnr(lq_orig, imresize(sr, 1 / scale)) with the uint8 imresize."""
        lr_rec = M.imresize_u8(_dev(sr, self.device), 1.0 / scale)
        return M.calc_psnr(_dev(lq_orig, self.device).float()[None], lr_rec.float()[None], rgb_range=255)


def format_measurements(meas):
    """test.py:179-185."""
    s_out = []
    for k, v in meas.items():
        v = f"{v:0.4f}" if isinstance(v, float) else v
        s_out.append(f"{k}: {v}")
    return ", ".join(s_out)


def evaluate_srflow_dir(model, prior, lr_dir, hr_dir, test_dir, scale, conf="SRFlow-LP", measure=None, pad_factor=2,
                        write_png=True, log=print):
    """The image loop of test.py:112-176.  `model` is a `SRFlowNetEngine` (its `sr_image` is test.py:121-151: reflect pad to a
    multiple of `pad_factor`, LP path, clamp, uint8, crop); `measure` defaults to the device `Measure`.  Returns the DataFrame
    that was written to `<test_dir>/measure_full.csv`."""
    import pandas as pd
    measure = measure or Measure()
    lr_paths = fiFindByWildcard(os.path.join(lr_dir, "*.png"))
    hr_paths = fiFindByWildcard(os.path.join(hr_dir, "*.png"))
    os.makedirs(test_dir, exist_ok=True)
    fname = "measure_full.csv"
    path_out_measures = os.path.join(test_dir, fname + "_")
    path_out_measures_final = os.path.join(test_dir, fname)
    if os.path.isfile(path_out_measures_final):
        df = pd.read_csv(path_out_measures_final)
    elif os.path.isfile(path_out_measures):
        df = pd.read_csv(path_out_measures)
    else:
        df = None
    for idx_test, (lr_path, hr_path) in enumerate(zip(lr_paths, hr_paths)):
        lr = imread(lr_path)
        hr = imread(hr_path)
        sr = model.sr_image(lr, prior, pad_factor=pad_factor)
        meas = OrderedDict(conf=conf, name=idx_test)
        if write_png:
            imwrite(os.path.join(test_dir, "{:06d}.png".format(idx_test)), sr)
        meas["PSNR"], meas["SSIM"], meas["LPIPS"] = measure.measure(sr, hr)
        meas["LRC PSNR"] = float(measure.lr_consistency_psnr(lr, sr, scale))
        log(format_measurements(meas))
        df = pd.DataFrame([meas]) if df is None else pd.concat([pd.DataFrame([meas]), df])
        df.to_csv(path_out_measures + "_", index=False)
        os.rename(path_out_measures + "_", path_out_measures)
    if df is None:
        raise FileNotFoundError(f"no *.png pairs under {lr_dir} / {hr_dir}")
    df.to_csv(path_out_measures, index=False)
    os.rename(path_out_measures, path_out_measures_final)
    log(f"Results in: {path_out_measures_final}")
    log("Mean: " + format_measurements(df.mean(numeric_only=True)))
    return df
