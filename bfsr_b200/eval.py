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


# ================================================================================================
# LINF-LP/test.py:50-230 (eval_psnr) around the engine
class Averager:
    """LINF-LP/utils.py:15-26."""

    def __init__(self):
        self.n = 0.0
        self.v = 0.0

    def add(self, v, n=1.0):
        self.v = (self.v * self.n + v * n) / (self.n + n)
        self.n += n

    def item(self):
        return self.v


def eval_psnr(loader, model, prior_model=None, data_norm=None, eval_type=None, eval_bsize=None, window_size=0, scale_max=4,
              verbose=False, sample=0, detail=False, randomness=False, temperature=0, patch=False, save_path=None,
              device="cuda", metrics=None, lpips_fn=None):
    """`eval_psnr` of LINF-LP/test.py:50-230 for the LINF / LINF-LP models (window_size must be 0: the SwinIR reflection padding
    belongs to an encoder outside the hot path).  `model` / `prior_model` are the engines (`LINFEngine`, `LINFPriorEngine`),
    driven through the reference's operator dispatch and the row-chunked `batched_predict*`; PSNR / SSIM / LR-consistency are
    computed on the device (`bfsr_b200.metrics`; `metrics` lets a test substitute the backend), LPIPS is NaN unless `lpips_fn`
    is given.  The reference saves sample PNGs into a module-level `save_path`; here it is a parameter."""
    import torch.nn.functional as F
    from functools import partial
    from .models.linf import batched_predict, batched_predict_log_p
    Mx = metrics or M
    if window_size != 0:
        raise NotImplementedError("window_size != 0 (SwinIR padding) is outside the LINF hot path")
    if prior_model is not None:
        prior_model.eval()
    model.eval()
    if data_norm is None:
        data_norm = {"inp": {"sub": [0], "div": [1]}, "gt": {"sub": [0], "div": [1]}}
    t = data_norm["inp"]
    inp_sub = torch.FloatTensor(t["sub"]).view(1, -1, 1, 1).to(device)
    inp_div = torch.FloatTensor(t["div"]).view(1, -1, 1, 1).to(device)
    t = data_norm["gt"]          # the reference shapes these (1,1,-1); every shipped config has one value, which broadcasts
    gt_sub = torch.FloatTensor(t["sub"]).view(1, -1, 1, 1).to(device)
    gt_div = torch.FloatTensor(t["div"]).view(1, -1, 1, 1).to(device)
    scale = None
    if eval_type is None:
        psnr_fn = Mx.calc_psnr
    elif eval_type.startswith("div2k"):
        scale = int(eval_type.split("-")[1])
        psnr_fn = partial(Mx.calc_psnr, dataset="div2k", scale=scale)
    elif eval_type.startswith("benchmark"):
        scale = int(eval_type.split("-")[1])
        psnr_fn = partial(Mx.calc_psnr, dataset="benchmark", scale=scale)
    else:
        raise NotImplementedError
    if detail and scale is None:
        raise ValueError("detail=True needs an eval_type with a scale (the reference reads `scale` from it for the LR consistency)")
    val_psnr, val_lr, val_ssim, val_lpips, val_diversity = Averager(), Averager(), Averager(), Averager(), Averager()

    def to01(img):
        return torch.clamp(img * gt_div + gt_sub, 0, 1)

    def detail_metrics(img, batch):
        ssim = Mx.calculate_ssim(to01(img)[0], batch["gt"][0], mul=255.0)
        lp = float("nan") if lpips_fn is None else float(lpips_fn(torch.clamp(img, -1, 1), (batch["gt"] - gt_sub) / gt_div))
        lr_recon = Mx.imresize(to01(img)[0], 1 / scale)[None]
        return ssim, lp, psnr_fn(lr_recon, batch["inp"])

    for idx, batch in enumerate(loader):
        batch = {k: v.to(device) for k, v in batch.items()}
        inp = (batch["inp"] - inp_sub) / inp_div
        coord, cell = batch["coord"], batch["cell"]
        lr_up_residual = batch["gt_lr_up"] if prior_model is not None else None
        gh, gw = batch["gt"].shape[-2:]
        preds = []
        with torch.no_grad():
            if eval_bsize is None:
                if prior_model is not None:
                    _, z_lr = model("log_p", inp=inp, coord=coord, cell=cell, gt=lr_up_residual)
                    z_lr_learned = prior_model(z_lr, inp)
                    pred = model("rgb", inp=inp, coord=coord, cell=cell, temperature=temperature, zmap=z_lr_learned)
                else:
                    pred = model("rgb", inp=inp, coord=coord, cell=cell, temperature=temperature)
                if patch:      # central pixel of every patch only (evaluation during training, test.py:127-141)
                    ps = model.patch_size
                    pred = pred[:, :, ps // 2::ps, ps // 2::ps]
                    pred = pred + F.grid_sample(inp, coord.flip(-1), mode="bilinear", padding_mode="border", align_corners=False)
                preds = [pred]
            else:
                z_lr_learned = None
                if prior_model is not None:
                    z_lr = batched_predict_log_p(model, inp, coord, cell, lr_up_residual).detach().contiguous()
                    z_lr_learned = prior_model(z_lr, inp)
                    if z_lr_learned.shape != z_lr.shape:
                        z_lr_learned = F.interpolate(z_lr_learned, size=z_lr.shape[-2:], mode="bilinear", align_corners=False)
                for _ in range(5 if randomness else 1):
                    pred = batched_predict(model, inp, coord, cell, temperature, z_lr_learned)
                    pred = pred[..., :gh, :gw]
                    if patch:
                        pred = pred + F.interpolate(inp, pred.shape[-2:], mode="bilinear", align_corners=False)
                    preds.append(pred)
        n = inp.shape[0]
        if detail:
            res = [detail_metrics(p, batch) for p in preds]
            val_ssim.add(sum(r[0] for r in res) / len(res), n)
            val_lpips.add(sum(r[1] for r in res) / len(res), n)
            val_lr.add(sum(r[2] for r in res) / len(res), n)
        if randomness and len(preds) > 1:
            q = [torch.round(to01(p) * 255.0).unsqueeze(1) for p in preds]
            val_diversity.add(float(torch.std(torch.cat(q, 1), dim=1).mean()), n)
        p01 = [to01(p) for p in preds]
        val_psnr.add(sum(psnr_fn(p, batch["gt"]) for p in p01) / len(p01), n)
        if idx < sample and save_path is not None:
            from PIL import Image
            img = (p01[0][0].permute(1, 2, 0) * 255.0).cpu().numpy()
            os.makedirs(save_path, exist_ok=True)
            Image.fromarray(img.round().astype(np.uint8), mode="RGB").save(
                os.path.join(save_path, "{}x{}.png".format(800 + idx + 1, scale if scale is not None else 1)))
        if verbose:
            print("psnr {:.4f}".format(val_psnr.item()))
    if detail:
        result = {"psnr": val_psnr.item(), "ssim": val_ssim.item(), "lpips": val_lpips.item(), "LR recon": val_lr.item()}
        if randomness:
            result["diversity"] = val_diversity.item()
        return result
    return val_psnr.item()
