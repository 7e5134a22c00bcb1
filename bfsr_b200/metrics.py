"""Evaluation metrics of the reference's test drivers, computed on the device through the C ABI.

Mirrors LINF-LP/utils.py:132-193 (`calc_psnr(sr, hr, dataset, scale, rgb_range)`, `calculate_ssim(img1, img2)`).  The SRFlow-LP
driver takes PSNR and SSIM from skimage (SRFlow-LP/code/Measure.py:46-53): its PSNR on uint8 images is `calc_psnr(..., rgb_range=255)`;
its SSIM is skimage's default (7x7 uniform window, sample covariance) = `ssim_skimage`, a different definition from the 11x11
Gaussian `calculate_ssim`; its LR-consistency PSNR resizes a uint8 image (`imresize_u8`).  Inputs are CUDA tensors; there is no
CPU fallback.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib

_MODES = {None: 0, "benchmark": 1, "div2k": 2}


_check = _lib.check


def calc_psnr(sr, hr, dataset=None, scale=1, rgb_range=1):
    """utils.calc_psnr: -10 log10(mean(((sr - hr) / rgb_range)^2)) over the whole (B,C,H,W) tensor; dataset='benchmark' takes the
    luma of the difference and shaves `scale` border pixels, 'div2k' only shaves."""
    if dataset not in _MODES:
        raise NotImplementedError(dataset)
    assert sr.is_cuda and hr.is_cuda and sr.shape == hr.shape and sr.dim() == 4
    sr, hr = sr.contiguous().float(), hr.contiguous().float()
    B, Cc, H, W = sr.shape
    out = C.c_double()
    with torch.cuda.device(sr.device):
        _check(_lib.lib().bfsr_metric_psnr(sr.data_ptr(), hr.data_ptr(), B, Cc, H, W, _MODES[dataset], int(scale), float(rgb_range),
                                           C.byref(out), _lib.stream_ptr(sr.device)))
    return out.value


def calculate_ssim(img1, img2, mul=1.0):
    """utils.calculate_ssim on (C,H,W) CUDA tensors (the reference takes HWC numpy arrays in [0,255]: pass mul=255 for [0,1]
    inputs).  11x11 Gaussian window, sigma 1.5, valid region, fp64, mean over channels."""
    assert img1.is_cuda and img2.is_cuda and img1.shape == img2.shape
    if img1.dim() == 2:
        img1, img2 = img1[None], img2[None]
    if img1.dim() != 3:
        raise ValueError("Wrong input image dimensions.")
    img1, img2 = img1.contiguous().float(), img2.contiguous().float()
    Cc, H, W = img1.shape
    out = C.c_double()
    with torch.cuda.device(img1.device):
        _check(_lib.lib().bfsr_metric_ssim(img1.data_ptr(), img2.data_ptr(), Cc, H, W, float(mul), C.byref(out),
                                           _lib.stream_ptr(img1.device)))
    return out.value


def imresize(img, scalar_scale):
    """imresize.imresize(img, scalar_scale) (bicubic, antialiased, MATLAB-compatible) on a (C,H,W) CUDA tensor; returns float32
    (the reference's float64 result cast as test.py:185 does).  LR consistency: calc_psnr(imresize(sr, 1/s)[None], lr)."""
    assert img.is_cuda and img.dim() == 3
    img = img.contiguous().float()
    Cc, H, W = img.shape
    oh, ow = C.c_int32(), C.c_int32()
    L = _lib.lib()
    _check(L.bfsr_imresize_bicubic(None, Cc, H, W, float(scalar_scale), None, C.byref(oh), C.byref(ow), None))
    out = torch.empty((Cc, oh.value, ow.value), device=img.device, dtype=torch.float32)
    with torch.cuda.device(img.device):
        _check(L.bfsr_imresize_bicubic(img.data_ptr(), Cc, H, W, float(scalar_scale), out.data_ptr(), C.byref(oh), C.byref(ow),
                                       _lib.stream_ptr(img.device)))
    return out


def ssim_skimage(img1, img2, mul=1.0, win_size=7, sample_cov=True):
    """skimage.metrics.structural_similarity(imgA, imgB, multichannel=True) as SRFlow-LP/code/Measure.py:46-49 calls it on uint8
    images, on (C,H,W) CUDA tensors holding 0..255 values (pass mul=255 for [0,1] inputs): 7x7 uniform window, sample
    covariance, data range 255, mean over the region the window fits in, then over channels."""
    assert img1.is_cuda and img2.is_cuda and img1.shape == img2.shape
    if img1.dim() == 2:
        img1, img2 = img1[None], img2[None]
    if img1.dim() != 3:
        raise ValueError("Wrong input image dimensions.")
    img1, img2 = img1.contiguous().float(), img2.contiguous().float()
    Cc, H, W = img1.shape
    out = C.c_double()
    with torch.cuda.device(img1.device):
        _check(_lib.lib().bfsr_metric_ssim_uniform(img1.data_ptr(), img2.data_ptr(), Cc, H, W, float(mul), int(win_size),
                                                   int(bool(sample_cov)), C.byref(out), _lib.stream_ptr(img1.device)))
    return out.value


def imresize_u8(img, scalar_scale):
    """imresize.imresize on a uint8 image (SRFlow-LP/code/test.py:159): (C,H,W) CUDA tensor of 0..255 values (uint8 or float);
    every pass rounds half-to-even and clips like the reference's uint8 branch.  Returns a uint8 tensor."""
    assert img.is_cuda and img.dim() == 3
    img = img.contiguous().float()
    Cc, H, W = img.shape
    oh, ow = C.c_int32(), C.c_int32()
    L = _lib.lib()
    _check(L.bfsr_imresize_bicubic_u8(None, Cc, H, W, float(scalar_scale), None, C.byref(oh), C.byref(ow), None))
    out = torch.empty((Cc, oh.value, ow.value), device=img.device, dtype=torch.float32)
    with torch.cuda.device(img.device):
        _check(L.bfsr_imresize_bicubic_u8(img.data_ptr(), Cc, H, W, float(scalar_scale), out.data_ptr(), C.byref(oh), C.byref(ow),
                                          _lib.stream_ptr(img.device)))
    return out.to(torch.uint8)
