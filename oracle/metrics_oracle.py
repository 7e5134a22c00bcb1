"""TEST INFRASTRUCTURE ONLY — numpy restatement of the reference's evaluation metrics (LINF-LP/utils.py:132-193).

Pinned against outputs of the unmodified reference functions (tests/golden/metrics.npz, oracle/make_golden_metrics.py).

`skimage_ssim` / `skimage_psnr` restate scikit-image's structural_similarity / peak_signal_noise_ratio (third-party dependency of
SRFlow-LP/code/Measure.py:26-27, `scikit-image` in SRFlow-LP/requirements.txt; the package is absent from this container and
from /root/reference): PARITY UNPINNED for those two -- they follow the published algorithm (Wang et al. 2004 as implemented by
skimage >= 0.16: uniform 7x7 filter, sample covariance, border of (win-1)/2 cropped) and are checked only against closed-form
cases (tests/test_metrics.py)."""
import numpy as np


def calc_psnr(sr, hr, dataset=None, scale=1, rgb_range=1):
    """utils.py:132-151.  sr, hr: (B,C,H,W) float arrays."""
    diff = (sr.astype(np.float64) - hr.astype(np.float64)) / rgb_range
    if dataset is not None:
        if dataset == "benchmark":
            shave = scale
            if diff.shape[1] > 1:
                conv = np.array([65.738, 129.057, 25.064], dtype=np.float32).astype(np.float64).reshape(1, 3, 1, 1) / 256
                diff = (diff * conv).sum(axis=1)
        elif dataset == "div2k":
            shave = scale
        else:
            raise NotImplementedError
        diff = diff[..., shave:-shave, shave:-shave]
    return float(-10 * np.log10((diff ** 2).mean()))


def _gauss_window():
    k = np.exp(-((np.arange(11) - 5.0) ** 2) / (2 * 1.5 ** 2))   # cv2.getGaussianKernel(11, 1.5)
    k /= k.sum()
    return np.outer(k, k)


def ssim(img1, img2):
    """utils.py:154-174 on one 2-D plane in [0,255] (valid region of an 11x11 Gaussian filter, float64)."""
    C1, C2 = (0.01 * 255) ** 2, (0.03 * 255) ** 2
    a, b = img1.astype(np.float64), img2.astype(np.float64)
    win = _gauss_window()
    H, W = a.shape

    def filt(x):
        out = np.zeros((H - 10, W - 10))
        for i in range(11):
            for j in range(11):
                out += win[i, j] * x[i:i + H - 10, j:j + W - 10]
        return out

    mu1, mu2 = filt(a), filt(b)
    s1, s2, s12 = filt(a * a) - mu1 ** 2, filt(b * b) - mu2 ** 2, filt(a * b) - mu1 * mu2
    m = ((2 * mu1 * mu2 + C1) * (2 * s12 + C2)) / ((mu1 ** 2 + mu2 ** 2 + C1) * (s1 + s2 + C2))
    return float(m.mean())


def calculate_ssim(img1, img2):
    """utils.py:177-193.  HWC (or HW) arrays in [0,255]."""
    if img1.shape != img2.shape:
        raise ValueError("Input images must have the same dimensions.")
    if img1.ndim == 2:
        return ssim(img1, img2)
    if img1.ndim == 3:
        return float(np.mean([ssim(img1[:, :, i], img2[:, :, i]) for i in range(img1.shape[2])]))
    raise ValueError("Wrong input image dimensions.")


def _cubic(x):
    a = np.abs(x); a2 = a * a; a3 = a2 * a
    return (1.5 * a3 - 2.5 * a2 + 1) * (a <= 1) + (-0.5 * a3 + 2.5 * a2 - 4 * a + 2) * ((1 < a) & (a <= 2))


def _contributions(in_len, out_len, scale):
    """imresize.py:40-64 with the bicubic kernel (width 4), antialiasing when scale < 1, mirror padding."""
    kw = 4.0 / scale if scale < 1 else 4.0
    u = np.arange(1, out_len + 1, dtype=np.float64) / scale + 0.5 * (1 - 1 / scale)
    left = np.floor(u - kw / 2)
    P = int(np.ceil(kw)) + 2
    ind = (left[:, None] + np.arange(P) - 1).astype(np.int32)
    d = u[:, None] - ind - 1
    w = scale * _cubic(scale * d) if scale < 1 else _cubic(d)
    w = w / w.sum(axis=1, keepdims=True)
    aux = np.concatenate((np.arange(in_len), np.arange(in_len - 1, -1, -1))).astype(np.int32)
    return w, aux[np.mod(ind, aux.size)]


def imresize(img, scalar_scale):
    """imresize.py:136-175 for an HWC (or HW) float image and one scalar scale: rows first, then columns, float64."""
    s = float(scalar_scale)
    H, W = img.shape[:2]
    oh, ow = int(np.ceil(s * H)), int(np.ceil(s * W))
    x = img.astype(np.float64)
    x = x[:, :, None] if x.ndim == 2 else x
    w0, i0 = _contributions(H, oh, s)
    x = np.einsum("op,opwc->owc", w0, x[i0])
    w1, i1 = _contributions(W, ow, s)
    x = np.einsum("op,hopc->hoc", w1, x[:, i1])
    return x[:, :, 0] if img.ndim == 2 else x


def imresize_u8(img, scalar_scale):
    """imresize.py:136-175 for a uint8 HWC image: as `imresize`, but each pass ends with around(clip(., 0, 255)).astype(uint8)
    (imresize.py:122-124)."""
    assert img.dtype == np.uint8
    s = float(scalar_scale)
    H, W = img.shape[:2]
    oh, ow = int(np.ceil(s * H)), int(np.ceil(s * W))
    x = img[:, :, None] if img.ndim == 2 else img
    w0, i0 = _contributions(H, oh, s)
    x = np.around(np.clip(np.einsum("op,opwc->owc", w0, x[i0].astype(np.float64)), 0, 255)).astype(np.uint8)
    w1, i1 = _contributions(W, ow, s)
    x = np.around(np.clip(np.einsum("op,hopc->hoc", w1, x[:, i1].astype(np.float64)), 0, 255)).astype(np.uint8)
    return x[:, :, 0] if img.ndim == 2 else x


def skimage_psnr(img_true, img_test, data_range=255.0):
    """skimage.metrics.peak_signal_noise_ratio on uint8 images (Measure.py:51-53): 10 log10(R^2 / mse), float64."""
    err = np.mean((img_true.astype(np.float64) - img_test.astype(np.float64)) ** 2)
    return float(10 * np.log10(data_range ** 2 / err))


def skimage_ssim(img1, img2, win_size=7, data_range=255.0, sample_cov=True):
    """skimage.metrics.structural_similarity(img1, img2, multichannel=True) on uint8 HWC (or HW) images (Measure.py:46-49)."""
    from scipy.ndimage import uniform_filter
    if img1.ndim == 3:
        return float(np.mean([skimage_ssim(img1[..., c], img2[..., c], win_size, data_range, sample_cov) for c in range(img1.shape[-1])]))
    x, y = img1.astype(np.float64), img2.astype(np.float64)
    NP = win_size ** 2
    cov_norm = NP / (NP - 1) if sample_cov else 1.0
    f = lambda t: uniform_filter(t, size=win_size)
    ux, uy = f(x), f(y)
    uxx, uyy, uxy = f(x * x), f(y * y), f(x * y)
    vx, vy, vxy = cov_norm * (uxx - ux * ux), cov_norm * (uyy - uy * uy), cov_norm * (uxy - ux * uy)
    C1, C2 = (0.01 * data_range) ** 2, (0.03 * data_range) ** 2
    S = ((2 * ux * uy + C1) * (2 * vxy + C2)) / ((ux ** 2 + uy ** 2 + C1) * (vx + vy + C2))
    pad = (win_size - 1) // 2
    return float(S[pad:S.shape[0] - pad, pad:S.shape[1] - pad].mean())
