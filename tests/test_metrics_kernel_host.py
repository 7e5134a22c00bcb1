"""The metric kernels' BODIES, compiled for the host from the very text of bfsr_b200/csrc/metrics.cu (one 'thread', CUDA
qualifiers defined away) and checked against the numpy oracle and the reference goldens.  This is a CPU-side check of the
arithmetic of `ssim_kernel` (both window definitions), `resize_dim_kernel` (float and uint8 rounding) and the host-side
`contributions` tables; the real kernels are checked on the device by the `gpu` tests of tests/test_metrics.py."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from tests.util import golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SHIM_HEAD = r"""
#include <cmath>
#include <cstdint>
#include <vector>
struct D3 { unsigned x, y, z; };
static D3 blockIdx{0, 0, 0}, threadIdx{0, 0, 0}, blockDim{1, 1, 1}, gridDim{1, 1, 1};
#define __global__
#define __device__
#define __forceinline__ inline
#define __restrict__
static double block_sum(double v) { return v; }
static void atomicAdd(double* p, double v) { *p += v; }
namespace bfsr {
"""
SHIM_TAIL = r"""
}
extern "C" double host_ssim(const float* a, const float* b, int C, int H, int W, float mul, int K, double cov_norm, const double* win) {
  double acc = 0;
  bfsr::ssim_kernel(a, b, C, H, W, mul, K, cov_norm, win, &acc);
  return acc / ((double)C * (H - K + 1) * (W - K + 1));
}
extern "C" void host_resize(const float* img, int C, int H, int W, double scale, int round_u8, float* out) {
  const int oh = (int)std::ceil(scale * H), ow = (int)std::ceil(scale * W);
  std::vector<double> w0, w1; std::vector<int> i0, i1; int P0 = 0, P1 = 0;
  bfsr::contributions(H, oh, scale, w0, i0, P0);
  bfsr::contributions(W, ow, scale, w1, i1, P1);
  std::vector<double> t0((size_t)C * oh * W), t1((size_t)C * oh * ow);
  bfsr::resize_dim_kernel<float>(img, t0.data(), C, H, W, 0, oh, P0, w0.data(), i0.data(), round_u8);
  bfsr::resize_dim_kernel<double>(t0.data(), t1.data(), C, oh, W, 1, ow, P1, w1.data(), i1.data(), round_u8);
  for (size_t i = 0; i < t1.size(); ++i) out[i] = (float)t1[i];
}
"""


@pytest.fixture(scope="module")
def host(tmp_path_factory):
    src = open(os.path.join(ROOT, "bfsr_b200", "csrc", "metrics.cu")).read()
    a = src.index("// one thread per valid pixel")
    b = src.index("}  // namespace bfsr")
    d = tmp_path_factory.mktemp("metrics_host")
    cpp, so = str(d / "m.cpp"), str(d / "m.so")
    open(cpp, "w").write(SHIM_HEAD + src[a:b] + SHIM_TAIL)
    subprocess.run(["g++", "-O1", "-ffp-contract=off", "-shared", "-fPIC", "-o", so, cpp], check=True)
    lib = C.CDLL(so)
    lib.host_ssim.restype = C.c_double
    lib.host_ssim.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, C.c_double, C.c_void_p]
    lib.host_resize.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.c_void_p]
    return lib


def _chw(img_hwc):
    return np.ascontiguousarray(img_hwc.transpose(2, 0, 1).astype(np.float32))


def test_ssim_kernel_body_both_windows(host):
    from oracle import metrics_oracle as MO
    g = golden("metrics")
    a8 = g["sr8"][0]
    b8 = (np.clip(g["hr"][0].transpose(1, 2, 0), 0, 1) * 255).astype(np.uint8)
    a, b = _chw(a8), _chw(b8)
    Cc, H, W = a.shape
    win7 = np.full(49, 1.0 / 49)
    got = host.host_ssim(a.ctypes.data, b.ctypes.data, Cc, H, W, 1.0, 7, 49.0 / 48.0, win7.ctypes.data)
    assert abs(got - MO.skimage_ssim(a8, b8)) < 1e-9
    got = host.host_ssim(a.ctypes.data, b.ctypes.data, Cc, H, W, 1.0, 7, 1.0, win7.ctypes.data)
    assert abs(got - MO.skimage_ssim(a8, b8, sample_cov=False)) < 1e-9
    # the 11x11 Gaussian definition on [0,1] images scaled by mul = 255 (pinned on the reference golden)
    k = np.exp(-((np.arange(11) - 5.0) ** 2) / (2 * 1.5 ** 2)); k /= k.sum()
    win11 = np.ascontiguousarray(np.outer(k, k).ravel())
    sr, hr = np.ascontiguousarray(g["sr"][0]), np.ascontiguousarray(g["hr"][0])
    got = host.host_ssim(sr.ctypes.data, hr.ctypes.data, 3, sr.shape[1], sr.shape[2], 255.0, 11, 1.0, win11.ctypes.data)
    assert abs(got - float(g["ssim_rgb"])) < 1e-9


def test_resize_kernel_body_float_and_uint8(host):
    g = golden("metrics")
    for key, i, sc, crop in (("lr8_x4", 0, 1 / 4, None), ("lr8_x3", 1, 1 / 3, None), ("lr8_x8", 0, 1 / 8, (40, 56))):
        img = g["sr8"][i] if crop is None else g["sr8"][i][:crop[0], :crop[1]]
        x = _chw(img)
        want = g[key]
        out = np.empty((3, want.shape[0], want.shape[1]), np.float32)
        host.host_resize(x.ctypes.data, 3, x.shape[1], x.shape[2], sc, 1, out.ctypes.data)
        assert np.array_equal(out, np.rint(out)) and out.min() >= 0 and out.max() <= 255
        d = np.abs(out.transpose(1, 2, 0).astype(np.int32) - want.astype(np.int32))
        assert d.max() <= 1 and int((d > 0).sum()) <= max(1, d.size // 200), (key, int(d.max()), int((d > 0).sum()))
    for key, img, sc in (("lr_x4", g["sr"][0], 1 / 4), ("lr_x3", g["sr"][1], 1 / 3), ("up_x2", g["sr"][0][:, :12, :10], 2)):
        x = np.ascontiguousarray(img.astype(np.float32))
        want = g[key]
        out = np.empty((3, want.shape[0], want.shape[1]), np.float32)
        host.host_resize(x.ctypes.data, 3, x.shape[1], x.shape[2], sc, 0, out.ctypes.data)
        assert np.abs(out.transpose(1, 2, 0) - want.astype(np.float32)).max() < 1e-6, key
