"""Golden values of the reference's metric functions (LINF-LP/utils.py calc_psnr / calculate_ssim, unmodified, imported from
/root/reference with the offline stubs) on seeded image-like tensors -> tests/golden/metrics.npz.  Build container only:
    python -m oracle.make_golden_metrics"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def main():
    sys.path.insert(0, ROOT)
    from tools import synth
    sys.path.insert(0, os.path.join(HERE, "stubs"))
    sys.path.insert(0, "/root/reference/LINF-LP")
    import utils as ref_utils   # LINF-LP/utils.py
    out = {}
    hr = synth.img(2, 44, 60, 901)
    sr = (hr + 0.03 * torch.randn(hr.shape, generator=torch.Generator().manual_seed(902))).clamp(0, 1)
    out["hr"], out["sr"] = hr.numpy(), sr.numpy()
    out["psnr_plain"] = float(ref_utils.calc_psnr(sr, hr))
    out["psnr_benchmark_x4"] = float(ref_utils.calc_psnr(sr, hr, dataset="benchmark", scale=4))
    out["psnr_div2k_x3"] = float(ref_utils.calc_psnr(sr, hr, dataset="div2k", scale=3))
    out["psnr_range255"] = float(ref_utils.calc_psnr(sr * 255, hr * 255, rgb_range=255))
    a = sr[0].permute(1, 2, 0).numpy() * 255.0
    b = hr[0].permute(1, 2, 0).numpy() * 255.0
    out["ssim_rgb"] = float(ref_utils.calculate_ssim(a, b))
    out["ssim_gray"] = float(ref_utils.calculate_ssim(a[:, :, 1], b[:, :, 1]))
    import imresize as ref_imresize   # LINF-LP/imresize.py
    out["lr_x4"] = ref_imresize.imresize(sr[0].permute(1, 2, 0).numpy(), 1 / 4)
    out["lr_x3"] = ref_imresize.imresize(sr[1].permute(1, 2, 0).numpy(), 1 / 3)
    out["up_x2"] = ref_imresize.imresize(sr[0, :, :12, :10].permute(1, 2, 0).numpy(), 2)
    # uint8 branch of the same imresize (SRFlow-LP/code/test.py:159: LR consistency of a uint8 SR image)
    sr8 = (sr.permute(0, 2, 3, 1).numpy() * 255).astype(np.uint8)
    out["sr8"] = sr8
    out["lr8_x4"] = ref_imresize.imresize(sr8[0], 1 / 4)
    out["lr8_x3"] = ref_imresize.imresize(sr8[1], 1 / 3)
    out["lr8_x8"] = ref_imresize.imresize(sr8[0][:40, :56], 1 / 8)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "metrics.npz"), **out)
    print({k: v for k, v in out.items() if not hasattr(v, "shape")})


if __name__ == "__main__":
    main()
