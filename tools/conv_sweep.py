"""Per-shape throughput of the conv kernels (CUDA events around the kernel only, via the library's profiling hooks)."""
import ctypes as C
import sys

import torch

sys.path.insert(0, ".")
from bfsr_b200 import _lib  # noqa: E402

SHAPES = [  # (B, H, W, Cin, Cout, label) at the config-2 chunk size (8 tiles of 160x160 LR)
    (8, 160, 160, 64, 32, "RDB conv1"), (8, 160, 160, 96, 32, "RDB conv2"), (8, 160, 160, 128, 32, "RDB conv3"),
    (8, 160, 160, 160, 32, "RDB conv4"), (8, 160, 160, 192, 64, "RDB conv5"),
    (8, 320, 320, 320, 1024, "L1 ft conv (16 steps batched)"), (8, 160, 160, 320, 1024, "L2 ft conv"),
    (8, 80, 80, 320, 1024, "L3 ft conv"), (8, 320, 320, 64, 24, "L1 fFeatures.4"), (8, 320, 320, 64, 12, "L1 fAffine.4"),
    (8, 80, 80, 64, 192, "L3 fFeatures.4"), (8, 320, 320, 72, 64, "UNet0 dense conv2"), (8, 320, 320, 264, 64, "UNet0 dense conv5"),
    (8, 160, 160, 128, 128, "UNet0 down0"), (8, 320, 320, 128, 64, "UNet0 up2"),
]


def run(shape, impl, reps=3):
    B, H, W, cin, cout, _ = shape
    L = _lib.lib()
    x = torch.randn(B, cin, H, W, device="cuda")
    w = torch.randn(cout, cin, 3, 3) / (cin * 9) ** 0.5
    b = torch.zeros(cout)
    y = torch.empty(B, cout, H, W, device="cuda")
    best = 1e30
    for _ in range(reps):
        L.bfsr_prof_enable(1)
        _lib.check(L.bfsr_op_conv2d(x.data_ptr(), B, cin, H, W, w.data_ptr(), b.data_ptr(), cout, 3, 0, impl, y.data_ptr(), None))
        ms, work, n = C.c_double(), C.c_double(), C.c_int64()
        L.bfsr_prof_summary(1 if impl else 0, C.byref(ms), C.byref(work), C.byref(n))
        L.bfsr_prof_enable(0)
        best = min(best, ms.value)
    return best, work.value


if __name__ == "__main__":
    print("| conv | B,H,W | Cin->Cout | fp32 ms (TF/s) | bf16x3 ms (TF/s) | bf16 ms (TF/s) |\n|---|---|---|---|---|---|")
    for sh in SHAPES:
        cells = []
        for impl in (0, 1, 2):
            ms, work = run(sh, impl)
            cells.append(f"{ms:.3f} ({work / ms / 1e9:.0f})")
        print(f"| {sh[5]} | {sh[0]},{sh[1]},{sh[2]} | {sh[3]}->{sh[4]} | " + " | ".join(cells) + " |", flush=True)
