"""Per-tap (impl 3) vs dx-folded (impl 6 / 5) evaluation of the small-Cout 3x3 convs at the config-2 chunk shapes:
    python tools/f3_sweep.py > gpurun_out/f3_sweep.md"""
import ctypes as C
import sys

import torch

sys.path.insert(0, ".")
from bfsr_b200 import _lib  # noqa: E402

SHAPES = [(32, 160, 160, 64, 32, "RDB conv1"), (32, 160, 160, 96, 32, "RDB conv2"), (32, 160, 160, 128, 32, "RDB conv3"),
          (32, 160, 160, 160, 32, "RDB conv4"), (32, 320, 320, 64, 12, "L1 fAffine.4"), (32, 320, 320, 64, 24, "L1 fFeatures.4"),
          (32, 160, 160, 64, 24, "L2 fAffine.4")]


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
        _lib.check(L.bfsr_op_conv2d(x.data_ptr(), B, cin, H, W, w.data_ptr(), b.data_ptr(), cout, 3, 1, impl, y.data_ptr(), None))
        ms, work, n = C.c_double(), C.c_double(), C.c_int64()
        L.bfsr_prof_summary(1, C.byref(ms), C.byref(work), C.byref(n))
        L.bfsr_prof_enable(0)
        best = min(best, ms.value)
    return best, work.value


if __name__ == "__main__":
    print("| conv | B,H,W | Cin->Cout | per-tap ms (TF/s) | dx-folded bf16x2-out ms (TF/s) | dx-folded fp32-out ms (TF/s) |\n|---|---|---|---|---|---|")
    for sh in SHAPES:
        cells = []
        for impl in (3, 6, 5):
            ms, work = run(sh, impl)
            cells.append(f"{ms:.3f} ({work / ms / 1e9:.0f})")
        print(f"| {sh[5]} | {sh[0]},{sh[1]},{sh[2]} | {sh[3]}->{sh[4]} | " + " | ".join(cells) + " |", flush=True)
