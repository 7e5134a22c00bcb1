"""Run one conv through the C ABI (for ncu captures / BFSR_LIB_PATH trace builds): python tools/conv_one.py B H W Cin Cout impl [reps] [ks]"""
import sys

import torch

sys.path.insert(0, ".")
from bfsr_b200 import _lib  # noqa: E402

B, H, W, cin, cout, impl = [int(v) for v in sys.argv[1:7]]
reps = int(sys.argv[7]) if len(sys.argv) > 7 else 2
ks = int(sys.argv[8]) if len(sys.argv) > 8 else 3
L = _lib.lib()
x = torch.randn(B, cin, H, W, device="cuda")
w = torch.randn(cout, cin, ks, ks) / (cin * ks * ks) ** 0.5
b = torch.zeros(cout)
y = torch.empty(B, cout, H, W, device="cuda")
for _ in range(reps):
    _lib.check(L.bfsr_op_conv2d(x.data_ptr(), B, cin, H, W, w.data_ptr(), b.data_ptr(), cout, ks, 0, impl, y.data_ptr(), None))
torch.cuda.synchronize()
print("ok", float(y.abs().mean()))
