"""Run one conv through the C ABI (for ncu captures): python tools/conv_one.py B H W Cin Cout impl [reps]"""
import sys

import torch

sys.path.insert(0, ".")
from bfsr_b200 import _lib  # noqa: E402

B, H, W, cin, cout, impl = [int(v) for v in sys.argv[1:7]]
reps = int(sys.argv[7]) if len(sys.argv) > 7 else 2
L = _lib.lib()
x = torch.randn(B, cin, H, W, device="cuda")
w = torch.randn(cout, cin, 3, 3) / (cin * 9) ** 0.5
b = torch.zeros(cout)
y = torch.empty(B, cout, H, W, device="cuda")
for _ in range(reps):
    _lib.check(L.bfsr_op_conv2d(x.data_ptr(), B, cin, H, W, w.data_ptr(), b.data_ptr(), cout, 3, 0, impl, y.data_ptr(), None))
torch.cuda.synchronize()
print("ok", float(y.abs().mean()))
