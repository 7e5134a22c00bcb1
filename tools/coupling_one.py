"""One coupling FlowStep at a config-2 level shape through bfsr_op_flowstep (profiling harness for ncu / step timing):
    python tools/coupling_one.py LEVEL(1|2|3) REVERSE(0|1) [reps] [batch]"""
import sys
import time

import torch

sys.path.insert(0, ".")
from bfsr_b200 import _lib  # noqa: E402
from tools import synth  # noqa: E402

level, rev = int(sys.argv[1]), int(sys.argv[2])
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 8
B = int(sys.argv[4]) if len(sys.argv) > 4 else 32
C, H = {1: (12, 320), 2: (24, 160), 3: (96, 80)}[level]
t = synth.SRFlowTopo(nb=1, blocks=(0, 0, 0, 0), K=1)
sd = synth.synth_srflow_state_dict(t, seed=3)
idx = {1: 3, 2: 8, 3: 12}[level]          # first coupling layer of each level for K = 1, two no-coupling steps per level
table, keep = _lib.tensor_table(sd)
L = _lib.lib()
z = torch.randn(B, C, H, H, device="cuda")
ft = torch.randn(B, 320, H, H, device="cuda") * 0.5
out = torch.empty_like(z)
p = f"flowUpsamplerNet.layers.{idx}".encode()
import ctypes as C_
for it in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    if it == 1:
        L.bfsr_prof_enable(1)
    _lib.check(L.bfsr_op_flowstep(table, len(table), p, C, 1, rev, z.data_ptr(), ft.data_ptr(), B, H, H, out.data_ptr(), 0, reps, None))
    torch.cuda.synchronize()
buf = C_.create_string_buffer(1 << 16)
L.bfsr_prof_dump(buf, len(buf))
L.bfsr_prof_enable(0)
print("ok", float(out.abs().mean()), f"{(time.perf_counter() - t0) * 1e3:.1f} ms wall for packing + ft convs + {reps} steps; per-launch device ms:")
for ln in buf.value.decode().splitlines():
    tag, n, ms, work = ln.split("\t")
    print(f"  {tag:40s} x{n:>3s}  {float(ms) / int(n):.3f} ms each")
