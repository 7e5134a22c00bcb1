"""Device time of the LINF-LP path at BASELINE config 3 (EDSR-baseline 4x, 48x48 LR patches, batch 64, q = 65 queries/side):
    python tools/linf_bench.py [batch] [encoder: edsr-baseline|rrdb]
Synthetic weights of the shipped architecture (timing is data-independent); inputs have the shapes of the paired test wrapper."""
import ctypes as C
import sys

import torch

sys.path.insert(0, ".")
from bfsr_b200 import _lib, models  # noqa: E402
from tools import synth  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
enc = sys.argv[2] if len(sys.argv) > 2 else "edsr-baseline"
h = w = 48; s = 4; q = (h * s) // 3 + 1
sd = synth.synth_linf_state_dict(synth.linf_param_shapes(enc), seed=5)
psd = synth.synth_unet_state_dict(synth.unet_linf_param_shapes(), seed=6)
spec = {"name": "linf-patch", "args": {"encoder_spec": {"name": enc, "args": {"no_upsampling": True}},
                                       "imnet_spec": {"name": "flow", "args": {"name": "flow"}}, "flow_layers": 10,
                                       "num_layer": 3, "hidden_dim": 256, "patch_size": 3}, "sd": sd}
model = models.make(spec, load_sd=True).cuda()
prior = models.make({"name": "unet", "args": {"in_chans": 27, "depth": 3, "dim": 64, "cell_input": False, "bilinear": True},
                     "sd": psd}, load_sd=True).cuda()
inp = (synth.img(B, h, w, 7) - 0.5) / 0.5
c = -1 + (2 * torch.arange(q) + 1) / q
coord = torch.stack(torch.meshgrid(c, c, indexing="ij"), -1)[None].expand(B, q, q, 2).contiguous()
cell = torch.full((B, 2), 2.0 / (h * s))
gt = 0.05 * torch.randn(B, 27, q, q)
args = [t.cuda() for t in (inp, coord, cell, gt)]
L = _lib.lib()
for _ in range(3):
    out = model.lp_sr(*args, prior, (h * s, w * s))
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 20
e0.record()
for _ in range(n):
    out = model.lp_sr(*args, prior, (h * s, w * s))
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
print(f"LINF-LP {enc} 4x, {B} x {h}x{w} LR, q={q}: {ms:.2f} ms/batch -> {B * (h * s) * (w * s) / ms / 1e3:.1f} HR-Mpix/s, finite={bool(torch.isfinite(out).all())}")
L.bfsr_prof_enable(1)
model.lp_sr(*args, prior, (h * s, w * s))
buf = C.create_string_buffer(1 << 20)
L.bfsr_prof_dump(buf, len(buf)); L.bfsr_prof_enable(0)
rows = [ln.split("\t") for ln in buf.value.decode().splitlines()]
rows = sorted(((r[0], int(r[1]), float(r[2])) for r in rows), key=lambda r: -r[2])
print("profiled launches:", sum(r[1] for r in rows), "sum ms", round(sum(r[2] for r in rows), 2))
for tag, cnt, t in rows[:14]:
    print(f"  {tag}\t{cnt}\t{t:.3f}")
