"""Per-shape device-time breakdown of one SRFlow-LP step at BASELINE config 2 (CUDA events around every launch):
    python tools/step_profile.py [batch] [precision] [tile_chunk] > gpurun_out/step_profile.tsv"""
import ctypes as C
import sys

import torch

sys.path.insert(0, ".")
from bfsr_b200 import _lib, models  # noqa: E402
from tools import synth  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
prec = int(sys.argv[2]) if len(sys.argv) > 2 else 0
chunk = int(sys.argv[3]) if len(sys.argv) > 3 else 0
t = synth.SRFlowTopo()
sd = synth.synth_srflow_state_dict(t, seed=0)
usd = synth.synth_unet_state_dict(synth.unet_srflow_param_shapes(), seed=1)
net = models.define_Flow(t.opt(), device="cuda:0", precision=prec, tile_chunk=chunk)
net.load_state_dict(sd, strict=True)
prior = models.make({"name": "unet", "args": {"depth": 3, "dim": 64, "bilinear": True}, "sd": usd}, load_sd=True)
lr = synth.img(B, 160, 160, 1236).cuda()
L = _lib.lib()
for _ in range(2):
    net.lp_sr(lr, prior)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); net.lp_sr(lr, prior); e1.record(); torch.cuda.synchronize()
step_ms = e0.elapsed_time(e1)
L.bfsr_prof_enable(1)
net.lp_sr(lr, prior)
buf = C.create_string_buffer(1 << 20)
L.bfsr_prof_dump(buf, len(buf))
L.bfsr_prof_enable(0)
rows = [ln.split("\t") for ln in buf.value.decode().splitlines()]
rows = [(r[0], int(r[1]), float(r[2]), float(r[3])) for r in rows]
rows.sort(key=lambda r: -r[2])
tot = sum(r[2] for r in rows)
print(f"# step {step_ms:.1f} ms unprofiled; sum of per-launch event times {tot:.1f} ms; batch {B} precision {prec}")
print("tag\tlaunches\tms\tshare\twork/ms (GFLOP/s or GB/s)")
for tag, n, ms, work in rows:
    print(f"{tag}\t{n}\t{ms:.3f}\t{ms / tot:.3f}\t{work / ms / 1e6 if ms else 0:.1f}")
