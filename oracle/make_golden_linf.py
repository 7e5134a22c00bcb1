"""LINF-LP fixtures from the UNMODIFIED reference (run via `python -m oracle.make_golden linf`, build container only).

Drives /root/reference/LINF-LP exactly as LINF-LP/test.py:143-171 does with `--patch` and `eval_bsize` set
(`batched_predict_log_p` -> prior -> `batched_predict` -> crop -> + bilinear(inp)), with `.cuda()` neutralised so it
runs on the CPU.  Cases with the REAL shipped checkpoints pin the oracle; the synthetic-weights case exists so the GPU
parity tests have a fixture that does not need the 205 MB checkpoint files.  The checkpoints' model state_dicts are also
exported (fp32, without optimizer state) to tests/golden/_linf_ckpt/ — git-ignored, but shipped to the GPU box with the
working tree — so the GPU tests can run the real-weights cases there.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLD = os.path.join(ROOT, "tests", "golden")
REF = "/root/reference/LINF-LP"


def golden_linf():
    from oracle import linf_oracle as LO
    from tools import synth
    sys.path.insert(0, os.path.join(HERE, "stubs"))
    sys.path.insert(0, REF)
    # run the reference on the CPU: `.cuda()` is hard-coded in linf.py:262,337,398 and test.py
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.nn.Module.cuda = lambda self, *a, **k: self
    import models as ref_models          # noqa: E402  (LINF-LP/models)
    import test as ref_test              # noqa: E402  (LINF-LP/test.py: batched_predict*)
    torch.set_num_threads(8)

    ck = os.path.join(GOLD, "_linf_ckpt")
    os.makedirs(ck, exist_ok=True)
    real = {}
    for enc, f, fp in (("edsr-baseline", "edsr-baseline-linf.pth", "edsr-baseline-linf-LP.pth"),
                       ("rrdb", "rrdb-linf.pth", "rrdb-linf-LP.pth")):
        m = torch.load(os.path.join(REF, f), map_location="cpu")["model"]
        p = torch.load(os.path.join(REF, fp), map_location="cpu")["prior_model"]
        real[enc] = (m, p)
        torch.save({"model": m, "prior_model": p}, os.path.join(ck, enc + ".pt"))

    ssd = synth.synth_linf_state_dict(synth.linf_param_shapes("edsr-baseline"), seed=5)
    spd = synth.synth_unet_state_dict(synth.unet_linf_param_shapes(), seed=6)
    synth_specs = ({"name": "linf-patch", "args": real["edsr-baseline"][0]["args"], "sd": ssd},
                   {"name": "unet", "args": real["edsr-baseline"][1]["args"], "sd": spd})

    # synthetic RRDB-encoder model: keeps the RRDB branch of the engine under test on the GPU box, where the 89 MB real
    # rrdb checkpoint does not travel
    rsd = synth.synth_linf_state_dict(synth.linf_param_shapes("rrdb"), seed=7)
    synth_rrdb_specs = ({"name": "linf-patch", "args": real["rrdb"][0]["args"], "sd": rsd},
                        {"name": "unet", "args": real["rrdb"][1]["args"], "sd": spd})
    only = os.environ.get("BFSR_GOLDEN_ONLY")
    cases = {
        # name: (model spec, prior spec, B, h, w, scale, always_pad, input seed)
        "linf_edsr_real_x4": (*real["edsr-baseline"], 2, 24, 24, 4, True, 301),     # paired wrapper, config-3 shape family
        "linf_edsr_real_x3": (*real["edsr-baseline"], 1, 20, 16, 3, False, 302),    # arbitrary-scale wrapper, ragged
        "linf_rrdb_real_x2": (*real["rrdb"], 1, 16, 16, 2, False, 303),
        "linf_edsr_synth_x4": (*synth_specs, 2, 16, 20, 4, True, 304),
        "linf_rrdb_synth_x2": (*synth_rrdb_specs, 1, 12, 14, 2, False, 305),
    }
    if only:
        cases = {k: v for k, v in cases.items() if k in only.split(",")}
    for name, (mspec, pspec, B, h, w, s, always_pad, seed) in cases.items():
        model = ref_models.make(mspec, load_sd=True).eval()
        prior = ref_models.make(pspec, load_sd=True).eval()
        lr01 = synth.img(B, h, w, seed)
        ins = [LO.build_inputs(lr01[i], s, 3, always_pad) for i in range(B)]
        inp = torch.stack([x[0] for x in ins]); coord = torch.stack([x[1] for x in ins])
        cell = torch.stack([x[2] for x in ins]); gt_lr_up = torch.stack([x[3] for x in ins])
        H, W = ins[0][4]
        with torch.no_grad():
            z_lr = ref_test.batched_predict_log_p(model, inp, coord, cell, gt_lr_up).detach().contiguous()
            z_learned = prior(z_lr, inp)
            if z_learned.shape != z_lr.shape:
                z_learned = torch.nn.functional.interpolate(z_learned, size=z_lr.shape[-2:], mode="bilinear", align_corners=False)
            pred = ref_test.batched_predict(model, inp, coord, cell, 0, z_learned)
            pred = pred[..., :H, :W]
            pred = pred + torch.nn.functional.interpolate(inp, pred.shape[-2:], mode="bilinear", align_corners=False)
            # invertibility probe (P2): query_rgb(zmap = z_lr) reproduces the LR residual patches
            rt = ref_test.batched_predict(model, inp, coord, cell, 0, z_lr)
            rt_err = float((torch.nn.functional.pixel_unshuffle(rt, 3) - gt_lr_up).abs().max())
        path = os.path.join(GOLD, name + ".npz")
        np.savez_compressed(path, lr01=lr01.numpy(), z_lr=z_lr.numpy(), z_learned=z_learned.numpy(), pred=pred.numpy(),
                            meta=np.array([B, h, w, s, int(always_pad), seed], dtype=np.int64),
                            roundtrip_maxabs=np.float32(rt_err))
        print(name, "pred range", float(pred.min()), float(pred.max()), "z_lr std", float(z_lr.std()), "z_learned std",
              float(z_learned.std()), "roundtrip", rt_err, "bytes", os.path.getsize(path))
