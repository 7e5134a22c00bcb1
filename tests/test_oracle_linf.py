"""CPU: the LINF-LP oracle reproduces the outputs recorded from the UNMODIFIED reference (real shipped checkpoints when
available in this container or exported under tests/golden/_linf_ckpt/, synthetic weights always)."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import linf_oracle as LO
from tests.util import GOLD, golden, max_abs, rel_l2
from tools import synth

CASES = {"linf_edsr_real_x4": "edsr-baseline", "linf_edsr_real_x3": "edsr-baseline", "linf_rrdb_real_x2": "rrdb",
         "linf_edsr_synth_x4": "synth", "linf_rrdb_synth_x2": "synth-rrdb"}


def load_case(name):
    g = golden(name)
    B, h, w, s, always_pad, seed = [int(v) for v in g["meta"]]
    kind = CASES[name]
    if kind in ("synth", "synth-rrdb"):
        enc = "edsr-baseline" if kind == "synth" else "rrdb"
        sd = synth.synth_linf_state_dict(synth.linf_param_shapes(enc), seed=5 if kind == "synth" else 7)
        psd = synth.synth_unet_state_dict(synth.unet_linf_param_shapes(), seed=6)
    else:
        enc = kind
        path = os.path.join(GOLD, "_linf_ckpt", enc + ".pt")
        if not os.path.exists(path):
            pytest.skip("real LINF checkpoints not exported (run `python -m oracle.make_golden linf` in the build container)")
        ck = torch.load(path, map_location="cpu")
        sd, psd = ck["model"]["sd"], ck["prior_model"]["sd"]
    lr01 = synth.img(B, h, w, seed)
    assert torch.equal(lr01, torch.from_numpy(g["lr01"]))
    ins = [LO.build_inputs(lr01[i], s, 3, bool(always_pad)) for i in range(B)]
    inp = torch.stack([x[0] for x in ins]); coord = torch.stack([x[1] for x in ins])
    cell = torch.stack([x[2] for x in ins]); gt = torch.stack([x[3] for x in ins])
    return g, enc, sd, psd, inp, coord, cell, gt, ins[0][4]


@pytest.mark.parametrize("name", list(CASES))
def test_linf_oracle_matches_reference(name):
    g, enc, sd, psd, inp, coord, cell, gt, hw = load_case(name)
    pred, z_lr, z_learned = LO.lp_sr(sd, psd, enc, inp, coord, cell, gt, hw, literal=False, return_all=True)
    assert rel_l2(g["z_lr"], z_lr) < 1e-5
    assert rel_l2(g["z_learned"], z_learned) < 1e-5
    assert rel_l2(g["pred"], pred) < 1e-5


def test_linf_index_laws():
    """F.fold(k=s=3) == pixel_shuffle(3) (linf.py:401-406); nearest gather index = clamp(rne(((c+1)n-1)/2))."""
    x = torch.arange(2 * 27 * 4 * 5, dtype=torch.float32).view(2, 27, 4, 5)
    f = F.fold(x.view(2, 27, -1), output_size=(12, 15), kernel_size=(3, 3), stride=3)
    assert torch.equal(f, F.pixel_shuffle(x, 3))
    n = 48
    c = torch.linspace(-1 + 1e-6, 1 - 1e-6, 4001)
    grid = torch.stack([c, torch.zeros_like(c)], -1).view(1, 1, -1, 2)
    src = torch.arange(n, dtype=torch.float32).view(1, 1, 1, n)
    got = F.grid_sample(src, grid, mode="nearest", align_corners=False).view(-1)
    idx = torch.clamp(torch.round(((c + 1) * n - 1) / 2), 0, n - 1)
    assert torch.equal(got, idx)
