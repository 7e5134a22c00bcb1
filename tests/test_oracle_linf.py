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


# ------------------------------------------------------------------ round-2 pins (oracle/make_golden_r2.py)
def wrapper_cases():
    g = golden("linf_wrappers")
    for i in range(int(g["n"])):
        paired, h, w = [int(v) for v in g[f"c{i}_meta"]]
        s = float(g[f"c{i}_scale"])
        yield (i, bool(paired), h, w, int(s) if s == int(s) else s, torch.from_numpy(g[f"c{i}_lr01"]), torch.from_numpy(g[f"c{i}_coord"]),
               torch.from_numpy(g[f"c{i}_cell"]), torch.from_numpy(g[f"c{i}_gt_lr_up"]), tuple(int(v) for v in g[f"c{i}_hw"]))


def test_build_inputs_matches_reference_wrappers():
    """b7: the restated input construction equals the items of the reference's OWN dataset wrappers
    (SRImplicitPairedFastPatch wrappers.py:155-238, SRImplicitDownsampledFastPatchTest :517-613) fed an in-memory dataset --
    coord / cell / gt_lr_up bit for bit, including the always-pad rule of the paired wrapper and non-integer scales."""
    n = 0
    for i, paired, h, w, s, lr01, coord, cell, gt, hw in wrapper_cases():
        inp, r_coord, r_cell, r_gt, r_hw = LO.build_inputs(lr01, s, 3, always_pad=paired)
        assert tuple(r_hw) == hw, (i, r_hw, hw)
        assert torch.equal(inp, (lr01 - 0.5) / 0.5)
        assert torch.equal(r_coord, coord), i
        assert torch.equal(r_cell, cell), i
        assert torch.equal(r_gt, gt), i
        n += 1
    assert n >= 10


R2_CASES = {"linf_edsr_real_x4_48": "edsr-baseline", "linf_rrdb_real_x6": "rrdb", "linf_rrdb_real_x8": "rrdb"}


def load_case_r2(name):
    """Cases recorded through the reference's wrappers: the fixture carries the wrapper's LR (`inp` before normalisation)."""
    g = golden(name)
    B, h, w, paired, seed = [int(v) for v in g["meta"]]
    s = float(g["scale"]); s = int(s) if s == int(s) else s
    enc = R2_CASES[name]
    path = os.path.join(GOLD, "_linf_ckpt", enc + ".pt")
    if not os.path.exists(path):
        pytest.skip("real LINF checkpoints not exported (run `python -m oracle.make_golden linf` in the build container)")
    ck = torch.load(path, map_location="cpu")
    lr01 = torch.from_numpy(g["lr01"])
    ins = [LO.build_inputs(lr01[i], s, 3, bool(paired)) for i in range(B)]
    inp = torch.stack([x[0] for x in ins]); coord = torch.stack([x[1] for x in ins])
    cell = torch.stack([x[2] for x in ins]); gt = torch.stack([x[3] for x in ins])
    return g, enc, ck["model"]["sd"], ck["prior_model"]["sd"], lr01, s, bool(paired), inp, coord, cell, gt, ins[0][4]


@pytest.mark.parametrize("name", list(R2_CASES))
def test_linf_oracle_matches_reference_at_config_shapes(name):
    """Config 3 geometry (48x48 LR, q = 65, B = 2, real EDSR weights) and the config-5 scales 6 and 8 with rrdb-linf.pth."""
    g, enc, sd, psd, lr01, s, paired, inp, coord, cell, gt, hw = load_case_r2(name)
    pred, z_lr, z_learned = LO.lp_sr(sd, psd, enc, inp, coord, cell, gt, hw, literal=False, return_all=True)
    assert rel_l2(g["z_lr"], z_lr) < 1e-5
    assert rel_l2(g["z_learned_s2"], z_learned[..., ::2, ::2]) < 1e-5
    assert rel_l2(g["pred"], pred) < 1e-5
