"""Deterministic synthetic workloads: inputs and checkpoints in the reference's layouts.

Neutral workload generator shared by the tests, bench.py and the oracle scripts; it
contains no model arithmetic and is never imported by the product package
(`bfsr_b200/`).  The reference ships no SRFlow weights
(SRFlow-LP/setup.sh:32-38 downloads them; no network here), so every SRFlow
configuration is exercised with synthetic checkpoints written in the reference's
own state_dict layout:

* key families / shapes follow SRFlow-LP/code/models/modules/RRDBNet_arch.py:64-87,
  FlowUpsamplerNet.py:94-187, FlowStep.py:49-79, FlowAffineCouplingsAblation.py:25-55,
  flow.py:26-83, Split.py:26-37 and SRFlow-LP/code/models/unet.py:109-152;
* the generation script that feeds them to the UNMODIFIED reference
  (`load_state_dict(strict=True)`) is oracle/make_golden.py.

Values are drawn from numpy's legacy RandomState (bit-reproducible across
machines) key by key in sorted-key order, with analytic ActNorm scales so the
flow stays well conditioned in both directions (SURVEY.md App. D explains why a
naive randomisation gives NaNs in the inverse).
"""
from __future__ import annotations

import math
from collections import OrderedDict

import numpy as np
import torch

# --------------------------------------------------------------------------
# topology description shared by the oracle, the synth generator and the tests
# --------------------------------------------------------------------------


class SRFlowTopo:
    """Static description of an SRFlowNet (FlowUpsamplerNet.py:94-115)."""

    def __init__(self, scale=4, nf=64, nb=23, gc=32, K=16, L=3, n_no_affine=2,
                 blocks=(1, 8, 15, 22), hidden=64, fea_up0=True, split=True):
        assert scale in (4, 8)
        self.scale, self.nf, self.nb, self.gc = scale, nf, nb, gc
        self.K, self.L, self.n_no_affine = K, L, n_no_affine
        self.blocks = tuple(blocks)
        self.hidden = hidden
        self.fea_up0 = fea_up0
        self.split = split
        self.n_cond = (len(self.blocks) + 1) * nf  # 320
        # layer list: (kind, C_in, level) ; kind in squeeze/nocoupling/coupling/split
        layers = []
        C = 3
        for level in range(1, L + 1):
            C *= 4
            layers.append(("squeeze", C, level))
            for _ in range(n_no_affine):
                layers.append(("nocoupling", C, level))
            for _ in range(K):
                layers.append(("coupling", C, level))
            # arch_split: `L < levels - correction` with correction = 1
            if split and level < L - 1:
                layers.append(("split", C, level))
                C = C - int(round(C * 0.5))
        self.layers = layers
        self.C_final = C

    def opt(self):
        """The `opt` dict the reference's define_Flow consumes (confs/SRFlow-LP_DF2K_4X.yml)."""
        return {
            "scale": self.scale,
            "network_G": {
                "which_model_G": "SRFlowNet", "in_nc": 3, "out_nc": 3, "nf": self.nf, "nb": self.nb,
                "upscale": self.scale, "train_RRDB": False, "train_RRDB_delay": 0.5,
                "flow": {
                    "K": self.K, "L": self.L, "noInitialInj": True,
                    "coupling": "CondAffineSeparatedAndCond",
                    "additionalFlowNoAffine": self.n_no_affine,
                    "split": {"enable": self.split},
                    "fea_up0": self.fea_up0,
                    "stackRRDB": {"blocks": list(self.blocks), "concat": True},
                },
            },
        }


def srflow_param_shapes(t: SRFlowTopo) -> "OrderedDict[str, tuple]":
    s = OrderedDict()
    nf, gc = t.nf, t.gc
    s["RRDB.conv_first.weight"] = (nf, 3, 3, 3)
    s["RRDB.conv_first.bias"] = (nf,)
    for i in range(t.nb):
        for r in (1, 2, 3):
            for c in range(1, 6):
                cin = nf + (c - 1) * gc
                cout = gc if c < 5 else nf
                p = f"RRDB.RRDB_trunk.{i}.RDB{r}.conv{c}"
                s[p + ".weight"] = (cout, cin, 3, 3)
                s[p + ".bias"] = (cout,)
    ups = ["trunk_conv", "upconv1", "upconv2"]
    if t.scale >= 8:
        ups.append("upconv3")
    ups.append("HRconv")
    for n in ups:
        s[f"RRDB.{n}.weight"] = (nf, nf, 3, 3)
        s[f"RRDB.{n}.bias"] = (nf,)
    s["RRDB.conv_last.weight"] = (3, nf, 3, 3)
    s["RRDB.conv_last.bias"] = (3,)
    H = t.hidden
    for i, (kind, C, _lvl) in enumerate(t.layers):
        p = f"flowUpsamplerNet.layers.{i}"
        if kind in ("nocoupling", "coupling"):
            s[p + ".actnorm.bias"] = (1, C, 1, 1)
            s[p + ".actnorm.logs"] = (1, C, 1, 1)
            s[p + ".invconv.weight"] = (C, C)
        if kind == "coupling":
            for name, cin, cout in (("fAffine", C // 2 + t.n_cond, (C - C // 2) * 2),
                                    ("fFeatures", t.n_cond, C * 2)):
                q = f"{p}.affine.{name}"
                s[q + ".0.weight"] = (H, cin, 3, 3)
                s[q + ".0.actnorm.bias"] = (1, H, 1, 1)
                s[q + ".0.actnorm.logs"] = (1, H, 1, 1)
                s[q + ".2.weight"] = (H, H, 1, 1)
                s[q + ".2.actnorm.bias"] = (1, H, 1, 1)
                s[q + ".2.actnorm.logs"] = (1, H, 1, 1)
                s[q + ".4.weight"] = (cout, H, 3, 3)
                s[q + ".4.bias"] = (cout,)
                s[q + ".4.logs"] = (cout, 1, 1)
        if kind == "split":
            cons = int(round(C * 0.5))
            s[p + ".conv.weight"] = (cons * 2, C - cons, 3, 3)
            s[p + ".conv.bias"] = (cons * 2,)
            s[p + ".conv.logs"] = (cons * 2, 1, 1)
    fo = 2 * 3 * 64 // 2 // 2 if t.split else 2 * 3 * 64
    s["flowUpsamplerNet.f.0.weight"] = (fo, t.n_cond, 3, 3)
    s["flowUpsamplerNet.f.0.bias"] = (fo,)
    return s


def _orthogonal(rs: np.random.RandomState, n: int) -> np.ndarray:
    """Product of n Householder reflections (plain fp64 arithmetic, no LAPACK)."""
    q = np.eye(n)
    for _ in range(min(n, 12)):
        v = rs.randn(n)
        v /= math.sqrt(float((v * v).sum()))
        q = q - 2.0 * np.outer(q @ v, v)
    return q


def synth_srflow_state_dict(t: SRFlowTopo, seed: int = 0, ft_std: float = None):
    """Well-conditioned random SRFlowNet weights (fp32 torch tensors, CPU)."""
    if ft_std is None:
        # every RRDB returns ~1.2x its input (out*0.2 + x around near-identity RDBs, RRDBNet_arch.py:53-57),
        # so the conditioning features grow as 1.2**nb; the hidden ActNorms are scaled to absorb it
        ft_std = 0.4 * 1.2 ** t.nb
    rs = np.random.RandomState(seed)
    sd = OrderedDict()
    shapes = srflow_param_shapes(t)
    for k in sorted(shapes):
        shp = shapes[k]
        if k.startswith("RRDB."):
            if k.endswith(".weight"):
                fan_in = shp[1] * 9
                if ".RDB" in k:
                    std = math.sqrt(2.0 / fan_in) * 0.03  # kaiming fan_in, damped so the 69 residual adds keep |fea| O(1)
                else:
                    std = math.sqrt(1.0 / fan_in)
                v = rs.randn(*shp) * std
            else:
                v = rs.randn(*shp) * 0.02
        elif k.endswith("invconv.weight"):
            n = shp[0]
            v = _orthogonal(rs, n) + 0.02 * rs.randn(n, n)
        elif k.endswith(".actnorm.bias") and ".affine." not in k:
            v = rs.randn(*shp) * 0.1
        elif k.endswith(".actnorm.logs") and ".affine." not in k:
            # compensates the mean contraction of the two affine couplings so |z| stays O(1)
            v = 0.17 + rs.randn(*shp) * 0.05
        elif ".affine." in k:
            if k.endswith(".0.weight") or k.endswith(".2.weight"):
                v = rs.randn(*shp) * 0.05  # flow.Conv2d weight_std
            elif k.endswith(".0.actnorm.logs"):
                fan_in = shapes[k.replace(".actnorm.logs", ".weight")][1] * 9
                v = -math.log(0.05 * math.sqrt(fan_in) * ft_std) + rs.randn(*shp) * 0.05
            elif k.endswith(".2.actnorm.logs"):
                v = -math.log(0.05 * 8.0 * 0.6) + rs.randn(*shp) * 0.05
            elif k.endswith(".actnorm.bias"):
                v = rs.randn(*shp) * 0.05
            elif k.endswith(".4.weight"):
                v = rs.randn(*shp) * 0.02
            elif k.endswith(".4.bias"):
                v = rs.randn(*shp) * 0.02
            elif k.endswith(".4.logs"):
                v = rs.randn(*shp) * 0.05
            else:
                raise KeyError(k)
        elif ".conv." in k:  # Split2d's Conv2dZeros
            if k.endswith(".weight"):
                v = rs.randn(*shp) * 0.02
            elif k.endswith(".bias"):
                v = rs.randn(*shp) * 0.02
            else:
                v = rs.randn(*shp) * 0.05
        elif k.startswith("flowUpsamplerNet.f.0"):
            v = rs.randn(*shp) * 0.01  # unused by the forward pass, must exist (FlowUpsamplerNet.py:107-110)
        else:
            raise KeyError(k)
        sd[k] = torch.from_numpy(np.ascontiguousarray(v, dtype=np.float32))
    return sd


# --------------------------------------------------------------------------
# SRFlow-LP prior (SRFlow-LP/code/models/unet.py:109-152)
# --------------------------------------------------------------------------


def unet_srflow_param_shapes(depth=3, dim=64, bilinear=True, latent_ch=(6, 96)):
    s = OrderedDict()
    factor = 2 if bilinear else 1

    def dense(p, nf, gc, out):
        for c in range(1, 6):
            cin = nf + (c - 1) * gc
            cout = gc if c < 5 else out
            s[f"{p}.conv{c}.weight"] = (cout, cin, 3, 3)
            s[f"{p}.conv{c}.bias"] = (cout,)

    def dconv(p, cin, cout, mid=None):
        mid = mid or cout
        for j, (a, b) in zip((0, 3), ((cin, mid), (mid, cout))):
            s[f"{p}.double_conv.{j}.weight"] = (b, a, 3, 3)
            s[f"{p}.double_conv.{j+1}.weight"] = (b,)
            s[f"{p}.double_conv.{j+1}.bias"] = (b,)
            s[f"{p}.double_conv.{j+1}.running_mean"] = (b,)
            s[f"{p}.double_conv.{j+1}.running_var"] = (b,)
            s[f"{p}.double_conv.{j+1}.num_batches_tracked"] = ()

    for b, nf in enumerate(latent_ch):
        dense(f"input_proj{b}", nf, dim, dim)
    for b in range(len(latent_ch)):
        for i in range(depth):
            cout = dim * 2 ** (i + 1) // (factor if i == depth - 1 else 1)
            dconv(f"down_layers{b}.{i}.maxpool_conv.1", dim * 2 ** i, cout)
        for i in range(depth):
            cin = dim * 2 ** (depth - i)
            cout = dim * 2 ** (depth - i - 1) // (factor if i < depth - 1 else 1)
            assert bilinear
            dconv(f"up_layers{b}.{i}.conv", cin, cout, cin // 2)
    for b in range(len(latent_ch)):
        dconv(f"inc{b}", dim, dim)
    for b, nf in enumerate(latent_ch):
        s[f"outc{b}.conv.weight"] = (nf, dim, 1, 1)
        s[f"outc{b}.conv.bias"] = (nf,)
    return s


def synth_unet_state_dict(shapes, seed: int = 1):
    rs = np.random.RandomState(seed)
    sd = OrderedDict()
    for k in sorted(shapes):
        shp = shapes[k]
        if k.endswith("num_batches_tracked"):
            sd[k] = torch.tensor(100, dtype=torch.int64)
            continue
        if k.endswith("running_mean"):
            v = rs.randn(*shp) * 0.1
        elif k.endswith("running_var"):
            v = 0.5 + rs.rand(*shp)
        elif len(shp) == 4:
            fan_in = shp[1] * shp[2] * shp[3]
            std = math.sqrt(2.0 / fan_in)
            if "input_proj" in k or "lr_proj" in k:
                std *= 0.5
            if k.startswith("outc"):
                # an untrained prior must stay near the flow's mode: O(1) random latents drive the untrained inverse
                # into its exploding regime (scale -> 1e-4 saturation; inf on ~1/4 of 160x160 tiles when undamped)
                std *= 0.1
            v = rs.randn(*shp) * std
        elif k.endswith(".weight"):  # BN gamma
            v = 1.0 + rs.randn(*shp) * 0.1
        else:  # biases / BN beta
            v = rs.randn(*shp) * (0.005 if k.startswith("outc") else 0.05)
        sd[k] = torch.from_numpy(np.ascontiguousarray(v, dtype=np.float32))
    return sd


# --------------------------------------------------------------------------
# inputs
# --------------------------------------------------------------------------


def img(B: int, h: int, w: int, seed: int) -> torch.Tensor:
    """Seeded image-like LR batch in [0,1], fp32 NCHW."""
    g = torch.Generator().manual_seed(seed)
    base = torch.rand(B, 3, max(h // 4, 1), max(w // 4, 1), generator=g)
    x = torch.nn.functional.interpolate(base, size=(h, w), mode="bicubic", align_corners=False)
    x = x.clamp(0, 1) + 0.02 * torch.randn(B, 3, h, w, generator=g)
    return x.clamp(0, 1).contiguous()


# --------------------------------------------------------------------------
# LINF-LP (LINF-LP/models/linf.py:221-242, edsr.py:107-132, rrdb.py:79-103, flow.py:12-27, unet.py:105-142)
# --------------------------------------------------------------------------


def linf_param_shapes(encoder="edsr-baseline", hidden=256, flow_layers=10, ps=3, nb=23):
    s = OrderedDict()
    D = 3 * ps * ps
    if encoder == "edsr-baseline":
        for n in ("sub_mean", "add_mean"):
            s[f"encoder.{n}.weight"] = (3, 3, 1, 1)
            s[f"encoder.{n}.bias"] = (3,)
        s["encoder.head.0.weight"] = (64, 3, 3, 3)
        s["encoder.head.0.bias"] = (64,)
        for i in range(16):
            for j in (0, 2):
                s[f"encoder.body.{i}.body.{j}.weight"] = (64, 64, 3, 3)
                s[f"encoder.body.{i}.body.{j}.bias"] = (64,)
        s["encoder.body.16.weight"] = (64, 64, 3, 3)
        s["encoder.body.16.bias"] = (64,)
    else:
        s["encoder.conv_first.weight"] = (64, 3, 3, 3)
        s["encoder.conv_first.bias"] = (64,)
        for i in range(nb):
            for r in (1, 2, 3):
                for c in range(1, 6):
                    p = f"encoder.RRDB_trunk.{i}.RDB{r}.conv{c}"
                    s[p + ".weight"] = (32 if c < 5 else 64, 64 + (c - 1) * 32, 3, 3)
                    s[p + ".bias"] = (32 if c < 5 else 64,)
        for n in ("trunk_conv", "upconv1", "upconv2", "HRconv"):
            s[f"encoder.{n}.weight"] = (64, 64, 3, 3)
            s[f"encoder.{n}.bias"] = (64,)
        s["encoder.conv_last.weight"] = (3, 64, 3, 3)
        s["encoder.conv_last.bias"] = (3,)
    for n in ("coef", "freq"):
        s[f"{n}.weight"] = (hidden, 64, 3, 3)
        s[f"{n}.bias"] = (hidden,)
    s["phase.weight"] = (hidden // 2, 2)
    dims = [(hidden * 4, hidden), (hidden, hidden), (hidden, hidden), (hidden, flow_layers * D * 2)]
    for k, (a, b) in zip((0, 2, 4, 6), dims):
        s[f"layers.{k}.weight"] = (b, a, 1, 1)
        s[f"layers.{k}.bias"] = (b,)
    for i in range(flow_layers):
        s[f"imnet.linears.{i}.bias"] = (D,)
        s[f"imnet.linears.{i}._weight"] = (D, D)
    s["imnet.last.bias"] = (D,)
    s["imnet.last._weight"] = (D, D)
    return s


def synth_linf_state_dict(shapes, seed=5):
    rs = np.random.RandomState(seed)
    sd = OrderedDict()
    for k in sorted(shapes):
        shp = shapes[k]
        if "_mean." in k:
            v = np.eye(3).reshape(3, 3, 1, 1) if k.endswith("weight") else np.zeros(3)
        elif k.endswith("._weight"):
            v = _orthogonal(rs, shp[0]) + 0.03 * rs.randn(*shp)
        elif len(shp) == 4:
            fan_in = shp[1] * shp[2] * shp[3]
            std = math.sqrt(1.0 / fan_in)
            if ".RDB" in k:
                std = math.sqrt(2.0 / fan_in) * 0.03
            if ".body." in k and k.endswith(".body.2.weight"):
                std *= 0.3       # keep the 16 residual adds of EDSR O(1)
            if k.startswith("layers.6"):
                std *= 0.5
            if k.startswith("encoder.trunk_conv"):
                std *= 0.02      # near-identity RDBs make every RRDB a x1.2 gain (x + 0.2 x): keep fea + trunk O(1) after 23 of them
            v = rs.randn(*shp) * std
        elif k == "phase.weight":
            v = rs.randn(*shp) * 0.5
        else:
            v = rs.randn(*shp) * 0.02
        sd[k] = torch.from_numpy(np.ascontiguousarray(v, dtype=np.float32))
    return sd


def unet_linf_param_shapes(in_chans=27, depth=3, dim=64):
    s = OrderedDict()
    gc = dim // 2

    def dense(p, nf):
        for c in range(1, 6):
            s[f"{p}.conv{c}.weight"] = (gc, nf + (c - 1) * gc, 3, 3)
            s[f"{p}.conv{c}.bias"] = (gc,)

    def dconv(p, cin, cout, mid=None):
        mid = mid or cout
        for j, (a, b) in zip((0, 3), ((cin, mid), (mid, cout))):
            s[f"{p}.double_conv.{j}.weight"] = (b, a, 3, 3)
            for n, shp in (("weight", (b,)), ("bias", (b,)), ("running_mean", (b,)), ("running_var", (b,)),
                           ("num_batches_tracked", ())):
                s[f"{p}.double_conv.{j + 1}.{n}"] = shp

    dense("input_proj", in_chans)
    s["lr_proj.0.weight"] = (in_chans, 3, 3, 3)
    s["lr_proj.0.bias"] = (in_chans,)
    dense("lr_proj.2", in_chans)
    for i in range(depth):
        dconv(f"down_layers.{i}.maxpool_conv.1", dim * 2 ** i, dim * 2 ** (i + 1) // (2 if i == depth - 1 else 1))
    for i in range(depth):
        cin = dim * 2 ** (depth - i)
        dconv(f"up_layers.{i}.conv", cin, dim * 2 ** (depth - i - 1) // (2 if i < depth - 1 else 1), cin // 2)
    dconv("inc", dim, dim)
    s["outc.conv.weight"] = (in_chans, dim, 1, 1)
    s["outc.conv.bias"] = (in_chans,)
    return s
