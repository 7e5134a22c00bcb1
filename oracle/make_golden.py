"""Generate tests/golden/*.npz by running the UNMODIFIED reference in this container.

Run (build container only — /root/reference does not exist on the GPU box):

    python -m oracle.make_golden srflow      # SRFlow-LP fixtures
    python -m oracle.make_golden linf        # LINF-LP fixtures (real shipped checkpoints)

The reference is imported from /root/reference with the stub modules under
oracle/stubs/ on sys.path (missing offline deps) and `.cuda()`
neutralised; no reference file is modified or copied.  The synthetic SRFlow
checkpoints come from tools/synth.py and are loaded with strict=True, which also
pins the state_dict key/shape layout.  Fixtures hold only inputs and outputs
(weights are regenerated from their seed by the tests).
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLD = os.path.join(ROOT, "tests", "golden")
REF = "/root/reference"


def _ref_srflow_modules():
    sys.path.insert(0, os.path.join(HERE, "stubs"))
    sys.path.insert(0, os.path.join(REF, "SRFlow-LP", "code"))
    import models.networks as networks  # noqa
    import models as ref_models  # noqa
    import options.options as option  # noqa
    return networks, ref_models, option


def golden_srflow():
    from tools import synth
    networks, ref_models, option = _ref_srflow_modules()
    torch.manual_seed(0)
    torch.set_num_threads(8)

    cases = {
        # name: (topology kwargs, B, h, w, weight seed, input seed)
        "srflow_small": (dict(nb=4, blocks=(0, 1, 2, 3), K=2), 2, 24, 20, 11, 101),
        "srflow_full40": (dict(), 1, 40, 40, 0, 1235),   # BASELINE config 1 (shipped yml topology)
    }
    for name, (kw, B, h, w, wseed, iseed) in cases.items():
        topo = synth.SRFlowTopo(**kw)
        opt = option.dict_to_nonedict(topo.opt())
        net = networks.define_Flow(opt, 0)
        sd = synth.synth_srflow_state_dict(topo, seed=wseed)
        net.load_state_dict(sd, strict=True)
        net.eval()
        ushapes = synth.unet_srflow_param_shapes()
        usd = synth.synth_unet_state_dict(ushapes, seed=wseed + 1)
        prior = ref_models.make({"name": "unet", "args": {"depth": 3, "dim": 64, "bilinear": True}, "sd": usd},
                                load_sd=True)
        prior.eval()

        lr = synth.img(B, h, w, iseed)
        with torch.no_grad():
            # SRFlow-LP/code/test.py:135-148 with the wrapper calls unrolled (SRFlow_model.py:201-222)
            lr_up = torch.nn.functional.interpolate(lr, scale_factor=topo.scale, mode="bilinear", align_corners=False)
            epses_lr = []
            net(gt=lr_up, lr=lr, reverse=False, epses=epses_lr, add_gt_noise=False)
            epses = [e.detach() for e in epses_lr]
            for i in range(len(epses)):
                mean = torch.mean(epses[i], dim=[1], keepdim=True)
                std = torch.std(epses[i], dim=[1], keepdim=True)
                epses[i] = (epses[i] - mean) / (std + 1e-8)
            learned = prior(epses)
            sr, _ = net(lr=lr, z=None, eps_std=None, reverse=True, epses=learned, reverse_with_grad=True)
            # invertibility probe: decode(encode(x)) on the un-normalised latents
            rt, _ = net(lr=lr, z=None, eps_std=None, reverse=True, epses=epses_lr, reverse_with_grad=True)
        out = {"lr": lr.numpy(), "sr": sr.numpy(), "roundtrip_maxabs": np.float32((rt - lr_up).abs().max().item())}
        for i, e in enumerate(epses_lr):
            out[f"eps_lr{i}"] = e.numpy()
        for i, e in enumerate(learned):
            out[f"learned{i}"] = e.numpy()
        out["meta"] = np.array([B, h, w, wseed, iseed], dtype=np.int64)
        path = os.path.join(GOLD, name + ".npz")
        np.savez_compressed(path, **out)
        print(name, "sr range", float(sr.min()), float(sr.max()), "eps std", [float(e.std()) for e in epses_lr],
              "learned std", [float(e.std()) for e in learned], "roundtrip", out["roundtrip_maxabs"],
              "bytes", os.path.getsize(path))


def golden_srflow_x8():
    """8x topology (BASELINE config 4 family: L = 4, conditioning by fea_up4 / fea_up2 / fea_up1 / fea_up0, two Split2d): encode
    and decode only -- the shipped SRFlow-LP prior hard-codes the 4x latent layout (unet.py:117-118)."""
    from tools import synth
    networks, ref_models, option = _ref_srflow_modules()
    torch.set_num_threads(8)
    kw, B, h, w, wseed, iseed = dict(scale=8, L=4, nb=2, blocks=(0, 1, 0, 1), K=1), 1, 16, 12, 21, 201
    topo = synth.SRFlowTopo(**kw)
    net = networks.define_Flow(option.dict_to_nonedict(topo.opt()), 0)
    sd = synth.synth_srflow_state_dict(topo, seed=wseed)
    net.load_state_dict(sd, strict=True)
    net.eval()
    lr = synth.img(B, h, w, iseed)
    with torch.no_grad():
        lr_up = torch.nn.functional.interpolate(lr, scale_factor=8, mode="bilinear", align_corners=False)
        epses = []
        net(gt=lr_up, lr=lr, reverse=False, epses=epses, add_gt_noise=False)
        half = [0.5 * e for e in epses]
        sr, _ = net(lr=lr, z=None, eps_std=None, reverse=True, epses=list(half), reverse_with_grad=True)
        rt, _ = net(lr=lr, z=None, eps_std=None, reverse=True, epses=list(epses), reverse_with_grad=True)
    out = {"lr": lr.numpy(), "sr_half": sr.numpy(), "roundtrip_maxabs": np.float32((rt - lr_up).abs().max().item()),
           "meta": np.array([B, h, w, wseed, iseed], dtype=np.int64)}
    for i, e in enumerate(epses):
        out[f"eps{i}"] = e.numpy()
    path = os.path.join(GOLD, "srflow_x8_small.npz")
    np.savez_compressed(path, **out)
    print("srflow_x8_small latents", [tuple(e.shape) for e in epses], "sr range", float(sr.min()), float(sr.max()),
          "roundtrip", out["roundtrip_maxabs"], "bytes", os.path.getsize(path))


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "srflow"
    os.makedirs(GOLD, exist_ok=True)
    if which == "srflow":
        golden_srflow()
    elif which == "srflow_x8":
        golden_srflow_x8()
    elif which == "linf":
        from oracle.make_golden_linf import golden_linf
        golden_linf()
    else:
        raise SystemExit("usage: python -m oracle.make_golden [srflow|srflow_x8|linf]")
