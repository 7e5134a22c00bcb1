"""Round-2 fixtures from the UNMODIFIED reference at the BASELINE shapes (build container only; /root/reference does not
exist on the GPU box).  Test infrastructure: nothing under bfsr_b200/ imports this.

    python -m oracle.make_golden_r2 wrappers     # LINF dataset wrappers (datasets/wrappers.py:155-238, 517-613) -> linf_wrappers.npz
    python -m oracle.make_golden_r2 linf         # LINF-LP: 48x48 B=2 (config 3), RRDB x6 / x8 (config 5 scales), real checkpoints
    python -m oracle.make_golden_r2 srflow160    # SRFlow-LP config-2 tile (160x160 LR, shipped topology), strided record
    python -m oracle.make_golden_r2 srflow_x8    # config-4 topology (8x, K=16, L=4, nb=23) encode / decode on a 40x40 tile
    python -m oracle.make_golden_r2 flowstep     # P1: single FlowStep / Split2d modules of the reference, both directions
    python -m oracle.make_golden_r2 srflow_x8_lp # config 4 LP path: reference SRFlowNet + three-branch prior built from the reference's blocks

Large outputs are recorded on a stride (every 4th pixel) plus full-resolution corner / centre crops: every recorded value
is an output of the reference itself, and the GPU tests compare the same positions of the engine's output.
"""
from __future__ import annotations

import os
import random
import sys

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLD = os.path.join(ROOT, "tests", "golden")
REF = "/root/reference"


def _cpu_cuda():
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.nn.Module.cuda = lambda self, *a, **k: self


# ----------------------------------------------------------------------------------------- LINF dataset wrappers
class _ListDS:
    def __init__(self, items):
        self.items = items

    def __len__(self):
        return len(self.items)

    def __getitem__(self, i):
        return self.items[i]


WRAPPER_CASES = [
    # (kind, lr h, lr w, scale)   paired: SRImplicitPairedFastPatch ; down: SRImplicitDownsampledFastPatchTest
    ("paired", 24, 24, 4), ("paired", 17, 23, 4), ("paired", 16, 20, 2), ("paired", 48, 48, 4), ("paired", 15, 12, 3),
    ("down", 20, 16, 3), ("down", 24, 24, 2), ("down", 12, 10, 3.5), ("down", 24, 24, 6), ("down", 24, 24, 8), ("down", 11, 13, 4),
]


def _ref_linf():
    sys.path.insert(0, os.path.join(HERE, "stubs"))
    sys.path.insert(0, os.path.join(REF, "LINF-LP"))
    _cpu_cuda()
    import models as ref_models          # noqa: E402
    import test as ref_test              # noqa: E402
    from datasets import wrappers        # noqa: E402
    return ref_models, ref_test, wrappers


def wrapper_item(wrappers, kind, lr_or_hr, scale):
    """One item of the reference's test-time wrapper.  paired: lr_or_hr is the LR image (the HR partner only fixes the
    shape); down: it is the HR image, the wrapper makes the LR itself (PIL bicubic, wrappers.py:241-244)."""
    random.seed(0)
    if kind == "paired":
        lr = lr_or_hr
        hr = torch.zeros(3, lr.shape[-2] * scale, lr.shape[-1] * scale)
        ds = wrappers.SRImplicitPairedFastPatch(_ListDS([(lr, hr)]), patch_size=3)
    else:
        ds = wrappers.SRImplicitDownsampledFastPatchTest(_ListDS([lr_or_hr]), scale_min=scale, scale_max=scale, patch_size=3)
    return ds[0]


def golden_wrappers():
    from tools import synth
    _, _, wrappers = _ref_linf()
    out = {"n": np.int64(len(WRAPPER_CASES))}
    for i, (kind, h, w, s) in enumerate(WRAPPER_CASES):
        if kind == "paired":
            src = synth.img(1, h, w, 400 + i)[0]
        else:
            src = synth.img(1, round(h * s), round(w * s), 400 + i)[0]
        it = wrapper_item(wrappers, kind, src, s)
        assert tuple(it["inp"].shape) == (3, h, w), (it["inp"].shape, h, w)
        out[f"c{i}_meta"] = np.array([kind == "paired", h, w], dtype=np.int64)
        out[f"c{i}_scale"] = np.float64(s)
        out[f"c{i}_lr01"] = it["inp"].numpy()
        out[f"c{i}_coord"] = it["coord"].numpy()
        out[f"c{i}_cell"] = it["cell"].numpy()
        out[f"c{i}_gt_lr_up"] = it["gt_lr_up"].numpy()
        out[f"c{i}_hw"] = np.array(it["gt"].shape[-2:], dtype=np.int64)
        print(kind, h, w, s, "coord", tuple(it["coord"].shape), "gt_lr_up", tuple(it["gt_lr_up"].shape), "HW", tuple(it["gt"].shape[-2:]))
    path = os.path.join(GOLD, "linf_wrappers.npz")
    np.savez_compressed(path, **out)
    print("linf_wrappers bytes", os.path.getsize(path))


# ----------------------------------------------------------------------------------------- LINF-LP at the config shapes
def golden_linf_r2():
    from tools import synth
    ref_models, ref_test, wrappers = _ref_linf()
    torch.set_num_threads(8)
    lp = os.path.join(REF, "LINF-LP")
    real = {}
    for enc, f, fp in (("edsr-baseline", "edsr-baseline-linf.pth", "edsr-baseline-linf-LP.pth"), ("rrdb", "rrdb-linf.pth", "rrdb-linf-LP.pth")):
        real[enc] = (torch.load(os.path.join(lp, f), map_location="cpu")["model"],
                     torch.load(os.path.join(lp, fp), map_location="cpu")["prior_model"])
    cases = {
        # name: (encoder, wrapper kind, B, lr h, lr w, scale, seed)
        "linf_edsr_real_x4_48": ("edsr-baseline", "paired", 2, 48, 48, 4, 311),     # BASELINE config 3 geometry (q = 65)
        "linf_rrdb_real_x6": ("rrdb", "down", 1, 24, 24, 6, 312),                   # BASELINE config 5 scales with rrdb-linf.pth
        "linf_rrdb_real_x8": ("rrdb", "down", 1, 24, 24, 8, 313),
    }
    only = os.environ.get("BFSR_GOLDEN_ONLY")
    for name, (enc, kind, B, h, w, s, seed) in cases.items():
        if only and name not in only.split(","):
            continue
        model = ref_models.make(real[enc][0], load_sd=True).eval()
        prior = ref_models.make(real[enc][1], load_sd=True).eval()
        src = synth.img(B, h, w, seed) if kind == "paired" else synth.img(B, h * s, w * s, seed)
        items = [wrapper_item(wrappers, kind, src[i], s) for i in range(B)]
        # LINF-LP/test.py:98 (data_norm inp sub 0.5 div 0.5), then :143-171
        inp = torch.stack([(it["inp"] - 0.5) / 0.5 for it in items]); coord = torch.stack([it["coord"] for it in items])
        cell = torch.stack([it["cell"] for it in items]); gt_lr_up = torch.stack([it["gt_lr_up"] for it in items])
        H, W = items[0]["gt"].shape[-2:]
        with torch.no_grad():
            z_lr = ref_test.batched_predict_log_p(model, inp, coord, cell, gt_lr_up).detach().contiguous()
            z_learned = prior(z_lr, inp)
            if z_learned.shape != z_lr.shape:
                z_learned = F.interpolate(z_learned, size=z_lr.shape[-2:], mode="bilinear", align_corners=False)
            pred = ref_test.batched_predict(model, inp, coord, cell, 0, z_learned)
            pred = pred[..., :H, :W]
            pred = pred + F.interpolate(inp, pred.shape[-2:], mode="bilinear", align_corners=False)
        path = os.path.join(GOLD, name + ".npz")
        np.savez_compressed(path, lr01=torch.stack([it["inp"] for it in items]).numpy(), z_lr=z_lr.numpy(),
                            z_learned_s2=z_learned[..., ::2, ::2].contiguous().numpy(), pred=pred.numpy(),
                            scale=np.float64(s), meta=np.array([B, h, w, int(kind == "paired"), seed], dtype=np.int64))
        print(name, "q", tuple(z_lr.shape[-2:]), "pred", tuple(pred.shape), "range", float(pred.min()), float(pred.max()),
              "z_lr std", float(z_lr.std()), "bytes", os.path.getsize(path))


# ----------------------------------------------------------------------------------------- SRFlow-LP
def _ref_srflow():
    sys.path.insert(0, os.path.join(HERE, "stubs"))
    sys.path.insert(0, os.path.join(REF, "SRFlow-LP", "code"))
    import models.networks as networks  # noqa
    import models as ref_models  # noqa
    import options.options as option  # noqa
    return networks, ref_models, option


def strided(t, s=4):
    return t[..., ::s, ::s].contiguous().numpy()


def golden_srflow160():
    """BASELINE config 2, tile 0 of the bench batch (synth.img(32,160,160,1236)[0], weights seed 0 / prior seed 1): the
    reference's own LP path (test.py:135-148) on the CPU, ~15 s."""
    from tools import synth
    networks, ref_models, option = _ref_srflow()
    torch.set_num_threads(8)
    topo = synth.SRFlowTopo()
    net = networks.define_Flow(option.dict_to_nonedict(topo.opt()), 0)
    net.load_state_dict(synth.synth_srflow_state_dict(topo, seed=0), strict=True)
    net.eval()
    usd = synth.synth_unet_state_dict(synth.unet_srflow_param_shapes(), seed=1)
    prior = ref_models.make({"name": "unet", "args": {"depth": 3, "dim": 64, "bilinear": True}, "sd": usd}, load_sd=True).eval()
    lr = synth.img(32, 160, 160, 1236)[:1].contiguous()
    with torch.no_grad():
        lr_up = F.interpolate(lr, scale_factor=4, mode="bilinear", align_corners=False)
        epses_lr = []
        net(gt=lr_up, lr=lr, reverse=False, epses=epses_lr, add_gt_noise=False)
        epses = [e.detach() for e in epses_lr]
        for i in range(len(epses)):
            mean = torch.mean(epses[i], dim=[1], keepdim=True)
            std = torch.std(epses[i], dim=[1], keepdim=True)
            epses[i] = (epses[i] - mean) / (std + 1e-8)
        learned = prior(epses)
        sr, _ = net(lr=lr, z=None, eps_std=None, reverse=True, epses=learned, reverse_with_grad=True)
    out = {"meta": np.array([1, 160, 160, 0, 1236], dtype=np.int64), "sr_s4": strided(sr), "sr_tl": sr[..., :64, :64].numpy(),
           "sr_c": sr[..., 288:352, 288:352].numpy(), "sr_br": sr[..., -64:, -64:].numpy(),
           "eps_lr0_s4": strided(epses_lr[0]), "eps_lr1_s4": strided(epses_lr[1]),
           "learned0_s4": strided(learned[0]), "learned1_s4": strided(learned[1]),
           "sr_sum": np.float64(sr.double().sum().item()), "sr_l2": np.float64(sr.double().norm().item())}
    path = os.path.join(GOLD, "srflow_full160.npz")
    np.savez_compressed(path, **out)
    print("srflow_full160 sr range", float(sr.min()), float(sr.max()), "bytes", os.path.getsize(path))


def golden_srflow_x8():
    """BASELINE config 4 topology (8x, RRDB nb=23, K=16, L=4, two Split2d) on one 40x40 tile: encode of bilinear(lr), decode of 0.9 x those latents,
    and the round trip.  The learned prior of this configuration is an extension the reference cannot run (unet.py:117-118)."""
    from tools import synth
    networks, ref_models, option = _ref_srflow()
    torch.set_num_threads(8)
    topo = synth.SRFlowTopo(scale=8, L=4)
    net = networks.define_Flow(option.dict_to_nonedict(topo.opt()), 0)
    net.load_state_dict(synth.synth_srflow_state_dict(topo, seed=31), strict=True)
    net.eval()
    lr = synth.img(1, 40, 40, 231)
    with torch.no_grad():
        lr_up = F.interpolate(lr, scale_factor=8, mode="bilinear", align_corners=False)
        epses = []
        net(gt=lr_up, lr=lr, reverse=False, epses=epses, add_gt_noise=False)
        lat = [0.9 * e for e in epses]      # decode input: the image's own latents at 0.9x amplitude (random latents of any
        #                                     useful amplitude overflow the untrained 64-step inverse: "exploding inverse")
        sr, _ = net(lr=lr, z=None, eps_std=None, reverse=True, epses=list(lat), reverse_with_grad=True)
        rt, _ = net(lr=lr, z=None, eps_std=None, reverse=True, epses=list(epses), reverse_with_grad=True)
    out = {"meta": np.array([1, 40, 40, 31, 231], dtype=np.int64), "lr": lr.numpy(), "sr_s2": strided(sr, 2),
           "sr_tl": sr[..., :48, :48].numpy(), "roundtrip_maxabs": np.float32((rt - lr_up).abs().max().item()),
           "shapes": np.array([list(e.shape) for e in epses], dtype=np.int64)}
    for i, e in enumerate(epses):
        out[f"eps{i}"] = e.numpy()         # full: 0.9 * eps is the decode input of the test
    path = os.path.join(GOLD, "srflow_x8_k16.npz")
    np.savez_compressed(path, **out)
    print("srflow_x8_k16 latents", [tuple(e.shape) for e in epses], "sr range", float(sr.min()), float(sr.max()), "roundtrip",
          out["roundtrip_maxabs"], "bytes", os.path.getsize(path))


def ref_block_prior(ref_unet, latent_ch, depth=3, dim=64):
    """The reference's SRFlow-LP UNet (unet.py:109-181) generalised from its hard-coded two latents (6 / 96 channels, :117-118,
    :151-152) to one branch per latent, built from the reference's OWN blocks (DenseBlock_5C, DoubleConv, Down, Up, OutConv,
    unet.py:10-107) under the reference's attribute names -- a documented extension: the reference cannot run config 4's prior."""
    import torch.nn as nn

    class UNetN(nn.Module):
        def __init__(self):
            super().__init__()
            self.depth = depth
            for b, nf in enumerate(latent_ch):
                setattr(self, f"input_proj{b}", ref_unet.DenseBlock_5C(nf=nf, gc=dim, out_dim=dim, bias=True))
                down, up = nn.ModuleList(), nn.ModuleList()
                for i in range(depth):
                    down.append(ref_unet.Down(dim * 2 ** i, dim * 2 ** (i + 1) // (2 if i == depth - 1 else 1)))
                for i in range(depth):
                    up.append(ref_unet.Up(dim * 2 ** (depth - i), dim * 2 ** (depth - i - 1) // (2 if i < depth - 1 else 1), True))
                setattr(self, f"down_layers{b}", down); setattr(self, f"up_layers{b}", up)
                setattr(self, f"inc{b}", ref_unet.DoubleConv(dim, dim))
                setattr(self, f"outc{b}", ref_unet.OutConv(dim, nf))
            self.n = len(latent_ch)

        def forward(self, epses):
            out = []
            for b in range(self.n):
                z = getattr(self, f"inc{b}")(getattr(self, f"input_proj{b}")(epses[b]))
                feats = [z]
                for layer in getattr(self, f"down_layers{b}"):
                    z = layer(z); feats.append(z)
                for idx, layer in enumerate(getattr(self, f"up_layers{b}")):
                    z = layer(z, feats[self.depth - 1 - idx])
                out.append(getattr(self, f"outc{b}")(z))
            return out

    return UNetN()


def golden_srflow_x8_lp():
    """BASELINE config 4, whole LP path on one 40x40 tile: the reference's SRFlowNet (8x, L=4, nb=23, K=8) + the three-branch prior built
    from the reference's blocks (`ref_block_prior`): encode -> normalise -> prior -> decode (test.py:135-148)."""
    from tools import synth
    networks, ref_models, option = _ref_srflow()
    from models import unet as ref_unet
    torch.set_num_threads(8)
    # K = 8: with UNTRAINED synthetic weights the 64-step inverse of K = 16 overflows for any non-zero prior output (probed: NaN even
    # at 0.1 x the prior's latents; exact zeros give |SR| ~ 70), the 32-step one stays finite.  K = 16 encode / decode: srflow_x8_k16.
    topo = synth.SRFlowTopo(scale=8, L=4, K=8)
    net = networks.define_Flow(option.dict_to_nonedict(topo.opt()), 0)
    net.load_state_dict(synth.synth_srflow_state_dict(topo, seed=31), strict=True)
    net.eval()
    lat_ch = (6, 12, 192)
    prior = ref_block_prior(ref_unet, lat_ch)
    usd = synth.synth_unet_state_dict(synth.unet_srflow_param_shapes(latent_ch=lat_ch), seed=32)
    prior.load_state_dict(usd, strict=True)
    prior.eval()
    lr = synth.img(1, 40, 40, 231)
    with torch.no_grad():
        lr_up = F.interpolate(lr, scale_factor=8, mode="bilinear", align_corners=False)
        epses_lr = []
        net(gt=lr_up, lr=lr, reverse=False, epses=epses_lr, add_gt_noise=False)
        epses = [e.detach() for e in epses_lr]
        for i in range(len(epses)):
            mean = torch.mean(epses[i], dim=[1], keepdim=True)
            std = torch.std(epses[i], dim=[1], keepdim=True)
            epses[i] = (epses[i] - mean) / (std + 1e-8)
        learned = prior(epses)
        sr, _ = net(lr=lr, z=None, eps_std=None, reverse=True, epses=learned, reverse_with_grad=True)
    out = {"meta": np.array([1, 40, 40, 31, 231, 32], dtype=np.int64), "sr_s2": strided(sr, 2), "sr_tl": sr[..., :48, :48].numpy()}
    for i, e in enumerate(learned):
        out[f"learned{i}_s2"] = strided(e, 2)
    path = os.path.join(GOLD, "srflow_x8_lp.npz")
    np.savez_compressed(path, **out)
    print("srflow_x8_lp sr range", float(sr.min()), float(sr.max()), "finite", bool(torch.isfinite(sr).all()), "learned std",
          [float(e.std()) for e in learned], "bytes", os.path.getsize(path))


def module_inputs(idx, C, H, W, split=False):
    """Seeded inputs of the module-level cases (regenerated by the tests, not stored): z, and ft (320 ch) or eps."""
    g = torch.Generator().manual_seed(1000 + idx)
    z = torch.randn(2, C, H, W, generator=g)
    other = torch.randn(2, C // 2, H, W, generator=g) if split else torch.randn(2, 320, H, W, generator=g) * 0.5
    return z, other


MODULE_HW = {12: (12, 10), 24: (9, 14), 96: (6, 5)}


def golden_flowstep():
    """P1: single modules of the reference with the synthetic weights of one layer of the small topology --
    FlowStep.normal_flow / reverse_flow (FlowStep.py:88-129) for a coupling and a no-coupling step at C = 12, 24, 96, and
    Split2d forward / inverse (Split.py:49-77) -- on random z / ft."""
    from tools import synth
    networks, ref_models, option = _ref_srflow()
    torch.set_num_threads(8)
    topo = synth.SRFlowTopo(nb=4, blocks=(0, 1, 2, 3), K=2)
    net = networks.define_Flow(option.dict_to_nonedict(topo.opt()), 0)
    sd = synth.synth_srflow_state_dict(topo, seed=11)
    net.load_state_dict(sd, strict=True)
    net.eval()
    layers = net.flowUpsamplerNet.layers
    out = {}
    names = []
    for idx, layer in enumerate(layers):
        cls = type(layer).__name__
        if cls == "FlowStep":
            C = layer.actnorm.bias.shape[1]
            H, W = MODULE_HW[C]
            z, ft = module_inputs(idx, C, H, W)
            with torch.no_grad():
                zf, _ = layer(z, logdet=torch.zeros(2), reverse=False, rrdbResults=ft)
                zi, _ = layer(z, logdet=torch.zeros(2), reverse=True, rrdbResults=ft)
            out[f"l{idx}_fwd"] = zf.numpy(); out[f"l{idx}_inv"] = zi.numpy()
            names.append((idx, 1 if layer.flow_coupling != "noCoupling" else 0, C))
            print("FlowStep", idx, layer.flow_coupling, C, "fwd std", float(zf.std()), "inv std", float(zi.std()))
        elif cls == "Split2d":
            C = 12
            z, eps = module_inputs(idx, C, 12, 10, split=True)
            with torch.no_grad():
                z1, _, e = _split_fwd(layer, z)
                zr = _split_inv(layer, z[:, :C // 2], eps)
            out[f"l{idx}_z1"] = z1.numpy(); out[f"l{idx}_eps"] = e.numpy(); out[f"l{idx}_zinv"] = zr.numpy()
            names.append((idx, 2, C))
            print("Split2d", idx, "eps std", float(e.std()))
    out["layers"] = np.array(names, dtype=np.int64)
    path = os.path.join(GOLD, "srflow_modules.npz")
    np.savez_compressed(path, **out)
    print("srflow_modules bytes", os.path.getsize(path))


def _split_fwd(layer, z):
    r = layer(z, logdet=0., reverse=False, eps=None, eps_std=None, ft=None)
    # Split2d.forward returns (z1, logdet, eps) in the encode direction (Split.py:49-61)
    return r[0], r[1], r[2]


def _split_inv(layer, z1, eps):
    r = layer(z1, logdet=0., reverse=True, eps=eps, eps_std=None, ft=None)
    return r[0]


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else ""
    os.makedirs(GOLD, exist_ok=True)
    {"wrappers": golden_wrappers, "linf": golden_linf_r2, "srflow160": golden_srflow160, "srflow_x8": golden_srflow_x8,
     "flowstep": golden_flowstep, "srflow_x8_lp": golden_srflow_x8_lp}.get(which, lambda: sys.exit(__doc__))()
