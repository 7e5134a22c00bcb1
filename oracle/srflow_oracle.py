"""CPU oracle for the SRFlow-LP hot path — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A plain fp32 torch-on-CPU restatement of the reference's algorithm (the path of
SRFlow-LP/code/test.py:135-148), operating directly on a flat state_dict in the
reference's checkpoint layout.  Only `tests/`, `__graft_entry__.smoke()` and
`bench.py`'s cpu_baseline / `--impl reference` leg may import this module; the
product path (`bfsr_b200/`) never does and fails loudly without its CUDA library.

Pinning: the reference's own test suite holds no golden vectors (it has no
tests).  This oracle is pinned instead against OUTPUTS OF THE
UNMODIFIED REFERENCE run in the build container (oracle/make_golden.py imports
/root/reference, loads the same synthetic state_dict with strict=True and
records inputs/outputs under tests/golden/); tests/test_oracle_golden.py replays
them.  Every function cites the reference lines it restates (paths relative to
/root/reference/SRFlow-LP/code/).

`literal=True` executes exactly the reference's amount of work (encoder run in
both passes, dead RRDB heads computed) and is what the CPU baseline times.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from tools.synth import SRFlowTopo


# ----------------------------------------------------------------- RRDB encoder
def _rdb(sd, p, x):
    """ResidualDenseBlock_5C.forward — models/modules/RRDBNet_arch.py:36-42."""
    def c(i, t):
        return F.conv2d(t, sd[f"{p}.conv{i}.weight"], sd[f"{p}.conv{i}.bias"], padding=1)
    x1 = F.leaky_relu(c(1, x), 0.2)
    x2 = F.leaky_relu(c(2, torch.cat((x, x1), 1)), 0.2)
    x3 = F.leaky_relu(c(3, torch.cat((x, x1, x2), 1)), 0.2)
    x4 = F.leaky_relu(c(4, torch.cat((x, x1, x2, x3), 1)), 0.2)
    x5 = c(5, torch.cat((x, x1, x2, x3, x4), 1))
    return x5 * 0.2 + x


def rrdb_forward(sd, t: SRFlowTopo, lr, literal=False, prefix="RRDB."):
    """RRDBNet.forward(x, get_steps=True) — RRDBNet_arch.py:89-148.

    Note the aliasing of the in-place LeakyReLU (RRDBNet_arch.py:87,105-109): the
    `fea_up2` / `fea_up4` entries handed to the flow are POST-activation.
    """
    def conv(name, x):
        return F.conv2d(x, sd[f"{prefix}{name}.weight"], sd[f"{prefix}{name}.bias"], padding=1)
    fea = conv("conv_first", lr)
    first = fea
    blocks = {}
    for i in range(t.nb):
        x = fea
        out = x
        for r in (1, 2, 3):
            out = _rdb(sd, f"{prefix}RRDB_trunk.{i}.RDB{r}", out)
        fea = out * 0.2 + x  # RRDB.forward :53-57
        if i in t.blocks:
            blocks[f"block_{i}"] = fea
    trunk = conv("trunk_conv", fea)
    last_lr_fea = fea + trunk          # :103 (fea is the trunk output, not conv_first)
    del first
    res = {"last_lr_fea": last_lr_fea, "fea_up1": last_lr_fea}
    fea_up2 = F.leaky_relu(conv("upconv1", F.interpolate(last_lr_fea, scale_factor=2, mode="nearest")), 0.2)
    res["fea_up2"] = fea_up2
    need_up4 = literal or t.scale >= 8
    if need_up4:
        fea_up4 = F.leaky_relu(conv("upconv2", F.interpolate(fea_up2, scale_factor=2, mode="nearest")), 0.2)
        res["fea_up4"] = fea_up4
        f = fea_up4
        if t.scale >= 8:
            fea_up8 = F.leaky_relu(conv("upconv3", F.interpolate(fea_up4, scale_factor=2, mode="nearest")), 0.2)
            res["fea_up8"] = fea_up8
            f = fea_up8
        if literal:
            res["out"] = conv("conv_last", F.leaky_relu(conv("HRconv", f), 0.2))
    if t.fea_up0:
        res["fea_up0"] = F.interpolate(last_lr_fea, scale_factor=0.5, mode="bilinear", align_corners=False,
                                       recompute_scale_factor=True)
    res.update(blocks)
    return res


def rrdb_preprocessing(sd, t: SRFlowTopo, lr, literal=False):
    """SRFlowNet.rrdbPreprocessing — SRFlowNet_arch.py:118-138."""
    r = rrdb_forward(sd, t, lr, literal=literal)
    concat = torch.cat([r[f"block_{i}"] for i in t.blocks], dim=1)
    keys = ["last_lr_fea", "fea_up1", "fea_up2", "fea_up4"]
    if "fea_up0" in r:
        keys.append("fea_up0")
    if t.scale >= 8:
        keys.append("fea_up8")
    for k in keys:
        if k not in r:
            continue
        h, w = r[k].shape[2:]
        r[k] = torch.cat([r[k], F.interpolate(concat, (h, w))], dim=1)
    return r


def level_to_name(t: SRFlowTopo):
    """FlowUpsamplerNet.__init__ levelToName — FlowUpsamplerNet.py:58-74."""
    if t.scale == 8:
        return {0: "fea_up8", 1: "fea_up4", 2: "fea_up2", 3: "fea_up1", 4: "fea_up0"}
    return {0: "fea_up4", 1: "fea_up2", 2: "fea_up1", 3: "fea_up0", 4: "fea_up-1"}


# ----------------------------------------------------------------- flow pieces
def squeeze2d(x):
    """flow.squeeze2d — flow.py:122-134."""
    B, C, H, W = x.shape
    x = x.view(B, C, H // 2, 2, W // 2, 2).permute(0, 1, 3, 5, 2, 4).contiguous()
    return x.view(B, C * 4, H // 2, W // 2)


def unsqueeze2d(x):
    """flow.unsqueeze2d — flow.py:137-152."""
    B, C, H, W = x.shape
    x = x.view(B, C // 4, 2, 2, H, W).permute(0, 1, 4, 2, 5, 3).contiguous()
    return x.view(B, C // 4, H * 2, W * 2)


def _conv_an(sd, p, x, k):
    """flow.Conv2d (bias-free conv then ActNorm fwd) — flow.py:41-65, FlowActNorms.py:61-92."""
    y = F.conv2d(x, sd[p + ".weight"], None, padding=k // 2)
    return (y + sd[p + ".actnorm.bias"]) * torch.exp(sd[p + ".actnorm.logs"])


def _conv_zeros(sd, p, x):
    """flow.Conv2dZeros — flow.py:68-83."""
    y = F.conv2d(x, sd[p + ".weight"], sd[p + ".bias"], padding=1)
    return y * torch.exp(sd[p + ".logs"] * 3)


def _f(sd, p, x):
    """CondAffineSeparatedAndCond.F — FlowAffineCouplingsAblation.py:127-135."""
    h = F.relu(_conv_an(sd, p + ".0", x, 3))
    h = F.relu(_conv_an(sd, p + ".2", h, 1))
    return _conv_zeros(sd, p + ".4", h)


def _scale_shift(h, eps=1e-4):
    """feature_extract(_aff) — FlowAffineCouplingsAblation.py:108-119; thops.py:59-60 ('cross')."""
    shift, scale = h[:, 0::2], h[:, 1::2]
    return torch.sigmoid(scale + 2.0) + eps, shift


def flowstep_forward(sd, p, z, ft, coupling):
    """FlowStep.normal_flow — FlowStep.py:88-111."""
    z = (z + sd[p + ".actnorm.bias"]) * torch.exp(sd[p + ".actnorm.logs"])
    W = sd[p + ".invconv.weight"]
    z = F.conv2d(z, W.view(*W.shape, 1, 1))                     # Permutations.py:39,48
    if coupling:
        C = z.shape[1]
        scaleFt, shiftFt = _scale_shift(_f(sd, p + ".affine.fFeatures", ft))
        z = (z + shiftFt) * scaleFt
        z1, z2 = z[:, :C // 2], z[:, C // 2:]
        scale, shift = _scale_shift(_f(sd, p + ".affine.fAffine", torch.cat([z1, ft], 1)))
        z2 = (z2 + shift) * scale
        z = torch.cat([z1, z2], 1)
    return z


def flowstep_reverse(sd, p, z, ft, coupling):
    """FlowStep.reverse_flow — FlowStep.py:113-129."""
    if coupling:
        C = z.shape[1]
        z1, z2 = z[:, :C // 2], z[:, C // 2:]
        scale, shift = _scale_shift(_f(sd, p + ".affine.fAffine", torch.cat([z1, ft], 1)))
        z2 = z2 / scale - shift
        z = torch.cat([z1, z2], 1)
        scaleFt, shiftFt = _scale_shift(_f(sd, p + ".affine.fFeatures", ft))
        z = z / scaleFt - shiftFt
    W = sd[p + ".invconv.weight"]
    Winv = torch.inverse(W.double()).float()                    # Permutations.py:41-42
    z = F.conv2d(z, Winv.view(*W.shape, 1, 1))
    z = z * torch.exp(-sd[p + ".actnorm.logs"]) - sd[p + ".actnorm.bias"]
    return z


def split2d_forward(sd, p, z):
    """Split2d.forward(reverse=False), ft=None — Split.py:49-61."""
    C = z.shape[1]
    cons = int(round(C * 0.5))
    z1, z2 = z[:, :C - cons], z[:, C - cons:]
    h = _conv_zeros(sd, p + ".conv", z1)
    mean, logs = h[:, 0::2], h[:, 1::2]
    eps = (z2 - mean) / torch.exp(logs)
    return z1, eps


def split2d_reverse(sd, p, z1, eps):
    """Split2d.forward(reverse=True) — Split.py:62-77."""
    h = _conv_zeros(sd, p + ".conv", z1)
    mean, logs = h[:, 0::2], h[:, 1::2]
    z2 = mean + torch.exp(logs) * eps
    return torch.cat([z1, z2], 1)


def flow_encode(sd, t: SRFlowTopo, gt, rr):
    """FlowUpsamplerNet.encode — FlowUpsamplerNet.py:217-251.  Returns [eps_split..., z_final]."""
    names = level_to_name(t)
    z = gt
    epses = []
    for i, (kind, _C, level) in enumerate(t.layers):
        p = f"flowUpsamplerNet.layers.{i}"
        if kind == "squeeze":
            z = squeeze2d(z)
        elif kind == "split":
            z, eps = split2d_forward(sd, p, z)
            epses.append(eps)
        else:
            z = flowstep_forward(sd, p, z, rr[names[level]], kind == "coupling")
    epses.append(z)
    return epses


def flow_decode(sd, t: SRFlowTopo, epses, rr):
    """FlowUpsamplerNet.decode — FlowUpsamplerNet.py:267-296 (pops from the end, never mutates the caller's list)."""
    names = level_to_name(t)
    epses = list(epses)
    z = epses.pop()
    for i in reversed(range(len(t.layers))):
        kind, _C, level = t.layers[i]
        p = f"flowUpsamplerNet.layers.{i}"
        if kind == "squeeze":
            z = unsqueeze2d(z)
        elif kind == "split":
            z = split2d_reverse(sd, p, z, epses.pop())
        else:
            z = flowstep_reverse(sd, p, z, rr[names[level]], kind == "coupling")
    assert z.shape[1] == 3
    return z


# ----------------------------------------------------------------- prior (UNet)
def _dense5(sd, p, x):
    """DenseBlock_5C.forward — models/unet.py:30-36."""
    def c(i, t):
        return F.conv2d(t, sd[f"{p}.conv{i}.weight"], sd[f"{p}.conv{i}.bias"], padding=1)
    x1 = F.leaky_relu(c(1, x), 0.2)
    x2 = F.leaky_relu(c(2, torch.cat((x, x1), 1)), 0.2)
    x3 = F.leaky_relu(c(3, torch.cat((x, x1, x2), 1)), 0.2)
    x4 = F.leaky_relu(c(4, torch.cat((x, x1, x2, x3), 1)), 0.2)
    return c(5, torch.cat((x, x1, x2, x3, x4), 1))


def _double_conv(sd, p, x):
    """DoubleConv — models/unet.py:38-56 (BatchNorm in eval mode)."""
    for j in (0, 3):
        x = F.conv2d(x, sd[f"{p}.double_conv.{j}.weight"], None, padding=1)
        q = f"{p}.double_conv.{j + 1}"
        x = F.batch_norm(x, sd[q + ".running_mean"], sd[q + ".running_var"], sd[q + ".weight"], sd[q + ".bias"],
                         False, 0.1, 1e-5)
        x = F.leaky_relu(x, 0.2)
    return x


def _up(sd, p, x1, x2):
    """Up.forward (bilinear) — models/unet.py:84-98."""
    x1 = F.interpolate(x1, scale_factor=2, mode="bilinear", align_corners=True)
    dy, dx = x2.shape[2] - x1.shape[2], x2.shape[3] - x1.shape[3]
    x1 = F.pad(x1, [dx // 2, dx - dx // 2, dy // 2, dy - dy // 2])
    return _double_conv(sd, p + ".conv", torch.cat([x2, x1], 1))


def unet_body(sd, b, z, depth):
    """One branch of UNet.forward after input_proj — models/unet.py:158-179."""
    z = _double_conv(sd, f"inc{b}", z)
    feats = [z]
    for i in range(depth):
        z = _double_conv(sd, f"down_layers{b}.{i}.maxpool_conv.1", F.max_pool2d(z, 2))
        feats.append(z)
    for i in range(depth):
        z = _up(sd, f"up_layers{b}.{i}", z, feats[depth - 1 - i])
    return F.conv2d(z, sd[f"outc{b}.conv.weight"], sd[f"outc{b}.conv.bias"])


def unet_srflow_forward(sd, epses, depth=3):
    """SRFlow-LP UNet.forward(epses) — models/unet.py:154-181."""
    out = []
    for b, e in enumerate(epses):
        z = _dense5(sd, f"input_proj{b}", e)
        out.append(unet_body(sd, b, z, depth))
    return out


def normalise_latents(epses):
    """Per-pixel channel normalisation — SRFlow-LP/code/test.py:141-145 (unbiased std)."""
    out = []
    for e in epses:
        mean = torch.mean(e, dim=[1], keepdim=True)
        std = torch.std(e, dim=[1], keepdim=True)
        out.append((e - mean) / (std + 1e-8))
    return out


# ----------------------------------------------------------------- whole paths
@torch.no_grad()
def encode(sd, t, lr, gt, literal=False):
    """SRFlowModel.get_encode_z(lq, gt, epses=[], add_gt_noise=False) — SRFlow_model.py:201-206."""
    rr = rrdb_preprocessing(sd, t, lr, literal=literal)
    return flow_encode(sd, t, gt, rr)


@torch.no_grad()
def decode(sd, t, lr, epses, literal=False):
    """SRFlowModel.get_sr(lq, epses=...) — SRFlow_model.py:198-199,215-222."""
    rr = rrdb_preprocessing(sd, t, lr, literal=literal)
    return flow_decode(sd, t, epses, rr)


@torch.no_grad()
def lp_sr(sd, prior_sd, t, lr, literal=True, return_all=False):
    """The LP inference path of SRFlow-LP/code/test.py:135-148 (pre-clamp SR)."""
    lr_up = F.interpolate(lr, scale_factor=t.scale, mode="bilinear", align_corners=False)
    if literal:
        epses_lr = encode(sd, t, lr, lr_up, literal=True)
        epses = normalise_latents(epses_lr)
        learned = unet_srflow_forward(prior_sd, epses)
        sr = decode(sd, t, lr, learned, literal=True)
    else:
        rr = rrdb_preprocessing(sd, t, lr)
        epses_lr = flow_encode(sd, t, lr_up, rr)
        epses = normalise_latents(epses_lr)
        learned = unet_srflow_forward(prior_sd, epses)
        sr = flow_decode(sd, t, learned, rr)
    if return_all:
        return sr, epses_lr, epses, learned
    return sr
