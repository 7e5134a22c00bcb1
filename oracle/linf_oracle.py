"""CPU oracle for the LINF-LP hot path — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Plain fp32 torch-on-CPU restatement of LINF-LP/test.py:143-171 (`--patch` path):
encoder (EDSR-baseline / RRDB) -> coef/freq convs -> 4-neighbour local Fourier features -> 1x1-conv MLP ->
per-query affine parameters -> 27-dim flow forward (LR-derived latent) -> UNet prior -> flow inverse -> fold 3x3
patches -> crop -> + bilinear(LR).  Operates on the flat state_dicts of the shipped checkpoints
(`{'model'|'prior_model': {'name','args','sd'}}`, LINF-LP/train.py:234-248).

Pinned against outputs of the UNMODIFIED reference run in the build container with the REAL shipped checkpoints
(oracle/make_golden_linf.py -> tests/golden/linf_*.npz); only tests/, smoke() and bench.py's CPU legs import it.
Citations are relative to /root/reference/LINF-LP/.
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F


# ----------------------------------------------------------------- input construction (datasets/wrappers.py)
def make_coord(shape, flatten=False):
    """utils.make_coord — utils.py:105-120 (grid centres in [-1,1], 'ij' order, last dim = (row, col))."""
    seqs = []
    for n in shape:
        r = 1.0 / n
        seqs.append(-1 + r + (2 * r) * torch.arange(n).float())
    ret = torch.stack(torch.meshgrid(*seqs, indexing="ij"), dim=-1)
    return ret.view(-1, ret.shape[-1]) if flatten else ret


def build_inputs(lr01, scale, patch_size=3, always_pad=True):
    """Test-time inputs for one LR image in [0,1] (3,h,w) at an integer (or real) scale.

    always_pad=True  -> SRImplicitPairedFastPatch semantics (wrappers.py:208-232: pad = ps - h%ps, >= 1);
    always_pad=False -> SRImplicitDownsampledFastPatchTest semantics (wrappers.py:587-594: pad 0 when divisible).
    Returns inp (3,h,w) in [-1,1], coord (q,q,2), cell (2,), gt_lr_up (27,q,q), (H,W).
    """
    h, w = lr01.shape[-2:]
    H, W = round(h * scale), round(w * scale)
    hr_coord = make_coord([H, W])
    lr_up = F.interpolate(((lr01 - 0.5) / 0.5).unsqueeze(0), (H, W), mode="bilinear", align_corners=False).squeeze(0)
    lr_up_down = F.interpolate(lr_up.unsqueeze(0), (h, w), mode="bilinear", align_corners=False).squeeze(0)
    resid = lr_up - F.interpolate(lr_up_down.unsqueeze(0), (H, W), mode="bilinear", align_corners=False).squeeze(0)
    ps = patch_size
    if always_pad:
        pad_h, pad_w = ps - H % ps, ps - W % ps
    else:
        pad_h, pad_w = (ps - H % ps) % ps, (ps - W % ps) % ps
    coord_pad = F.pad(hr_coord.permute(2, 0, 1), (0, pad_w, 0, pad_h), "constant", 0)
    cu = coord_pad.unfold(1, ps, ps).unfold(2, ps, ps)
    coord = cu[:, :, :, ps // 2, ps // 2].permute(1, 2, 0).contiguous()
    lp = F.pad(resid, (0, pad_w, 0, pad_h), "constant", 0).unfold(1, ps, ps).unfold(2, ps, ps)
    c, a, b, _, _ = lp.shape
    gt_lr_up = lp.contiguous().view(c, a, b, ps * ps).permute(0, 3, 1, 2).contiguous().view(c * ps * ps, a, b)
    cell = torch.tensor([2 / H, 2 / W], dtype=torch.float32)
    inp = (lr01 - 0.5) / 0.5     # test.py:98 with data_norm inp sub 0.5 div 0.5
    return inp, coord, cell, gt_lr_up, (H, W)


# ----------------------------------------------------------------- encoders
def edsr_forward(sd, x, p="encoder.", n_resblocks=16):
    """EDSR.forward, no_upsampling (edsr.py:134-146); baseline: 16 ResBlocks, 64 feats, res_scale 1."""
    def conv(n, t):
        return F.conv2d(t, sd[p + n + ".weight"], sd[p + n + ".bias"], padding=1)
    x = conv("head.0", x)
    res = x
    for i in range(n_resblocks):
        r = conv(f"body.{i}.body.2", F.relu(conv(f"body.{i}.body.0", res)))
        res = r * 1 + res
    res = conv(f"body.{n_resblocks}", res)
    return res + x


def rrdb_forward(sd, x, p="encoder.", nb=23):
    """RRDBNet.forward, no_upsampling (rrdb.py:105-116): fea = conv_first(x) + trunk_conv(trunk(conv_first(x)))."""
    def conv(n, t):
        return F.conv2d(t, sd[p + n + ".weight"], sd[p + n + ".bias"], padding=1)
    fea = conv("conv_first", x)
    t = fea
    for i in range(nb):
        x0 = t
        out = x0
        for r in (1, 2, 3):
            q = f"RRDB_trunk.{i}.RDB{r}"
            xin = out
            x1 = F.leaky_relu(conv(q + ".conv1", xin), 0.2)
            x2 = F.leaky_relu(conv(q + ".conv2", torch.cat((xin, x1), 1)), 0.2)
            x3 = F.leaky_relu(conv(q + ".conv3", torch.cat((xin, x1, x2), 1)), 0.2)
            x4 = F.leaky_relu(conv(q + ".conv4", torch.cat((xin, x1, x2, x3), 1)), 0.2)
            x5 = conv(q + ".conv5", torch.cat((xin, x1, x2, x3, x4), 1))
            out = x5 * 0.2 + xin
        t = out * 0.2 + x0
    return fea + conv("trunk_conv", t)


def gen_feat(sd, enc_name, inp):
    return edsr_forward(sd, inp) if enc_name == "edsr-baseline" else rrdb_forward(sd, inp)


# ----------------------------------------------------------------- query: local Fourier features + MLP
def affine_info(sd, feat, coord, cell):
    """LINFPatch.query_* up to `affine_info = self.layers(features)` (linf.py:248-312 == :324-391)."""
    coef = F.conv2d(feat, sd["coef.weight"], sd["coef.bias"], padding=1)
    freq = F.conv2d(feat, sd["freq.weight"], sd["freq.bias"], padding=1)
    B, _, h, w = feat.shape
    rx, ry = 2 / h / 2, 2 / w / 2
    feat_coord = make_coord((h, w)).permute(2, 0, 1).unsqueeze(0).expand(B, 2, h, w)
    freqs, coefs, areas = [], [], []
    for vx in (-1, 1):
        for vy in (-1, 1):
            coord_ = coord.clone()
            coord_[:, :, :, 0] += vx * rx + 1e-6
            coord_[:, :, :, 1] += vy * ry + 1e-6
            coord_.clamp_(-1 + 1e-6, 1 - 1e-6)
            q_coord = F.grid_sample(feat_coord, coord_.flip(-1), mode="nearest", align_corners=False)
            rel = coord.permute(0, 3, 1, 2) - q_coord
            rel[:, 0] *= h
            rel[:, 1] *= w
            rel_cell = cell.clone()
            rel_cell[:, 0] *= h
            rel_cell[:, 1] *= w
            coef_ = F.grid_sample(coef, coord_.flip(-1), mode="nearest", align_corners=False)
            freq_ = F.grid_sample(freq, coord_.flip(-1), mode="nearest", align_corners=False)
            freq_ = torch.stack(torch.split(freq_, freq.shape[1] // 2, dim=1), dim=2)
            freq_ = torch.sum(freq_ * rel.unsqueeze(1), dim=2)
            freq_ = freq_ + F.linear(rel_cell, sd["phase.weight"]).unsqueeze(-1).unsqueeze(-1)
            freq_ = torch.cat((torch.cos(np.pi * freq_), torch.sin(np.pi * freq_)), dim=1)
            freqs.append(freq_)
            coefs.append(coef_)
            areas.append(torch.abs(rel[:, 0] * rel[:, 1]) + 1e-9)
    tot = torch.stack(areas).sum(dim=0)
    areas[0], areas[3] = areas[3], areas[0]
    areas[1], areas[2] = areas[2], areas[1]
    feats = [((areas[i] / tot).unsqueeze(1) * coefs[i]) * freqs[i] for i in range(4)]
    x = torch.cat(feats, dim=1)
    for i in (0, 2, 4):
        x = F.relu(F.conv2d(x, sd[f"layers.{i}.weight"], sd[f"layers.{i}.bias"]))
    return F.conv2d(x, sd["layers.6.weight"], sd["layers.6.bias"])


# ----------------------------------------------------------------- flow (models/flow.py)
def flow_forward(sd, x, aff, n_layers=10, D=27):
    """Flow.forward (flow.py:44-55) without the log-prob bookkeeping.  x, aff: (N, D), (N, 2*D*n_layers)."""
    z = x
    for i in range(n_layers):
        z = F.linear(z, sd[f"imnet.linears.{i}._weight"], sd[f"imnet.linears.{i}.bias"])
        a = aff[:, i * 2 * D:(i + 1) * 2 * D]
        scale = torch.sigmoid(a[:, :D] + 2.0) + 1e-4
        z = z * scale + a[:, D:]
    return F.linear(z, sd["imnet.last._weight"], sd["imnet.last.bias"])


def flow_inverse(sd, z, aff, n_layers=10, D=27):
    """Flow.inverse (flow.py:57-63); NaiveLinear.inverse solves W x = y - b (flow.py:110-122)."""
    def lin_inv(p, y):
        return torch.linalg.solve(sd[p + "._weight"], (y - sd[p + ".bias"]).t()).t()
    x = lin_inv("imnet.last", z)
    for i in reversed(range(n_layers)):
        a = aff[:, i * 2 * D:(i + 1) * 2 * D]
        scale = torch.sigmoid(a[:, :D] + 2.0) + 1e-4
        x = (x - a[:, D:]) / scale
        x = lin_inv(f"imnet.linears.{i}", x)
    return x


def query_log_p(sd, feat, coord, cell, gt):
    """LINFPatch.query_log_p -> z (B,27,q,q)   (linf.py:248-322)."""
    aff = affine_info(sd, feat, coord, cell)
    B, qh, qw, _ = coord.shape
    z = flow_forward(sd, gt.permute(0, 2, 3, 1).reshape(B * qh * qw, -1), aff.permute(0, 2, 3, 1).reshape(B * qh * qw, -1))
    return z.reshape(B, qh, qw, -1).permute(0, 3, 1, 2)


def query_rgb(sd, feat, coord, cell, zmap, ps=3):
    """LINFPatch.query_rgb with zmap -> (B,3,3q,3q)   (linf.py:324-407; F.fold(k=s=3) == pixel_shuffle(3))."""
    aff = affine_info(sd, feat, coord, cell)
    B, qh, qw, _ = coord.shape
    pred = flow_inverse(sd, zmap.permute(0, 2, 3, 1).reshape(-1, 3 * ps * ps), aff.permute(0, 2, 3, 1).reshape(B * qh * qw, -1))
    pred = pred.view(B, qh, qw, -1).permute(0, 3, 1, 2).contiguous()
    return F.fold(pred.view(B, ps * ps * 3, -1), output_size=(qh * ps, qw * ps), kernel_size=(ps, ps), stride=ps)


# ----------------------------------------------------------------- LINF-LP prior (models/unet.py)
def unet_linf_forward(sd, x, lr, depth=3):
    """LINF-LP UNet.forward(x, lr) (unet.py:144-167)."""
    from .srflow_oracle import _dense5, _double_conv, _up
    x = _dense5(sd, "input_proj", x)
    e = F.conv2d(lr, sd["lr_proj.0.weight"], sd["lr_proj.0.bias"], stride=3, padding=1)
    e = _dense5(sd, "lr_proj.2", F.leaky_relu(e, 0.2))
    if e.shape != x.shape:
        e = F.interpolate(e, size=x.shape[2:], mode="bilinear", align_corners=False)
    z = _double_conv(sd, "inc", torch.cat([x, e], 1))
    feats = [z]
    for i in range(depth):
        z = _double_conv(sd, f"down_layers.{i}.maxpool_conv.1", F.max_pool2d(z, 2))
        feats.append(z)
    for i in range(depth):
        z = _up(sd, f"up_layers.{i}", z, feats[depth - 1 - i])
    return F.conv2d(z, sd["outc.conv.weight"], sd["outc.conv.bias"])


# ----------------------------------------------------------------- whole path (test.py:143-171, eval_bsize set, --patch)
@torch.no_grad()
def lp_sr(sd, prior_sd, enc_name, inp, coord, cell, gt_lr_up, out_hw, literal=True, return_all=False):
    feat = gen_feat(sd, enc_name, inp)
    z_lr = query_log_p(sd, feat, coord, cell, gt_lr_up)                 # batched_predict_log_p (test.py:36-47)
    z_learned = unet_linf_forward(prior_sd, z_lr.contiguous(), inp)     # test.py:147
    if z_learned.shape != z_lr.shape:
        z_learned = F.interpolate(z_learned, size=z_lr.shape[-2:], mode="bilinear", align_corners=False)
    if literal:
        feat = gen_feat(sd, enc_name, inp)                              # batched_predict re-runs the encoder (test.py:22)
    pred = query_rgb(sd, feat, coord, cell, z_learned)
    pred = pred[..., :out_hw[0], :out_hw[1]]                            # test.py:168
    pred = pred + F.interpolate(inp, pred.shape[-2:], mode="bilinear", align_corners=False)   # test.py:171
    if return_all:
        return pred, z_lr, z_learned
    return pred
