"""Learned-prior latent modules ('unet' in both reference registries) on the B200 engine.

* SRFlow-LP: `make_unet(depth, dim=64, bilinear=True)` -> UNet.forward(epses) -> [z0, z1]
  (SRFlow-LP/code/models/unet.py:109-186);
* LINF-LP:   `make_unet(in_chans, depth, dim=64, bilinear=True, cell_input=False)` -> UNet.forward(x, lr)
  (LINF-LP/models/unet.py:105-172).

State-dict keys are the reference's, so the shipped `*-LP.pth` specs load through `models.make(spec, load_sd=True)`.
"""
from __future__ import annotations

import ctypes as C
from collections import OrderedDict

import torch
from torch import nn

from .. import _lib, param_tree
from .models import register


def _dense_shapes(s, p, nf, gc, out):
    for c in range(1, 6):
        s[f"{p}.conv{c}.weight"] = (gc if c < 5 else out, nf + (c - 1) * gc, 3, 3)
        s[f"{p}.conv{c}.bias"] = (gc if c < 5 else out,)


def _dconv_shapes(s, bufs, p, cin, cout, mid=None):
    mid = mid or cout
    for j, (a, b) in zip((0, 3), ((cin, mid), (mid, cout))):
        s[f"{p}.double_conv.{j}.weight"] = (b, a, 3, 3)
        s[f"{p}.double_conv.{j + 1}.weight"] = (b,)
        s[f"{p}.double_conv.{j + 1}.bias"] = (b,)
        for n, shp in (("running_mean", (b,)), ("running_var", (b,)), ("num_batches_tracked", ())):
            s[f"{p}.double_conv.{j + 1}.{n}"] = shp
            bufs.append(f"{p}.double_conv.{j + 1}.{n}")


def _body_shapes(s, bufs, depth, dim, sfx):
    for i in range(depth):
        cout = dim * 2 ** (i + 1) // (2 if i == depth - 1 else 1)
        _dconv_shapes(s, bufs, f"down_layers{sfx}.{i}.maxpool_conv.1", dim * 2 ** i, cout)
    for i in range(depth):
        cin = dim * 2 ** (depth - i)
        cout = dim * 2 ** (depth - i - 1) // (2 if i < depth - 1 else 1)
        _dconv_shapes(s, bufs, f"up_layers{sfx}.{i}.conv", cin, cout, cin // 2)


class _PriorBase(nn.Module):
    # conv arithmetic of the STANDALONE prior calls (0 = split-bf16 x3, 1 = bf16 single pass, 2 = fp32 CUDA cores); inside the
    # fused lp_sr entry points the prior runs under the generator's precision.  Set it to the generator's value when mixing
    # step-by-step and fused calls on one model.
    precision = 0

    def __init__(self):
        super().__init__()
        self._handles = {}

    def load_state_dict(self, state_dict, strict=True, **kw):
        r = super().load_state_dict(state_dict, strict=strict, **kw)
        self.refresh()
        return r

    def refresh(self):
        for h in self._handles.values():
            _lib.lib().bfsr_unet_destroy(h)
        self._handles = {}

    def __del__(self):
        try:
            self.refresh()
        except Exception:
            pass

    def cuda(self, device=None):
        """`.cuda()` (SRFlow-LP/code/test.py:91): weights are packed on the device lazily; parameters stay on the host."""
        return self

    def _desc(self):
        raise NotImplementedError

    def handle(self, device):
        dev_i = _lib.cuda_device(device).index
        key = (dev_i, int(self.precision))
        if key not in self._handles:
            if not torch.cuda.is_available():
                raise _lib.BfsrError("bfsr_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
            table, keep = _lib.tensor_table(self.state_dict())
            h = C.c_void_p()
            d = self._desc()
            d.precision = int(self.precision)
            _lib.check(_lib.lib().bfsr_unet_create(C.byref(h), C.byref(d), table, len(table), dev_i))
            del keep
            self._handles[key] = h
        return self._handles[key]


class SRFlowPriorEngine(_PriorBase):
    """SRFlow-LP UNet (two independent branches, one per latent tensor)."""

    def __init__(self, depth=3, dim=64, bilinear=True, latent_ch=(6, 96)):
        super().__init__()
        assert bilinear, "only bilinear=True priors are shipped/supported"
        self.depth, self.dim, self.bilinear, self.latent_ch = depth, dim, bilinear, tuple(latent_ch)
        s, bufs = OrderedDict(), []
        for b, nf in enumerate(self.latent_ch):
            _dense_shapes(s, f"input_proj{b}", nf, dim, dim)
        for b in range(len(self.latent_ch)):
            _body_shapes(s, bufs, depth, dim, str(b))
        for b in range(len(self.latent_ch)):
            _dconv_shapes(s, bufs, f"inc{b}", dim, dim)
        for b, nf in enumerate(self.latent_ch):
            s[f"outc{b}.conv.weight"] = (nf, dim, 1, 1)
            s[f"outc{b}.conv.bias"] = (nf,)
        param_tree.build(self, s, bufs)
        # BatchNorm defaults so a freshly made prior is usable (gamma=1, var=1)
        for k, v in self.state_dict().items():
            if k.endswith("running_var") or (k.endswith(".weight") and v.dim() == 1):
                v.fill_(1.0)

    def _desc(self):
        d = _lib.UNetDesc()
        d.variant, d.depth, d.dim, d.bilinear = 0, self.depth, self.dim, int(self.bilinear)
        d.n_latents = len(self.latent_ch)
        for i, c in enumerate(self.latent_ch):
            d.latent_ch[i] = c
        return d

    def forward(self, epses):
        """UNet.forward(epses) -> [z0, z1] (unet.py:154-181)."""
        assert len(epses) == len(self.latent_ch)
        dev = epses[0].device if epses[0].is_cuda else torch.device("cuda", torch.cuda.current_device())
        xs = [e.detach().to(dev, torch.float32).contiguous() for e in epses]
        B = xs[0].shape[0]
        for x, c in zip(xs, self.latent_ch):
            assert x.shape[0] == B and x.shape[1] == c, (tuple(x.shape), c)
        outs = [torch.empty_like(x) for x in xs]
        H = (C.c_int32 * len(xs))(*[x.shape[2] for x in xs])
        W = (C.c_int32 * len(xs))(*[x.shape[3] for x in xs])
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().bfsr_unet_forward_srflow(self.handle(dev), _lib.ptr_array(xs), H, W, B,
                                                          _lib.ptr_array(outs), _lib.stream_ptr(dev)))
        return outs


@register('unet')
def make_unet(depth=3, dim=64, bilinear=True, in_chans=None, cell_input=False, latent_ch=(6, 96)):
    """One name, two signatures: `in_chans` selects the LINF-LP variant."""
    if in_chans is None:
        return SRFlowPriorEngine(depth=depth, dim=dim, bilinear=bilinear, latent_ch=latent_ch)
    from .linf import LINFPriorEngine  # noqa: WPS433 (optional family)
    return LINFPriorEngine(in_chans=in_chans, depth=depth, dim=dim, bilinear=bilinear, cell_input=cell_input)
