"""SRFlow generator engine — drop-in for `SRFlowNet` behind `SRFlowModel.get_sr()/get_encode_z()`.

Mirrors the call surface of SRFlow-LP/code/models/modules/SRFlowNet_arch.py:30-82 (constructor
arguments, `forward(gt, lr, z, eps_std, reverse, epses, ...)`, `.flowUpsamplerNet.C/.scaleH/.scaleW`,
`.RRDB_training`, `.set_rrdb_training`) and holds parameters under exactly the reference's state_dict
keys, so `SRFlow_DF2K_4X.pth` / `RRDB_DF2K_4X.pth` load with `strict=True` (base_model.py:112-124).
All arithmetic runs in libbfsr_b200.so (hand-written sm_100a CUDA) through the C ABI; there is no
PyTorch compute path and no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from collections import OrderedDict

import torch
from torch import nn

from .. import _lib, param_tree


def _opt_get(opt, keys, default=None):
    """utils/util.py:167-175."""
    if opt is None:
        return default
    ret = opt
    for k in keys:
        ret = ret.get(k, None) if hasattr(ret, "get") else None
        if ret is None:
            return default
    return ret


def srflow_param_shapes(nf, nb, gc, scale, K, L, n_no_affine, blocks, hidden, split):
    """State-dict layout of SRFlowNet (RRDBNet_arch.py:64-87, FlowUpsamplerNet.py:94-187, FlowStep.py:49-79,
    FlowAffineCouplingsAblation.py:25-55, flow.py:26-83, Split.py:26-37)."""
    s = OrderedDict()
    s["RRDB.conv_first.weight"] = (nf, 3, 3, 3)
    s["RRDB.conv_first.bias"] = (nf,)
    for i in range(nb):
        for r in (1, 2, 3):
            for c in range(1, 6):
                p = f"RRDB.RRDB_trunk.{i}.RDB{r}.conv{c}"
                s[p + ".weight"] = (gc if c < 5 else nf, nf + (c - 1) * gc, 3, 3)
                s[p + ".bias"] = (gc if c < 5 else nf,)
    heads = ["trunk_conv", "upconv1", "upconv2"] + (["upconv3"] if scale >= 8 else []) + ["HRconv"]
    for n in heads:
        s[f"RRDB.{n}.weight"] = (nf, nf, 3, 3)
        s[f"RRDB.{n}.bias"] = (nf,)
    s["RRDB.conv_last.weight"] = (3, nf, 3, 3)
    s["RRDB.conv_last.bias"] = (3,)
    n_cond = (len(blocks) + 1) * nf
    C, idx = 3, 0
    latents = []
    for level in range(1, L + 1):
        C *= 4
        idx += 1  # SqueezeLayer has no parameters
        for k in range(n_no_affine + K):
            p = f"flowUpsamplerNet.layers.{idx}"
            s[p + ".actnorm.bias"] = (1, C, 1, 1)
            s[p + ".actnorm.logs"] = (1, C, 1, 1)
            s[p + ".invconv.weight"] = (C, C)
            if k >= n_no_affine:
                for name, cin, cout in (("fAffine", C // 2 + n_cond, (C - C // 2) * 2), ("fFeatures", n_cond, C * 2)):
                    q = f"{p}.affine.{name}"
                    s[q + ".0.weight"] = (hidden, cin, 3, 3)
                    s[q + ".0.actnorm.bias"] = (1, hidden, 1, 1)
                    s[q + ".0.actnorm.logs"] = (1, hidden, 1, 1)
                    s[q + ".2.weight"] = (hidden, hidden, 1, 1)
                    s[q + ".2.actnorm.bias"] = (1, hidden, 1, 1)
                    s[q + ".2.actnorm.logs"] = (1, hidden, 1, 1)
                    s[q + ".4.weight"] = (cout, hidden, 3, 3)
                    s[q + ".4.bias"] = (cout,)
                    s[q + ".4.logs"] = (cout, 1, 1)
            idx += 1
        if split and level < L - 1:
            cons = int(round(C * 0.5))
            p = f"flowUpsamplerNet.layers.{idx}.conv"
            s[p + ".weight"] = (cons * 2, C - cons, 3, 3)
            s[p + ".bias"] = (cons * 2,)
            s[p + ".logs"] = (cons * 2, 1, 1)
            latents.append((cons, level))
            C -= cons
            idx += 1
    latents.append((C, L))
    fo = 2 * 3 * 64 // 2 // 2 if split else 2 * 3 * 64
    s["flowUpsamplerNet.f.0.weight"] = (fo, n_cond, 3, 3)   # unused by forward, present in every checkpoint
    s["flowUpsamplerNet.f.0.bias"] = (fo,)
    return s, C, latents


class SRFlowNetEngine(nn.Module):
    """`SRFlowNet(in_nc, out_nc, nf, nb, gc=32, scale=4, K=None, opt=None, step=None)` on the B200 engine."""

    def __init__(self, in_nc=3, out_nc=3, nf=64, nb=23, gc=32, scale=4, K=None, opt=None, step=None,
                 device=None, tile_chunk=0, precision=0):
        super().__init__()
        assert in_nc == 3 and out_nc == 3, "SRFlow is RGB -> RGB"
        self.opt = opt
        self.quant = _opt_get(opt, ["datasets", "train", "quant"]) or 255
        flow = _opt_get(opt, ["network_G", "flow"]) or {}
        self.scale = int(_opt_get(opt, ["scale"], scale))
        self.K = int(K if K is not None else flow.get("K", 16))
        self.L = int(flow.get("L", 3) or 3)
        self.n_no_affine = int(flow.get("additionalFlowNoAffine", 0) or 0)
        coupling = flow.get("coupling", "CondAffineSeparatedAndCond")
        if coupling != "CondAffineSeparatedAndCond":
            raise NotImplementedError(f"flow coupling {coupling!r} (the shipped yml uses CondAffineSeparatedAndCond)")
        self.blocks = list(_opt_get(opt, ["network_G", "flow", "stackRRDB", "blocks"]) or [])
        if not _opt_get(opt, ["network_G", "flow", "stackRRDB", "concat"], False):
            raise NotImplementedError("stackRRDB.concat must be true (shipped yml)")
        self.split = bool(_opt_get(opt, ["network_G", "flow", "split", "enable"], False))
        self.hidden = int(_opt_get(opt, ["network_G", "flow", "hidden_channels"]) or 64)
        self.nf, self.nb, self.gc = nf, nb, gc
        shapes, c_final, latents = srflow_param_shapes(nf, nb, gc, self.scale, self.K, self.L, self.n_no_affine,
                                                       self.blocks, self.hidden, self.split)
        param_tree.build(self, shapes)
        self.latent_specs = latents                      # [(channels, level)] in encode order
        # attributes read by SRFlowModel.get_z (SRFlow_model.py:224-237)
        fu = self._modules["flowUpsamplerNet"]
        fu.C = c_final
        fu.H = fu.W = 160 // (2 ** self.L)
        fu.scaleH = fu.scaleW = 160 / fu.H
        self.RRDB_training = True
        self._device = torch.device(device) if device is not None else None
        self.tile_chunk = int(tile_chunk)
        self.precision = int(precision)
        self._handle = None
        self._handle_key = None

    # ---- reference surface -------------------------------------------------------------
    @property
    def module(self):
        """`netG.module` (DataParallel unwrap in SRFlowModel, SRFlow_model.py:227) resolves to the engine itself."""
        return self

    def set_rrdb_training(self, trainable):
        """SRFlowNet.set_rrdb_training (SRFlowNet_arch.py:52-58); inference engine: only tracks the flag."""
        if self.RRDB_training != trainable:
            self.RRDB_training = trainable
            return True
        return False

    def load_state_dict(self, state_dict, strict=True, **kw):
        r = super().load_state_dict(state_dict, strict=strict, **kw)
        self.refresh()
        return r

    def refresh(self):
        """Drop the packed device weights; they are rebuilt from the current parameters on the next call."""
        self._destroy()

    # ---- engine plumbing ---------------------------------------------------------------
    def device(self):
        if self._device is None:
            p = next(self.parameters())
            self._device = p.device if p.is_cuda else None
        # stored with an explicit index: the handle, the buffers and the stream must name the same GPU
        self._device = _lib.cuda_device(self._device)
        return self._device

    def _destroy(self):
        if self._handle is not None:
            _lib.lib().bfsr_srflow_destroy(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self._destroy()
        except Exception:
            pass

    def handle(self):
        if self._handle is None:
            if not torch.cuda.is_available():
                raise _lib.BfsrError("bfsr_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
            d = _lib.SRFlowDesc()
            d.scale, d.nf, d.nb, d.gc = self.scale, self.nf, self.nb, self.gc
            d.K, d.L, d.n_no_affine, d.hidden = self.K, self.L, self.n_no_affine, self.hidden
            d.n_blocks = len(self.blocks)
            for i, b in enumerate(self.blocks):
                d.blocks[i] = int(b)
            d.split_enable = int(self.split)
            d.tile_chunk = self.tile_chunk
            d.precision = self.precision
            table, keep = _lib.tensor_table(self.state_dict())
            h = C.c_void_p()
            dev = self.device()
            _lib.check(_lib.lib().bfsr_srflow_create(C.byref(h), C.byref(d), table, len(table), dev.index))
            del keep
            self._handle = h
        return self._handle

    def _prep(self, t):
        return t.detach().to(self.device(), torch.float32).contiguous()

    def latent_shapes(self, lr_h, lr_w):
        return [(c, (lr_h * self.scale) >> lv, (lr_w * self.scale) >> lv) for c, lv in self.latent_specs]

    # ---- SRFlowNet.forward (SRFlowNet_arch.py:60-82) -------------------------------------
    def forward(self, gt=None, lr=None, z=None, eps_std=None, reverse=False, epses=None, reverse_with_grad=False,
                lr_enc=None, add_gt_noise=False, step=None, y_label=None):
        if lr_enc is not None:
            raise NotImplementedError("lr_enc: the engine caches the encoder internally (use lp_sr for the fused path)")
        if not reverse:
            return self.normal_flow(gt, lr, epses=epses, add_gt_noise=add_gt_noise)
        assert lr.shape[1] == 3
        return self.reverse_flow(lr, z, eps_std=eps_std, epses=epses)

    def normal_flow(self, gt, lr, epses=None, add_gt_noise=True):
        """SRFlowNet.normal_flow (SRFlowNet_arch.py:83-116).  `epses`, when a list, is appended to IN PLACE
        (FlowUpsamplerNet.py:250-251,263-264 — SRFlow-LP/code/test.py:138-141 relies on it).  nll / logdet are
        dead on the inference path (SRFlow_model.py:204) and are returned as NaN placeholders."""
        lib = _lib.lib()
        lr_d, gt_d = self._prep(lr), self._prep(gt)
        B, _, h, w = lr_d.shape
        assert gt_d.shape == (B, 3, h * self.scale, w * self.scale), (gt_d.shape, lr_d.shape)
        if add_gt_noise:  # SRFlowNet_arch.py:92-97
            gt_d = gt_d + (torch.rand(gt_d.shape, device=gt_d.device) - 0.5) / self.quant
        outs = [torch.empty((B, c, H, W), device=lr_d.device, dtype=torch.float32)
                for c, H, W in self.latent_shapes(h, w)]
        with torch.cuda.device(lr_d.device):
            _lib.check(lib.bfsr_srflow_encode(self.handle(), lr_d.data_ptr(), gt_d.data_ptr(), B, h, w,
                                              _lib.ptr_array(outs), _lib.stream_ptr(lr_d.device)))
        nan = torch.full((B,), float("nan"), device=lr_d.device)
        if isinstance(epses, list):
            epses.extend(outs)
            return epses, nan, nan.clone()
        return outs[-1], nan, nan.clone()

    def reverse_flow(self, lr, z, eps_std=None, epses=None):
        """SRFlowNet.reverse_flow (SRFlowNet_arch.py:145-158).  Never mutates `epses` (FlowUpsamplerNet.py:206)."""
        lib = _lib.lib()
        lr_d = self._prep(lr)
        B, _, h, w = lr_d.shape
        shapes = self.latent_shapes(h, w)
        if isinstance(epses, (list, tuple)):
            lat = [self._prep(e) for e in epses]
        else:
            # z-only call (get_sr(lq, heat=tau)): Split2d draws its eps ~ N(0, eps_std) (Split.py:66-68)
            assert z is not None, "reverse flow needs z or epses"
            lat = []
            for c, H, W in shapes[:-1]:
                e = torch.zeros((B, c, H, W), device=lr_d.device)
                if eps_std:
                    e.normal_(0.0, float(eps_std))
                lat.append(e)
            lat.append(self._prep(z))
        assert len(lat) == len(shapes), f"{len(lat)} latents given, topology has {len(shapes)}"
        for t, (c, H, W) in zip(lat, shapes):
            assert tuple(t.shape) == (B, c, H, W), (tuple(t.shape), (B, c, H, W))
        sr = torch.empty((B, 3, h * self.scale, w * self.scale), device=lr_d.device, dtype=torch.float32)
        with torch.cuda.device(lr_d.device):
            _lib.check(lib.bfsr_srflow_decode(self.handle(), lr_d.data_ptr(), _lib.ptr_array(lat), B, h, w,
                                              sr.data_ptr(), _lib.stream_ptr(lr_d.device)))
        return sr, torch.full((B,), float("nan"), device=lr_d.device)

    # ---- fused LP path (SRFlow-LP/code/test.py:135-148 in one call) ----------------------
    def lp_sr(self, lr, prior, out=None):
        """bilinear(lr) -> encode -> normalise -> prior -> decode, encoder and feature-only convs run once.  `out`: optional
        preallocated (B,3,s*h,s*w) fp32 CUDA tensor for the result (a stable output buffer lets the engine replay one CUDA graph)."""
        lr_d = self._prep(lr)
        B, _, h, w = lr_d.shape
        if out is not None:
            assert out.is_cuda and out.dtype == torch.float32 and out.is_contiguous() and \
                tuple(out.shape) == (B, 3, h * self.scale, w * self.scale) and out.device == lr_d.device
        sr = out if out is not None else torch.empty((B, 3, h * self.scale, w * self.scale), device=lr_d.device, dtype=torch.float32)
        with torch.cuda.device(lr_d.device):
            _lib.check(_lib.lib().bfsr_srflow_lp_sr(self.handle(), prior.handle(self.device()), lr_d.data_ptr(), B, h, w,
                                                    sr.data_ptr(), _lib.stream_ptr(lr_d.device)))
        return sr

    def sr_image(self, lr_img, prior, pad_factor=2):
        """Image-level loop body of SRFlow-LP/code/test.py:121-151 for one HWC uint8 RGB image (numpy): reflect-pad bottom / right
        to a multiple of `pad_factor` (impad, :80-81,126-130), `t()` (:57), the LP path, clamp, `rgb()` (:59-61: x255 and a
        truncating uint8 cast) and the crop back to (h*scale, w*scale) (:151).  Returns an HWC uint8 numpy array."""
        import numpy as np
        assert lr_img.ndim == 3 and lr_img.shape[2] == 3 and lr_img.dtype == np.uint8
        h, w, _ = lr_img.shape
        pb, pr = int(np.ceil(h / pad_factor) * pad_factor - h), int(np.ceil(w / pad_factor) * pad_factor - w)
        lr = np.pad(lr_img, [(0, pb), (0, pr), (0, 0)], "reflect")
        lr_t = torch.from_numpy(np.ascontiguousarray(lr.transpose(2, 0, 1))[None].astype(np.float32)) / 255
        sr_t = self.lp_sr(lr_t, prior)
        sr = (torch.clamp(sr_t[0], 0, 1) * 255).to(torch.uint8).permute(1, 2, 0).cpu().numpy()
        return sr[:h * self.scale, :w * self.scale]

    def lp_sr_host(self, lr_host, prior, out=None):
        """Same through the host-buffer entry point: lr_host is a CPU tensor (pinned for speed), result on the CPU."""
        assert not lr_host.is_cuda and lr_host.dtype == torch.float32 and lr_host.is_contiguous()
        B, _, h, w = lr_host.shape
        if out is None:
            out = torch.empty((B, 3, h * self.scale, w * self.scale), dtype=torch.float32, pin_memory=True)
        dev = self.device()
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().bfsr_srflow_lp_sr_host(self.handle(), prior.handle(dev), lr_host.data_ptr(), B, h, w,
                                                         out.data_ptr(), _lib.stream_ptr(dev)))
        return out


class SRFlowModel:
    """The inference surface of the reference's `SRFlowModel` (SRFlow-LP/code/models/SRFlow_model.py:198-237) over the engine:
    `get_sr`, `get_sr_with_z`, `get_encode_z`, `get_z` with the reference's argument meaning, including its quirks -- the net is
    left in train() mode after every call (:205,221) and `get_z` reads `netG.module.flowUpsamplerNet.{C,scaleH,scaleW}`
    (:227-229).  `SRFlow-LP/code/test.py:135-148` runs unchanged on top of it (`model.get_encode_z(...)`, `model.get_sr(...)`).
    Host glue only: every tensor op is a call into the engine."""

    def __init__(self, opt, step=0, netG=None, **kw):
        self.opt = opt
        self.netG = netG if netG is not None else define_Flow(opt, step, **kw)

    def load_network(self, load_path_or_sd, network=None, strict=True, submodule=None):
        """base_model.load_network (base_model.py:112-124): flat state_dict, `module.` prefixes stripped, optional submodule."""
        sd = torch.load(load_path_or_sd, map_location="cpu") if isinstance(load_path_or_sd, str) else load_path_or_sd
        sd = OrderedDict((k[7:] if k.startswith("module.") else k, v) for k, v in sd.items())
        net = self.netG if network is None else network
        if submodule is not None:
            full = net.state_dict()
            full.update({f"{submodule}.{k}": v for k, v in sd.items()})
            sd = full
        net.load_state_dict(sd, strict=strict)

    def get_sr(self, lq, heat=None, seed=None, z=None, epses=None):
        return self.get_sr_with_z(lq, heat, seed, z, epses)[0]

    def get_encode_z(self, lq, gt, epses=None, add_gt_noise=True):
        self.netG.eval()
        with torch.no_grad():
            z, _, _ = self.netG(gt=gt, lr=lq, reverse=False, epses=epses, add_gt_noise=add_gt_noise)
        self.netG.train()
        return z

    def get_sr_with_z(self, lq, heat=None, seed=None, z=None, epses=None):
        self.netG.eval()
        z = self.get_z(heat, seed, batch_size=lq.shape[0], lr_shape=lq.shape) if z is None and epses is None else z
        sr, logdet = self.netG(lr=lq, z=z, eps_std=heat, reverse=True, epses=epses, reverse_with_grad=True)
        self.netG.train()
        return sr, z

    def get_z(self, heat, seed=None, batch_size=1, lr_shape=None):
        if seed:
            torch.manual_seed(seed)
        fu = self.netG.module.flowUpsamplerNet
        H = int(self.opt["scale"] * lr_shape[2] // fu.scaleH)
        W = int(self.opt["scale"] * lr_shape[3] // fu.scaleW)
        if heat and heat > 0:
            return torch.normal(mean=0, std=heat, size=(batch_size, fu.C, H, W))
        return torch.zeros((batch_size, fu.C, H, W))


def define_Flow(opt, step=0, **kw):
    """networks.define_Flow (SRFlow-LP/code/models/networks.py:70-79) returning the engine."""
    n = opt["network_G"]
    return SRFlowNetEngine(in_nc=n["in_nc"], out_nc=n["out_nc"], nf=n["nf"], nb=n["nb"], scale=opt["scale"],
                           K=n["flow"]["K"], opt=opt, step=step, **kw)
