"""LINF engine ('linf-patch') and LINF-LP prior on the B200 engine.

Mirrors LINF-LP/models/linf.py:218-428: `model(op, inp=, feat=, coord=, cell=, gt=, temperature=, zmap=)` with
`op in {"gen_feat", "query_log_p", "query_rgb", "log_p", "rgb"}`, attribute `patch_size`, and the state_dict keys of the
shipped `edsr-baseline-linf.pth` / `rrdb-linf.pth`, so `models.make(torch.load(path)['model'], load_sd=True)` works
unchanged (LINF-LP/test.py:276-281).  `batched_predict` / `batched_predict_log_p` restate LINF-LP/test.py:20-47 on top of
that surface; `lp_sr` is the fused path of test.py:143-171.  All arithmetic runs in libbfsr_b200.so.
"""
from __future__ import annotations

import ctypes as C
from collections import OrderedDict

import torch
from torch import nn

from .. import _lib, param_tree
from .models import register
from .unet import _PriorBase, _body_shapes, _dconv_shapes, _dense_shapes


def _encoder_shapes(s, name, args):
    if name == "edsr-baseline":
        if not args.get("no_upsampling", True):
            raise NotImplementedError("edsr-baseline with upsampling tail (LINF uses no_upsampling=True)")
        nb = args.get("n_resblocks", 16)
        for n in ("sub_mean", "add_mean"):      # present in the checkpoint, unused by forward (edsr.py:134-146)
            s[f"encoder.{n}.weight"] = (3, 3, 1, 1)
            s[f"encoder.{n}.bias"] = (3,)
        s["encoder.head.0.weight"] = (64, 3, 3, 3)
        s["encoder.head.0.bias"] = (64,)
        for i in range(nb):
            for j in (0, 2):
                s[f"encoder.body.{i}.body.{j}.weight"] = (64, 64, 3, 3)
                s[f"encoder.body.{i}.body.{j}.bias"] = (64,)
        s[f"encoder.body.{nb}.weight"] = (64, 64, 3, 3)
        s[f"encoder.body.{nb}.bias"] = (64,)
        return 0, nb
    if name == "rrdb":
        nb = args.get("nb", 23)
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
        return 1, nb
    raise NotImplementedError(f"LINF encoder {name!r}: only 'edsr-baseline' and 'rrdb' ship checkpoints (SURVEY.md §2.1)")


@register('linf-patch')
class LINFEngine(nn.Module):
    """`LINFPatch(encoder_spec, imnet_spec, flow_layers=10, num_layer=3, hidden_dim=256, patch_size=3)`."""

    def __init__(self, encoder_spec, imnet_spec=None, flow_layers=10, num_layer=3, hidden_dim=256, patch_size=3,
                 device=None, precision=0, tile_chunk=0):
        super().__init__()
        if num_layer != 3:
            raise NotImplementedError("num_layer != 3 (shipped checkpoints use 3)")
        if patch_size != 3:
            raise NotImplementedError("patch_size != 3 (shipped checkpoints use 3)")
        self.patch_size, self.flow_layers, self.hidden_dim = patch_size, flow_layers, hidden_dim
        self.precision, self.tile_chunk = int(precision), int(tile_chunk)
        D = 3 * patch_size * patch_size
        s = OrderedDict()
        self._enc_kind, self._nb = _encoder_shapes(s, encoder_spec["name"], dict(encoder_spec.get("args") or {}))
        for n in ("coef", "freq"):
            s[f"{n}.weight"] = (hidden_dim, 64, 3, 3)
            s[f"{n}.bias"] = (hidden_dim,)
        s["phase.weight"] = (hidden_dim // 2, 2)
        dims = [hidden_dim * 4, hidden_dim, hidden_dim, hidden_dim, flow_layers * D * 2]
        for i in range(4):
            s[f"layers.{2 * i}.weight"] = (dims[i + 1], dims[i], 1, 1)
            s[f"layers.{2 * i}.bias"] = (dims[i + 1],)
        for i in range(flow_layers):
            s[f"imnet.linears.{i}.bias"] = (D,)
            s[f"imnet.linears.{i}._weight"] = (D, D)
        s["imnet.last.bias"] = (D,)
        s["imnet.last._weight"] = (D, D)
        param_tree.build(self, s)
        self._device = torch.device(device) if device is not None else None
        self._handle = None

    # ---- plumbing ----------------------------------------------------------------------
    def load_state_dict(self, state_dict, strict=True, **kw):
        r = super().load_state_dict(state_dict, strict=strict, **kw)
        self._destroy()
        return r

    def cuda(self, device=None):
        """`.cuda()` (LINF-LP/test.py:278): weights are packed on the device lazily; parameters stay on the host."""
        if device is not None:
            self._device = torch.device("cuda", device) if isinstance(device, int) else torch.device(device)
        return self

    def device(self):
        self._device = _lib.cuda_device(self._device)
        return self._device

    def _destroy(self):
        self.__dict__["_serial"] = self.__dict__.get("_serial", 0) + 1      # cached features / affine parameters belong to the old weights
        self.__dict__.pop("_feat_cache", None); self.__dict__.pop("_aff_cache", None)
        if self._handle is not None:
            _lib.lib().bfsr_linf_destroy(self._handle)
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
            d = _lib.LINFDesc()
            d.encoder, d.nb, d.hidden, d.flow_layers = self._enc_kind, self._nb, self.hidden_dim, self.flow_layers
            d.patch_size, d.tile_chunk, d.precision = self.patch_size, self.tile_chunk, self.precision
            table, keep = _lib.tensor_table(self.state_dict())
            h = C.c_void_p()
            _lib.check(_lib.lib().bfsr_linf_create(C.byref(h), C.byref(d), table, len(table), self.device().index))
            del keep
            self._handle = h
        return self._handle

    def _prep(self, t):
        return t.detach().to(self.device(), torch.float32).contiguous()

    # ---- LINFPatch operators (linf.py:244-428) --------------------------------------------
    # The reference's drivers call gen_feat twice per batch and query_* once per pass and per 256-row chunk, recomputing the
    # encoder, the coef / freq convs and the MLP every time (LINF-LP/test.py:20-47).  Through the same call surface the engine
    # remembers (a) the features of the last inputs and (b) the per-query affine parameters of the last (features, coord chunk)
    # pairs.  A cache entry is tied to the tensor OBJECT (weak reference, so an id is never reused) and its in-place version
    # counter, so modified or new tensors miss.  `cache_entries = 0` disables it.
    cache_entries = 8

    @staticmethod
    def _root(t):
        return t._base if t._base is not None else t     # row-chunk slices are fresh view objects of one live tensor

    @staticmethod
    def _key(t):
        r = t._base if t._base is not None else t
        return (id(r), t._version, t.data_ptr(), tuple(t.shape), tuple(t.stride()))

    def _cache_get(self, store, key_tensors):
        import weakref
        if not self.cache_entries:
            return None, None
        cache = self.__dict__.setdefault(store, OrderedDict())
        key = tuple(self._key(t) for t in key_tensors) + (self._handle_serial(),)
        hit = cache.get(key)
        if hit is not None and all(r() is self._root(t) for r, t in zip(hit[0], key_tensors)):
            cache.move_to_end(key)
            return key, hit[1]
        return key, None

    def _cache_put(self, store, key, key_tensors, value):
        import weakref
        if key is None:
            return
        cache = self.__dict__.setdefault(store, OrderedDict())
        cache[key] = ([weakref.ref(self._root(t)) for t in key_tensors], value)
        while len(cache) > self.cache_entries:
            cache.popitem(last=False)

    def _handle_serial(self):
        return self.__dict__.get("_serial", 0)

    def gen_feat(self, inp):
        key, hit = self._cache_get("_feat_cache", [inp])
        if hit is not None:
            return hit
        x = self._prep(inp)
        B, _, h, w = x.shape
        feat = torch.empty((B, 64, h, w), device=x.device, dtype=torch.float32)
        with torch.cuda.device(x.device):
            _lib.check(_lib.lib().bfsr_linf_gen_feat(self.handle(), x.data_ptr(), B, h, w, feat.data_ptr(),
                                                     _lib.stream_ptr(x.device)))
        self._cache_put("_feat_cache", key, [inp], feat)
        return feat

    def affine_info(self, feat, coord, cell):
        """coef / freq conv + local Fourier features + MLP of a query chunk (linf.py:251-321): (B,qh,qw,540) NHWC, cached."""
        key, hit = self._cache_get("_aff_cache", [feat, coord, cell])
        if hit is not None:
            return hit
        f, c, ce = self._prep(feat), self._prep(coord), self._prep(cell)
        B, _, h, w = f.shape
        _, qh, qw, _ = c.shape
        aff = torch.empty((B, qh, qw, 2 * 3 * self.patch_size ** 2 * self.flow_layers), device=f.device, dtype=torch.float32)
        with torch.cuda.device(f.device):
            _lib.check(_lib.lib().bfsr_linf_affine(self.handle(), f.data_ptr(), B, h, w, c.data_ptr(), ce.data_ptr(), qh, qw,
                                                   aff.data_ptr(), _lib.stream_ptr(f.device)))
        self._cache_put("_aff_cache", key, [feat, coord, cell], aff)
        return aff

    def _query(self, feat, coord, cell, zin, mode):
        aff = self.affine_info(feat, coord, cell)
        zin = self._prep(zin)
        B, qh, qw, _ = aff.shape
        D = 3 * self.patch_size ** 2
        assert tuple(zin.shape) == (B, D, qh, qw), (tuple(zin.shape), (B, D, qh, qw))
        shape = (B, D, qh, qw) if mode == 0 else (B, 3, qh * self.patch_size, qw * self.patch_size)
        out = torch.empty(shape, device=aff.device, dtype=torch.float32)
        with torch.cuda.device(aff.device):
            _lib.check(_lib.lib().bfsr_linf_flow(self.handle(), aff.data_ptr(), zin.data_ptr(), B, qh, qw, mode, out.data_ptr(),
                                                 _lib.stream_ptr(aff.device)))
        return out

    def query_log_p(self, inp, feat, coord, cell, gt):
        """-> (log_p, z); log_p is dead on the inference path (test.py:43 discards it) and returned as NaN."""
        z = self._query(feat, coord, cell, gt, 0)
        return torch.full((z.shape[0] * z.shape[2] * z.shape[3],), float("nan"), device=z.device), z

    def query_rgb(self, inp, feat, coord, cell, temperature=0, zmap=None):
        if zmap is None:   # sampling mode (linf.py:398): z ~ N(0, temperature^2)
            B, qh, qw, _ = coord.shape
            zmap = torch.randn((B, 3 * self.patch_size ** 2, qh, qw), device=self.device()) * temperature
        return self._query(feat, coord, cell, zmap, 1)

    def log_p(self, inp, coord, cell, gt):
        return self.query_log_p(inp, self.gen_feat(inp), coord, cell, gt)

    def rgb(self, inp, coord, cell, temperature=0, zmap=None):
        return self.query_rgb(inp, self.gen_feat(inp), coord, cell, temperature, zmap)

    def forward(self, op, inp=None, feat=None, coord=None, cell=None, gt=None, temperature=0, zmap=None):
        if op == "query_log_p":
            return self.query_log_p(inp, feat, coord, cell, gt)
        if op == "query_rgb":
            return self.query_rgb(inp, feat, coord, cell, temperature, zmap)
        if op == "log_p":
            return self.log_p(inp, coord, cell, gt)
        if op == "rgb":
            return self.rgb(inp, coord, cell, temperature, zmap)
        if op == "gen_feat":
            return self.gen_feat(inp)
        raise ValueError(f"unknown op {op!r}")

    # ---- fused LP path (test.py:143-171) -------------------------------------------------
    def lp_sr(self, inp, coord, cell, gt_lr_up, prior, out_hw):
        inp, coord, cell, gt = self._prep(inp), self._prep(coord), self._prep(cell), self._prep(gt_lr_up)
        B, _, h, w = inp.shape
        _, qh, qw, _ = coord.shape
        pred = torch.empty((B, 3, out_hw[0], out_hw[1]), device=inp.device, dtype=torch.float32)
        with torch.cuda.device(inp.device):
            _lib.check(_lib.lib().bfsr_linf_lp_sr(self.handle(), prior.handle(self.device()), inp.data_ptr(), B, h, w,
                                                  coord.data_ptr(), cell.data_ptr(), gt.data_ptr(), qh, qw, out_hw[0], out_hw[1],
                                                  pred.data_ptr(), _lib.stream_ptr(inp.device)))
        return pred

    def lp_sr_host(self, inp, coord, cell, gt_lr_up, prior, out_hw, out=None):
        for t in (inp, coord, cell, gt_lr_up):
            assert not t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()
        B, _, h, w = inp.shape
        _, qh, qw, _ = coord.shape
        if out is None:
            out = torch.empty((B, 3, out_hw[0], out_hw[1]), dtype=torch.float32, pin_memory=True)
        dev = self.device()
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().bfsr_linf_lp_sr_host(self.handle(), prior.handle(dev), inp.data_ptr(), B, h, w, coord.data_ptr(),
                                                       cell.data_ptr(), gt_lr_up.data_ptr(), qh, qw, out_hw[0], out_hw[1],
                                                       out.data_ptr(), _lib.stream_ptr(dev)))
        return out


class LINFPriorEngine(_PriorBase):
    """LINF-LP UNet(in_chans, depth, dim, bilinear) — LINF-LP/models/unet.py:105-172."""

    def __init__(self, in_chans=27, depth=3, dim=64, bilinear=True, cell_input=False):
        super().__init__()
        assert bilinear, "only bilinear=True priors are shipped/supported"
        self.in_chans, self.depth, self.dim, self.bilinear = in_chans, depth, dim, bilinear
        s, bufs = OrderedDict(), []
        _dense_shapes(s, "input_proj", in_chans, dim // 2, dim // 2)
        s["lr_proj.0.weight"] = (in_chans, 3, 3, 3)
        s["lr_proj.0.bias"] = (in_chans,)
        _dense_shapes(s, "lr_proj.2", in_chans, dim // 2, dim // 2)
        _body_shapes(s, bufs, depth, dim, "")
        _dconv_shapes(s, bufs, "inc", dim, dim)
        s["outc.conv.weight"] = (in_chans, dim, 1, 1)
        s["outc.conv.bias"] = (in_chans,)
        param_tree.build(self, s, bufs)

    def _desc(self):
        d = _lib.UNetDesc()
        d.variant, d.depth, d.dim, d.bilinear, d.in_chans = 1, self.depth, self.dim, int(self.bilinear), self.in_chans
        return d

    def forward(self, x, lr):
        """prior_model(z_lr, inp) (LINF-LP/test.py:147)."""
        dev = x.device if x.is_cuda else torch.device("cuda", torch.cuda.current_device())
        x = x.detach().to(dev, torch.float32).contiguous()
        lr = lr.detach().to(dev, torch.float32).contiguous()
        B, Cc, qh, qw = x.shape
        assert Cc == self.in_chans and lr.shape[0] == B and lr.shape[1] == 3
        out = torch.empty_like(x)
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().bfsr_unet_forward_linf(self.handle(dev), x.data_ptr(), lr.data_ptr(), B, qh, qw,
                                                         lr.shape[2], lr.shape[3], out.data_ptr(), _lib.stream_ptr(dev)))
        return out


# ---- LINF-LP/test.py:20-47 -------------------------------------------------------------------
def batched_predict(model, inp, coord, cell, temperature, zmap=None):
    with torch.no_grad():
        feat = model("gen_feat", inp=inp)
        _, h, w, _ = coord.shape
        row, preds = 0, []
        while row < h:
            z = None if zmap is None else zmap[:, :, row:row + 256, :]
            preds.append(model("query_rgb", inp=inp, feat=feat, coord=coord[:, row:row + 256, :, :], cell=cell,
                               temperature=temperature, zmap=z))
            row += 256
        return torch.cat(preds, dim=2)


def batched_predict_log_p(model, inp, coord, cell, gt):
    with torch.no_grad():
        feat = model("gen_feat", inp=inp)
        _, h, w, _ = coord.shape
        row, preds = 0, []
        while row < h:
            _, z = model("query_log_p", inp=inp, feat=feat, coord=coord[:, row:row + 256, :, :], cell=cell,
                         gt=gt[:, :, row:row + 256, :])
            preds.append(z)
            row += 256
        return torch.cat(preds, dim=2)


# ---- LINF-LP/datasets/wrappers.py:154-238, 516-613 -------------------------------------------
def build_inputs(lr01, scale, patch_size=3, always_pad=True):
    """Test-time inputs of the LINF wrappers for a batch of LR images in [0,1] (B,3,h,w) on a CUDA device:
    returns inp (B,3,h,w) in [-1,1], coord (B,qh,qw,2), cell (B,2), gt_lr_up (B,3*ps*ps,qh,qw), (H,W).
    always_pad=True is the paired test wrapper (one extra patch row/column even when H % ps == 0, wrappers.py:218-219),
    False the arbitrary-scale wrapper (wrappers.py:587-594)."""
    import ctypes as C
    assert lr01.is_cuda and lr01.dtype == torch.float32 and lr01.dim() == 4 and lr01.shape[1] == 3
    lr01 = lr01.contiguous()
    B, _, h, w = lr01.shape
    H, W = round(h * scale), round(w * scale)
    qh, qw = C.c_int32(), C.c_int32()
    L = _lib.lib()
    _lib.check(L.bfsr_linf_build_inputs(None, B, h, w, H, W, patch_size, int(always_pad), None, None, None, None, C.byref(qh), C.byref(qw), None))
    dev = lr01.device
    inp = torch.empty_like(lr01)
    coord = torch.empty((B, qh.value, qw.value, 2), device=dev, dtype=torch.float32)
    cell = torch.empty((B, 2), device=dev, dtype=torch.float32)
    gt = torch.empty((B, 3 * patch_size ** 2, qh.value, qw.value), device=dev, dtype=torch.float32)
    with torch.cuda.device(dev):
        _lib.check(L.bfsr_linf_build_inputs(lr01.data_ptr(), B, h, w, H, W, patch_size, int(always_pad), inp.data_ptr(), coord.data_ptr(),
                                            cell.data_ptr(), gt.data_ptr(), C.byref(qh), C.byref(qw), _lib.stream_ptr(dev)))
    return inp, coord, cell, gt, (H, W)


def lp_sr_mixed(model, prior, lr01, scales, world=1, rank=0, always_pad=False):
    """LP inference of a batch of equally sized LR images (B,3,h,w) in [0,1] with one scale PER IMAGE (BASELINE config 5).
    Images are bucketed by scale, every bucket is sharded over `world` ranks (`dist.bucket_by_scale`), each local bucket runs as
    one `build_inputs` + `lp_sr` call.  Returns {image index: (3, H, W) CUDA tensor} for this rank's images."""
    from ..dist import bucket_by_scale
    assert lr01.dim() == 4 and len(scales) == lr01.shape[0]
    dev = model.device()
    out = {}
    for s, idx in bucket_by_scale(list(scales), world, rank).items():
        sub = lr01[idx].to(dev)
        inp, coord, cell, gt, hw = build_inputs(sub, s, model.patch_size, always_pad)
        pred = model.lp_sr(inp, coord, cell, gt, prior, hw)
        for j, i in enumerate(idx):
            out[i] = pred[j]
    return out



# ---- stand-alone encoders under the reference's registry names (LINF-LP/models/edsr.py:168-181, rrdb.py:119-129) -------------
class EncoderEngine(nn.Module):
    """`models.make({'name': 'edsr-baseline' | 'rrdb', 'args': {..., 'no_upsampling': True}})`: the LR encoder of LINF on its own
    (EDSR.forward edsr.py:134-146 / RRDBNet.forward rrdb.py:105-116, `no_upsampling` variants -- the only ones LINF uses), with the
    encoder's own state_dict keys and `out_dim`.  Runs the `gen_feat` path of the engine (`bfsr_linf_gen_feat`)."""

    def __init__(self, name, args):
        super().__init__()
        s = OrderedDict()
        _encoder_shapes(s, name, dict(args))
        param_tree.build(self, OrderedDict((k[len("encoder."):], v) for k, v in s.items()))
        self._spec = {"name": name, "args": dict(args)}
        self.out_dim = 64
        self._inner = None

    def load_state_dict(self, state_dict, strict=True, **kw):
        r = super().load_state_dict(state_dict, strict=strict, **kw)
        self._inner = None
        return r

    def cuda(self, device=None):
        self._device = device
        return self

    def forward(self, x):
        if self._inner is None:
            inner = LINFEngine(encoder_spec=self._spec, device=getattr(self, "_device", None))
            sd = inner.state_dict()
            for k, v in self.state_dict().items():
                sd["encoder." + k] = v
            eye = torch.eye(3 * inner.patch_size ** 2)
            for k in sd:                      # the query side is unused by gen_feat; its flow matrices only have to be invertible
                if k.endswith("._weight"):
                    sd[k] = eye.clone()
            inner.load_state_dict(sd, strict=True)
            self._inner = inner
        return self._inner.gen_feat(x)


@register('edsr-baseline')
def make_edsr_baseline(n_resblocks=16, n_feats=64, res_scale=1, scale=2, no_upsampling=False, rgb_range=1):
    if n_feats != 64 or res_scale != 1 or not no_upsampling:
        raise NotImplementedError("edsr-baseline: the engine builds the LINF configuration (64 feats, res_scale 1, no_upsampling=True)")
    return EncoderEngine("edsr-baseline", {"n_resblocks": n_resblocks, "no_upsampling": True})


@register('rrdb')
def make_rrdb(in_nc=3, out_nc=3, nf=64, nb=23, gc=32, no_upsampling=True):
    if (in_nc, out_nc, nf, gc) != (3, 3, 64, 32) or not no_upsampling:
        raise NotImplementedError("rrdb: the engine builds the LINF configuration (3->64, gc 32, no_upsampling=True)")
    return EncoderEngine("rrdb", {"nb": nb, "no_upsampling": True})


# ---- stand-alone Flow under the reference's registry name (LINF-LP/models/flow.py:11-63) -------------------------------------
@register('flow')
class FlowEngine(nn.Module):
    """`Flow(flow_layers=10, patch_size=3)`: forward(x, affine_info) -> (z, log_det) and inverse(z, affine_info) -> x on (N, 27)
    vectors with (N, 54*flow_layers) affine parameters; state_dict keys `linears.<i>._weight/.bias`, `last._weight/.bias`.  The
    log-determinant is dead on the inference path and returned as NaN.  Only 3x3 patches (D = 27) are built."""

    def __init__(self, flow_layers=10, patch_size=3, name='flow'):
        super().__init__()
        if patch_size != 3:
            raise NotImplementedError("flow: only patch_size 3 (D = 27) is built (the shipped checkpoints)")
        self.n_layers, D = flow_layers, 3 * patch_size ** 2
        s = OrderedDict()
        for i in range(flow_layers):
            s[f"linears.{i}.bias"] = (D,)
            s[f"linears.{i}._weight"] = (D, D)
        s["last.bias"] = (D,)
        s["last._weight"] = (D, D)
        param_tree.build(self, s)
        for k, v in self.state_dict().items():
            if k.endswith("._weight"):
                v.copy_(torch.eye(D))

    def cuda(self, device=None):
        return self

    def _run(self, x, affine_info, inverse):
        dev = x.device if x.is_cuda else _lib.cuda_device(None)
        x = x.detach().to(dev, torch.float32).contiguous()
        a = affine_info.detach().to(dev, torch.float32).contiguous()
        N, D = x.shape
        assert D == 27 and tuple(a.shape) == (N, 2 * D * self.n_layers), (tuple(x.shape), tuple(a.shape))
        out = torch.empty_like(x)
        table, keep = _lib.tensor_table(self.state_dict())
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().bfsr_op_linf_flow(table, len(table), self.n_layers, int(inverse), x.data_ptr(), a.data_ptr(), N,
                                                    out.data_ptr(), _lib.stream_ptr(dev)))
        del keep
        return out

    def forward(self, x, affine_info):
        z = self._run(x, affine_info, False)
        return z, torch.full((z.shape[0],), float("nan"), device=z.device)

    def inverse(self, z, affine_info):
        return self._run(z, affine_info, True)
