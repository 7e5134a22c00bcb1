"""ctypes binding of libbfsr_b200.so (the C ABI declared in include/bfsr_b200.h).

The product path has NO fallback: if the shared library is missing or a call fails
the error is raised to the caller.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# BFSR_LIB_PATH selects an instrumented build of the same library (tools only; e.g. `make trace`)
LIB_PATH = os.environ.get("BFSR_LIB_PATH") or os.path.join(_HERE, "libbfsr_b200.so")


class BfsrError(RuntimeError):
    pass


class Tensor(C.Structure):
    _fields_ = [("name", C.c_char_p), ("data", C.c_void_p), ("ndim", C.c_int32), ("shape", C.c_int64 * 4)]


class SRFlowDesc(C.Structure):
    _fields_ = [("scale", C.c_int32), ("nf", C.c_int32), ("nb", C.c_int32), ("gc", C.c_int32), ("K", C.c_int32),
                ("L", C.c_int32), ("n_no_affine", C.c_int32), ("hidden", C.c_int32), ("n_blocks", C.c_int32),
                ("blocks", C.c_int32 * 8), ("split_enable", C.c_int32), ("tile_chunk", C.c_int32),
                ("precision", C.c_int32)]


class UNetDesc(C.Structure):
    _fields_ = [("variant", C.c_int32), ("depth", C.c_int32), ("dim", C.c_int32), ("bilinear", C.c_int32),
                ("n_latents", C.c_int32), ("latent_ch", C.c_int32 * 4), ("in_chans", C.c_int32),
                ("precision", C.c_int32)]


class LINFDesc(C.Structure):
    _fields_ = [("encoder", C.c_int32), ("nb", C.c_int32), ("hidden", C.c_int32), ("flow_layers", C.c_int32),
                ("patch_size", C.c_int32), ("tile_chunk", C.c_int32), ("precision", C.c_int32)]


_lib = None

# every symbol include/bfsr_b200.h declares: (restype, argtypes)
_P = C.c_void_p
_I = C.c_int32
SYMBOLS = {
    "bfsr_last_error": (C.c_char_p, []),
    "bfsr_version": (C.c_char_p, []),
    "bfsr_launch_count": (C.c_int64, [C.c_int]),
    "bfsr_prof_enable": (C.c_int, [C.c_int]),
    "bfsr_prof_dump": (C.c_int, [C.c_char_p, C.c_int]),
    "bfsr_prof_summary": (C.c_int, [C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_int64)]),
    "bfsr_srflow_create": (C.c_int, [C.POINTER(_P), C.POINTER(SRFlowDesc), C.POINTER(Tensor), _I, _I]),
    "bfsr_srflow_destroy": (None, [_P]),
    "bfsr_srflow_num_latents": (C.c_int, [_P]),
    "bfsr_srflow_latent_shape": (C.c_int, [_P, _I, _I, _I, C.POINTER(_I), C.POINTER(_I), C.POINTER(_I)]),
    "bfsr_srflow_encode": (C.c_int, [_P, _P, _P, _I, _I, _I, C.POINTER(_P), _P]),
    "bfsr_srflow_decode": (C.c_int, [_P, _P, C.POINTER(_P), _I, _I, _I, _P, _P]),
    "bfsr_srflow_lp_sr": (C.c_int, [_P, _P, _P, _I, _I, _I, _P, _P]),
    "bfsr_srflow_lp_sr_host": (C.c_int, [_P, _P, _P, _I, _I, _I, _P, _P]),
    "bfsr_srflow_workspace_bytes": (C.c_int64, [_P]),
    "bfsr_unet_create": (C.c_int, [C.POINTER(_P), C.POINTER(UNetDesc), C.POINTER(Tensor), _I, _I]),
    "bfsr_unet_destroy": (None, [_P]),
    "bfsr_unet_forward_srflow": (C.c_int, [_P, C.POINTER(_P), C.POINTER(_I), C.POINTER(_I), _I, C.POINTER(_P), _P]),
    "bfsr_unet_forward_linf": (C.c_int, [_P, _P, _P, _I, _I, _I, _I, _I, _P, _P]),
    "bfsr_linf_create": (C.c_int, [C.POINTER(_P), C.POINTER(LINFDesc), C.POINTER(Tensor), _I, _I]),
    "bfsr_linf_destroy": (None, [_P]),
    "bfsr_linf_gen_feat": (C.c_int, [_P, _P, _I, _I, _I, _P, _P]),
    "bfsr_linf_query": (C.c_int, [_P, _P, _I, _I, _I, _P, _P, _I, _I, _I, _P, _P, _P]),
    "bfsr_linf_affine": (C.c_int, [_P, _P, _I, _I, _I, _P, _P, _I, _I, _P, _P]),
    "bfsr_linf_flow": (C.c_int, [_P, _P, _P, _I, _I, _I, _I, _P, _P]),
    "bfsr_op_linf_flow": (C.c_int, [C.POINTER(Tensor), _I, _I, _I, _P, _P, C.c_int64, _P, _P]),
    "bfsr_linf_lp_sr": (C.c_int, [_P, _P, _P, _I, _I, _I, _P, _P, _P, _I, _I, _I, _I, _P, _P]),
    "bfsr_linf_build_inputs": (C.c_int, [_P, _I, _I, _I, _I, _I, _I, _I, _P, _P, _P, _P, C.POINTER(C.c_int32), C.POINTER(C.c_int32), _P]),
    "bfsr_linf_lp_sr_host": (C.c_int, [_P, _P, _P, _I, _I, _I, _P, _P, _P, _I, _I, _I, _I, _P, _P]),
    "bfsr_op_conv2d": (C.c_int, [_P, _I, _I, _I, _I, _P, _P, _I, _I, _I, _I, _P, _P]),
    "bfsr_metric_psnr": (C.c_int, [_P, _P, _I, _I, _I, _I, _I, _I, C.c_float, C.POINTER(C.c_double), _P]),
    "bfsr_metric_ssim": (C.c_int, [_P, _P, _I, _I, _I, C.c_float, C.POINTER(C.c_double), _P]),
    "bfsr_imresize_bicubic": (C.c_int, [_P, _I, _I, _I, C.c_double, _P, C.POINTER(C.c_int32), C.POINTER(C.c_int32), _P]),
    "bfsr_imresize_bicubic_u8": (C.c_int, [_P, _I, _I, _I, C.c_double, _P, C.POINTER(C.c_int32), C.POINTER(C.c_int32), _P]),
    "bfsr_metric_ssim_uniform": (C.c_int, [_P, _P, _I, _I, _I, C.c_float, _I, _I, C.POINTER(C.c_double), _P]),
    "bfsr_op_conv2d_hi_lo": (C.c_int, [_P, _P, _I, _I, _I, _I, _I, _P, _P, _I, _I, _P, _P]),
    "bfsr_op_conv2d_up2": (C.c_int, [_P, _I, _I, _I, _I, _P, _P, _I, _I, _P, _P]),
    "bfsr_op_squeeze2d": (C.c_int, [_P, _I, _I, _I, _I, _I, _P, _P]),
    "bfsr_op_flowstep": (C.c_int, [C.POINTER(Tensor), _I, C.c_char_p, _I, _I, _I, _P, _P, _I, _I, _I, _P, _I, _I, _P]),
    "bfsr_op_split2d": (C.c_int, [C.POINTER(Tensor), _I, C.c_char_p, _I, _I, _P, _P, _I, _I, _I, _P, _P, _P]),
}


def lib():
    """Load libbfsr_b200.so (once).  Raises if it has not been built (`make -C bfsr_b200/csrc`)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise BfsrError(f"{LIB_PATH} not found: build it with `make -C bfsr_b200/csrc` "
                            "(or `python -c 'import __graft_entry__ as g; g.build()'`); there is no CPU fallback")
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            f = getattr(l, name)
            f.restype = res
            f.argtypes = args
        _lib = l
    return _lib


def check(rc: int):
    if rc != 0:
        raise BfsrError(lib().bfsr_last_error().decode("utf-8", "replace"))


def tensor_table(sd):
    """state_dict -> (ctypes array of Tensor, keep-alive list).  Non-fp32 entries (BatchNorm counters) are skipped."""
    keep = []
    items = []
    for k, v in sd.items():
        if not torch.is_tensor(v) or not v.dtype.is_floating_point:
            continue
        t = v.detach().to("cpu", torch.float32).contiguous()
        keep.append(t)
        kb = k.encode()
        keep.append(kb)
        tt = Tensor()
        tt.name = kb
        tt.data = t.data_ptr()
        tt.ndim = t.dim()
        if t.dim() > 4:
            raise BfsrError(f"tensor {k} has rank {t.dim()} > 4")
        for i, s in enumerate(t.shape):
            tt.shape[i] = s
        items.append(tt)
    arr = (Tensor * len(items))(*items)
    return arr, keep


def cuda_device(device=None) -> torch.device:
    """Indexed CUDA device: `None`, "cuda" and torch.device("cuda") resolve to the CURRENT device (under torchrun each rank has
    called torch.cuda.set_device(rank), so an un-indexed device must never silently mean GPU 0)."""
    if device is None:
        return torch.device("cuda", torch.cuda.current_device())
    d = torch.device("cuda", device) if isinstance(device, int) else torch.device(device)
    if d.type != "cuda":
        raise BfsrError(f"bfsr_b200 runs on CUDA devices only (got {d}); there is no CPU fallback")
    return d if d.index is not None else torch.device("cuda", torch.cuda.current_device())


def stream_ptr(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def ptr_array(tensors):
    return (C.c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])
