"""Builds nn.Module parameter trees from dotted names so engines expose EXACTLY the reference's
state_dict keys (`load_state_dict(strict=True)` of reference checkpoints works unchanged)."""
from __future__ import annotations

import torch
from torch import nn


class _Node(nn.Module):
    """Name-only container (mirrors the reference's module nesting; holds no compute)."""


def build(root: nn.Module, shapes, buffers=()):
    """shapes: {dotted_name: shape}; names listed in `buffers` become buffers instead of parameters."""
    buffers = set(buffers)
    for name, shape in shapes.items():
        parts = name.split(".")
        m = root
        for p in parts[:-1]:
            if p not in m._modules:
                m.add_module(p, _Node())
            m = m._modules[p]
        leaf = parts[-1]
        if name in buffers:
            dtype = torch.int64 if leaf == "num_batches_tracked" else torch.float32
            m.register_buffer(leaf, torch.zeros(tuple(shape), dtype=dtype))
        else:
            m.register_parameter(leaf, nn.Parameter(torch.zeros(tuple(shape)), requires_grad=False))
    return root
