"""Shared helpers for the parity tests (the oracle is the checker, never the thing under test)."""
import os

import numpy as np
import torch

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def rel_l2(ref, got):
    ref = torch.as_tensor(ref).double().cpu()
    got = torch.as_tensor(got).double().cpu()
    return float((ref - got).norm() / ref.norm().clamp_min(1e-30))


def max_abs(ref, got):
    return float((torch.as_tensor(ref).double().cpu() - torch.as_tensor(got).double().cpu()).abs().max())


def golden(name):
    return np.load(os.path.join(GOLD, name + ".npz"))


SMALL = dict(nb=4, blocks=(0, 1, 2, 3), K=2)
