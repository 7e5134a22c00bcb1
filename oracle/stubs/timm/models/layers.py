import torch.nn as nn


class DropPath(nn.Identity):
    def __init__(self, *a, **k):
        super().__init__()


def to_2tuple(x):
    return tuple(x) if isinstance(x, (tuple, list)) else (x, x)


def trunc_normal_(t, std=0.02, **k):
    return nn.init.normal_(t, std=std)
