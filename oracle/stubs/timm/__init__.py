"""Stub of `timm` (absent offline): only what LINF-LP/models/swinir.py and swin_transformer.py import at module load."""
