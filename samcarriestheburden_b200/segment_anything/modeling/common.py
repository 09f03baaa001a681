"""Parameter containers shared by the SAM modules (reference: segment_anything/modeling/common.py).

The modules of this package own the reference-named parameters (so reference checkpoints load with
`load_state_dict(strict=True)`) but carry no PyTorch arithmetic: the top-level modules hand raw device
pointers to libb200sam.so.  Calling a fused-away sub-module directly raises.
"""
from __future__ import annotations

import torch
import torch.nn as nn


class FusedAway(nn.Module):
    """Base for sub-modules whose math runs inside a parent's fused CUDA path."""

    def forward(self, *args, **kwargs):  # pragma: no cover - guard rail
        raise NotImplementedError(
            f"{type(self).__name__} has no standalone forward in b200sam: it is fused into the CUDA path of its "
            "parent module (ImageEncoderViT.forward / Sam.decode_prompts)")


class MLPBlock(FusedAway):
    """lin1 -> act -> lin2 (reference common.py:13-26)."""

    def __init__(self, embedding_dim: int, mlp_dim: int, act=nn.GELU) -> None:
        super().__init__()
        self.lin1 = nn.Linear(embedding_dim, mlp_dim)
        self.lin2 = nn.Linear(mlp_dim, embedding_dim)
        self.act = act()


class LayerNorm2d(FusedAway):
    """Channel-dim LayerNorm of NCHW tensors, eps inside the sqrt (reference common.py:31-43)."""

    def __init__(self, num_channels: int, eps: float = 1e-6) -> None:
        super().__init__()
        self.weight = nn.Parameter(torch.ones(num_channels))
        self.bias = nn.Parameter(torch.zeros(num_channels))
        self.eps = eps
