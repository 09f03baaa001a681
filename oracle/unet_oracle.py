"""CPU oracle for the U-Net (test infrastructure only; never imported by the product path).

Functional, state-dict driven restatement of custom_arcitecture/classic_u_net.py (bilinear=False) in torch fp32.
Pinned against the reference's own `UNet` class by tests/golden/make_golden_unet.py -> tests/golden/unet_golden.npz
(tests/test_unet_oracle.py)."""
from __future__ import annotations

from typing import Dict

import numpy as np
import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]
IMG_MEAN, IMG_STD = 0.3505533917353781, 0.22763733675869177  # scripts/seg_grazpedwri_dataset.py:22-23


def _double_conv(sd: SD, p: str, x: torch.Tensor) -> torch.Tensor:
    """classic_u_net.py:9-27: (conv3x3 no bias, InstanceNorm2d affine eps 1e-5, LeakyReLU 0.01) x 2"""
    for conv, norm in (("0", "1"), ("3", "4")):
        x = F.conv2d(x, sd[f"{p}.{conv}.weight"], None, padding=1)
        x = F.instance_norm(x, weight=sd[f"{p}.{norm}.weight"], bias=sd[f"{p}.{norm}.bias"], eps=1e-5)
        x = F.leaky_relu(x, 0.01)
    return x


def unet_forward(sd: SD, x: torch.Tensor) -> torch.Tensor:
    """classic_u_net.py:108-119 -> logits [B, n_classes, H, W] (H, W multiples of 16: no skip padding needed)."""
    with torch.no_grad():
        x1 = _double_conv(sd, "inc.double_conv", x)
        skips = [x1]
        cur = x1
        for i in range(1, 5):
            cur = _double_conv(sd, f"down{i}.maxpool_conv.1.double_conv", F.max_pool2d(cur, 2))
            skips.append(cur)
        for i in range(1, 5):
            up = F.conv_transpose2d(cur, sd[f"up{i}.up.weight"], sd[f"up{i}.up.bias"], stride=2)  # :53,57
            cur = _double_conv(sd, f"up{i}.conv.double_conv", torch.cat([skips[4 - i], up], dim=1))  # :68-69
        return F.conv2d(cur, sd["outc.conv.weight"], sd["outc.conv.bias"])  # :72-78


from samcarriestheburden_b200.synthetic import random_unet_state_dict, synthetic_radiograph_small  # noqa: E402,F401
