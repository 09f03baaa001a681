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


def random_unet_state_dict(seed: int = 0, n_channels: int = 1, n_classes: int = 17, n_last: int = 64) -> SD:
    """Seeded stand-in for a trained checkpoint (the reference's weights live in ClearML): kaiming-like convolutions,
    norm scales near 1 with small offsets."""
    g = torch.Generator().manual_seed(9000 + seed)
    sd: SD = {}

    def conv(name, cout, cin, k):
        sd[name] = torch.randn((cout, cin, k, k), generator=g) * (2.0 / (cin * k * k)) ** 0.5

    def norm(name, c):
        sd[name + ".weight"] = 1.0 + 0.1 * torch.randn((c,), generator=g)
        sd[name + ".bias"] = 0.1 * torch.randn((c,), generator=g)

    def dconv(p, cin, cout):
        conv(p + ".0.weight", cout, cin, 3)
        norm(p + ".1", cout)
        conv(p + ".3.weight", cout, cout, 3)
        norm(p + ".4", cout)

    ch = [64, 128, 256, 512, 1024]
    dconv("inc.double_conv", n_channels, 64)
    for i in range(1, 5):
        dconv(f"down{i}.maxpool_conv.1.double_conv", ch[i - 1], ch[i])
    for i in range(1, 5):
        cin = ch[5 - i]
        cout = ch[4 - i] if i < 4 else n_last
        sd[f"up{i}.up.weight"] = torch.randn((cin, cin // 2, 2, 2), generator=g) * (1.0 / cin) ** 0.5
        sd[f"up{i}.up.bias"] = 0.05 * torch.randn((cin // 2,), generator=g)
        dconv(f"up{i}.conv.double_conv", cin, cout)
    sd["outc.conv.weight"] = torch.randn((n_classes, n_last, 1, 1), generator=g) * (4.0 / n_last) ** 0.5
    sd["outc.conv.bias"] = 0.5 * torch.randn((n_classes,), generator=g)
    return sd


def synthetic_radiograph_small(seed: int, H: int = 384, W: int = 224) -> torch.Tensor:
    """[1,1,H,W] normalised grey image: smooth blobs + noise, like the U-Net input of save_refined_segmentations.py."""
    rng = np.random.default_rng(7000 + seed)
    yy, xx = np.mgrid[0:H, 0:W].astype(np.float32)
    img = np.zeros((H, W), np.float32)
    for _ in range(6):
        cy, cx, s = rng.uniform(0, H), rng.uniform(0, W), rng.uniform(0.05, 0.3) * max(H, W)
        img += rng.uniform(0.2, 0.8) * np.exp(-((yy - cy) ** 2 + (xx - cx) ** 2) / (2 * s * s)).astype(np.float32)
    img = np.clip(img / max(float(img.max()), 1e-6) + 0.05 * rng.standard_normal((H, W)).astype(np.float32), 0, 1)
    return torch.from_numpy(((img - IMG_MEAN) / IMG_STD).astype(np.float32))[None, None]
