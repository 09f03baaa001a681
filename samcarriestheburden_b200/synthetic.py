"""Synthetic inputs of the shapes SURVEY.md 8(d) specifies (there is no network for datasets or checkpoints): seeded
random-init SAM / U-Net weights with the reference's parameter names and shapes, synthetic 1024^2 radiographs, synthetic
U-Net masks / probability maps.  Pure data generators (numpy / torch CPU), shared by the benchmark, the profiling tools
and - through re-export from `oracle/` - the parity tests, so that the product arm of `bench.py` never imports the oracle."""
from __future__ import annotations

import math
from typing import Dict

import numpy as np
import torch

SD = Dict[str, torch.Tensor]

VIT_CONFIGS = {
    "vit_h": dict(embed_dim=1280, depth=32, num_heads=16, global_attn_indexes=(7, 15, 23, 31)),
    "vit_l": dict(embed_dim=1024, depth=24, num_heads=16, global_attn_indexes=(5, 11, 17, 23)),
    "vit_b": dict(embed_dim=768, depth=12, num_heads=12, global_attn_indexes=(2, 5, 8, 11)),
}


def random_state_dict(model_type: str = "vit_b", seed: int = 0) -> SD:
    """Random-init weights with the reference's parameter names/shapes (build_sam.py:55-107), PyTorch
    default initialisers, plus N(0, 0.02) pos_embed / rel_pos tables (zero at init in the reference,
    image_encoder.py:68-70,221-222 — randomised so the rel-pos path is exercised)."""
    cfg = VIT_CONFIGS[model_type]
    g = torch.Generator().manual_seed(seed)
    sd: SD = {}

    def lin(name, out_f, in_f, bias=True):
        bound = 1.0 / math.sqrt(in_f)
        sd[name + ".weight"] = (torch.rand((out_f, in_f), generator=g) * 2 - 1) * bound
        if bias:
            sd[name + ".bias"] = (torch.rand((out_f,), generator=g) * 2 - 1) * bound

    def conv(name, out_c, in_c, k, bias=True, transposed=False):
        fan_in = (out_c if transposed else in_c) * k * k
        bound = 1.0 / math.sqrt(fan_in)
        shape = (in_c, out_c, k, k) if transposed else (out_c, in_c, k, k)
        sd[name + ".weight"] = (torch.rand(shape, generator=g) * 2 - 1) * bound
        if bias:
            sd[name + ".bias"] = (torch.rand((out_c,), generator=g) * 2 - 1) * bound

    def norm(name, n):
        sd[name + ".weight"] = 1.0 + 0.1 * torch.randn((n,), generator=g)
        sd[name + ".bias"] = 0.1 * torch.randn((n,), generator=g)

    D, depth, heads = cfg["embed_dim"], cfg["depth"], cfg["num_heads"]
    hd = D // heads
    ie = "image_encoder."
    conv(ie + "patch_embed.proj", D, 3, 16)
    sd[ie + "pos_embed"] = 0.02 * torch.randn((1, 64, 64, D), generator=g)
    for i in range(depth):
        b = f"{ie}blocks.{i}."
        S = 64 if i in cfg["global_attn_indexes"] else 14
        norm(b + "norm1", D)
        lin(b + "attn.qkv", 3 * D, D)
        lin(b + "attn.proj", D, D)
        sd[b + "attn.rel_pos_h"] = 0.02 * torch.randn((2 * S - 1, hd), generator=g)
        sd[b + "attn.rel_pos_w"] = 0.02 * torch.randn((2 * S - 1, hd), generator=g)
        norm(b + "norm2", D)
        lin(b + "mlp.lin1", 4 * D, D)
        lin(b + "mlp.lin2", D, 4 * D)
    conv(ie + "neck.0", 256, D, 1, bias=False)
    norm(ie + "neck.1", 256)
    conv(ie + "neck.2", 256, 256, 3, bias=False)
    norm(ie + "neck.3", 256)

    pe = "prompt_encoder."
    sd[pe + "pe_layer.positional_encoding_gaussian_matrix"] = torch.randn((2, 128), generator=g)
    for i in range(4):
        sd[f"{pe}point_embeddings.{i}.weight"] = torch.randn((1, 256), generator=g)
    sd[pe + "not_a_point_embed.weight"] = torch.randn((1, 256), generator=g)
    sd[pe + "no_mask_embed.weight"] = torch.randn((1, 256), generator=g)
    conv(pe + "mask_downscaling.0", 4, 1, 2)
    norm(pe + "mask_downscaling.1", 4)
    conv(pe + "mask_downscaling.3", 16, 4, 2)
    norm(pe + "mask_downscaling.4", 16)
    conv(pe + "mask_downscaling.6", 256, 16, 1)

    d = "mask_decoder."
    sd[d + "iou_token.weight"] = torch.randn((1, 256), generator=g)
    sd[d + "mask_tokens.weight"] = torch.randn((4, 256), generator=g)
    t = d + "transformer."

    def attn(name, internal):
        for pr in ("q_proj", "k_proj", "v_proj"):
            lin(f"{name}.{pr}", internal, 256)
        lin(f"{name}.out_proj", 256, internal)

    for i in range(2):
        L = f"{t}layers.{i}."
        attn(L + "self_attn", 256)
        norm(L + "norm1", 256)
        attn(L + "cross_attn_token_to_image", 128)
        norm(L + "norm2", 256)
        lin(L + "mlp.lin1", 2048, 256)
        lin(L + "mlp.lin2", 256, 2048)
        norm(L + "norm3", 256)
        norm(L + "norm4", 256)
        attn(L + "cross_attn_image_to_token", 128)
    attn(t + "final_attn_token_to_image", 128)
    norm(t + "norm_final_attn", 256)
    conv(d + "output_upscaling.0", 64, 256, 2, transposed=True)
    norm(d + "output_upscaling.1", 64)
    conv(d + "output_upscaling.3", 32, 64, 2, transposed=True)
    for i in range(4):
        m = f"{d}output_hypernetworks_mlps.{i}.layers."
        lin(m + "0", 256, 256); lin(m + "1", 256, 256); lin(m + "2", 32, 256)
    m = d + "iou_prediction_head.layers."
    lin(m + "0", 256, 256); lin(m + "1", 256, 256); lin(m + "2", 4, 256)
    return sd


def synthetic_radiograph(seed: int, h: int = 1024, w: int = 1024) -> np.ndarray:
    """Smooth blobs + noise, grayscale replicated to RGB like generate_img_embeddings.py:39-40. uint8 HWC."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    img = np.zeros((h, w), np.float32)
    for _ in range(6):
        cy, cx = rng.uniform(0, h), rng.uniform(0, w)
        sy, sx = rng.uniform(h / 12, h / 3), rng.uniform(w / 12, w / 3)
        img += rng.uniform(0.3, 1.0) * np.exp(-(((yy - cy) / sy) ** 2 + ((xx - cx) / sx) ** 2))
    img = img / max(float(img.max()), 1e-6) * 200.0 + rng.uniform(0, 40, size=(h, w)).astype(np.float32)
    g = np.clip(img, 0, 255).astype(np.uint8)
    return np.repeat(g[:, :, None], 3, axis=2)


def synthetic_unet_masks(seed: int, C: int = 17, H: int = 384, W: int = 224) -> np.ndarray:
    """Bool [C,H,W]: one ellipse per class (+ optional distractor blob), some overlaps, 1-2 empty classes."""
    rng = np.random.default_rng(1000 + seed)
    yy, xx = np.mgrid[0:H, 0:W].astype(np.float32)
    m = np.zeros((C, H, W), bool)
    empty = set(rng.choice(C, size=int(rng.integers(1, 3)), replace=False).tolist())
    for c in range(C):
        if c in empty:
            continue
        cy, cx = rng.uniform(0.1 * H, 0.9 * H), rng.uniform(0.15 * W, 0.85 * W)
        ry, rx = rng.uniform(8, 45), rng.uniform(6, 30)
        m[c] = ((yy - cy) / ry) ** 2 + ((xx - cx) / rx) ** 2 <= 1.0
        if rng.random() < 0.4:
            by, bx = rng.uniform(0, H), rng.uniform(0, W)
            m[c] |= ((yy - by) / 4.0) ** 2 + ((xx - bx) / 4.0) ** 2 <= 1.0
    return m



def synthetic_unet_probs(seed: int, C: int = 17, H: int = 384, W: int = 224) -> np.ndarray:
    """float32 [C,H,W] probabilities: per class a main ellipse (p in [0.6, 0.95]) + 0-2 distractor blobs of higher or
    lower confidence, smooth background < 0.5, 1-2 empty classes."""
    rng = np.random.default_rng(5000 + seed)
    yy, xx = np.mgrid[0:H, 0:W].astype(np.float32)
    out = np.zeros((C, H, W), np.float32)
    empty = set(rng.choice(C, size=int(rng.integers(1, 3)), replace=False).tolist())
    for c in range(C):
        p = (0.05 + 0.3 * rng.random((H, W))).astype(np.float32)
        if c not in empty:
            cy, cx = rng.uniform(0.1 * H, 0.9 * H), rng.uniform(0.15 * W, 0.85 * W)
            ry, rx = rng.uniform(8, 45), rng.uniform(6, 30)
            core = ((yy - cy) / ry) ** 2 + ((xx - cx) / rx) ** 2 <= 1.0
            p[core] = (rng.uniform(0.6, 0.95) + 0.04 * rng.standard_normal(int(core.sum()))).astype(np.float32)
            for _ in range(int(rng.integers(0, 3))):
                by, bx, br = rng.uniform(0, H), rng.uniform(0, W), rng.uniform(1.5, 9.0)
                blob = ((yy - by) / br) ** 2 + ((xx - bx) / br) ** 2 <= 1.0
                p[blob] = (rng.uniform(0.55, 0.99) + 0.02 * rng.standard_normal(int(blob.sum()))).astype(np.float32)
        out[c] = np.clip(p, 0.0, 1.0)
    return out


# ----------------------------------------------------------------------------------------------- U-Net
IMG_MEAN, IMG_STD = 0.3505533917353781, 0.22763733675869177  # scripts/seg_grazpedwri_dataset.py:22-23


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
