"""CPU oracle for the SAM pseudo-label refinement hot path.  TEST INFRASTRUCTURE ONLY.

A functional, state-dict-driven restatement (plain torch fp32 on CPU + numpy integer code) of the reference
algorithm in multimodallearning/SamCarriesTheBurden.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / `--impl reference` leg may import this module — never the product path.

Pinning: the reference ships no golden vectors or tests, so the oracle is pinned against
OUTPUTS OF THE REFERENCE ITSELF: tests/golden/make_golden.py imports /root/reference in the build container,
runs its modules on seeded inputs and stores inputs + outputs as fixtures; tests/test_oracle_golden.py checks
this file against those fixtures (and directly against the live reference when /root/reference exists).

Every function cites the reference file:line it follows (paths relative to the reference root).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]

# segment_anything/build_sam.py:14-44
from samcarriestheburden_b200.synthetic import (VIT_CONFIGS, random_state_dict, synthetic_radiograph,  # noqa: E402,F401
                                            synthetic_unet_masks, synthetic_unet_probs)
PIXEL_MEAN = (123.675, 116.28, 103.53)  # build_sam.py:99
PIXEL_STD = (58.395, 57.12, 57.375)     # build_sam.py:100


# ----------------------------------------------------------------------------------------------- encoder
def preprocess(x: torch.Tensor, img_size: int = 1024) -> torch.Tensor:
    """modeling/sam.py:164-174 — normalise then zero-pad bottom/right to img_size."""
    mean = torch.tensor(PIXEL_MEAN).view(-1, 1, 1)
    std = torch.tensor(PIXEL_STD).view(-1, 1, 1)
    x = (x - mean) / std
    h, w = x.shape[-2:]
    return F.pad(x, (0, img_size - w, 0, img_size - h))


def _rel_pos_bias(q: torch.Tensor, rel_h: torch.Tensor, rel_w: torch.Tensor, S: int) -> torch.Tensor:
    """modeling/image_encoder.py:292-361 for q_size == k_size == (S, S): table length is 2S-1 so no
    interpolation; bias[b, (qh,qw), (kh,kw)] = q . rel_h[qh-kh+S-1] + q . rel_w[qw-kw+S-1]."""
    assert rel_h.shape[0] == 2 * S - 1 and rel_w.shape[0] == 2 * S - 1
    idx = torch.arange(S)[:, None] - torch.arange(S)[None, :] + (S - 1)
    Rh, Rw = rel_h[idx], rel_w[idx]  # [S, S, hd]
    Bn, _, hd = q.shape
    rq = q.reshape(Bn, S, S, hd)
    bh = torch.einsum("bhwc,hkc->bhwk", rq, Rh)
    bw = torch.einsum("bhwc,wkc->bhwk", rq, Rw)
    return (bh[:, :, :, :, None] + bw[:, :, :, None, :]).reshape(Bn, S * S, S * S)


def _encoder_attention(sd: SD, p: str, x: torch.Tensor, heads: int) -> torch.Tensor:
    """modeling/image_encoder.py:224-240.  x: [B', S, S, D]."""
    Bn, S, _, D = x.shape
    hd = D // heads
    qkv = F.linear(x, sd[p + "qkv.weight"], sd[p + "qkv.bias"])
    qkv = qkv.reshape(Bn, S * S, 3, heads, hd).permute(2, 0, 3, 1, 4).reshape(3, Bn * heads, S * S, hd)
    q, k, v = qkv[0], qkv[1], qkv[2]
    attn = (q * hd ** -0.5) @ k.transpose(-2, -1)
    attn = attn + _rel_pos_bias(q, sd[p + "rel_pos_h"], sd[p + "rel_pos_w"], S)  # unscaled q (:234)
    attn = attn.softmax(dim=-1)
    out = (attn @ v).view(Bn, heads, S, S, hd).permute(0, 2, 3, 1, 4).reshape(Bn, S, S, D)
    return F.linear(out, sd[p + "proj.weight"], sd[p + "proj.bias"])


def _encoder_block(sd: SD, p: str, x: torch.Tensor, heads: int, window: int) -> torch.Tensor:
    """modeling/image_encoder.py:166-182 with window partition :243-289 (pad AFTER norm1)."""
    B, H, W, D = x.shape
    y = F.layer_norm(x, (D,), sd[p + "norm1.weight"], sd[p + "norm1.bias"], eps=1e-6)
    if window > 0:
        ph, pw = (-H) % window, (-W) % window
        y = F.pad(y, (0, 0, 0, pw, 0, ph))
        Hp, Wp = H + ph, W + pw
        y = y.view(B, Hp // window, window, Wp // window, window, D).permute(0, 1, 3, 2, 4, 5)
        y = y.reshape(-1, window, window, D)
        y = _encoder_attention(sd, p + "attn.", y, heads)
        y = y.view(B, Hp // window, Wp // window, window, window, D).permute(0, 1, 3, 2, 4, 5)
        y = y.reshape(B, Hp, Wp, D)[:, :H, :W, :]
    else:
        y = _encoder_attention(sd, p + "attn.", y, heads)
    x = x + y
    z = F.layer_norm(x, (D,), sd[p + "norm2.weight"], sd[p + "norm2.bias"], eps=1e-6)
    z = F.linear(F.gelu(F.linear(z, sd[p + "mlp.lin1.weight"], sd[p + "mlp.lin1.bias"])),
                 sd[p + "mlp.lin2.weight"], sd[p + "mlp.lin2.bias"])  # common.py:25-26 (erf GELU)
    return x + z


def layer_norm_2d(x: torch.Tensor, w: torch.Tensor, b: torch.Tensor, eps: float = 1e-6) -> torch.Tensor:
    """modeling/common.py:31-43 — LayerNorm over the channel dim of NCHW (biased variance)."""
    u = x.mean(1, keepdim=True)
    s = (x - u).pow(2).mean(1, keepdim=True)
    return w[:, None, None] * ((x - u) / torch.sqrt(s + eps)) + b[:, None, None]


@torch.no_grad()
def image_encoder(sd: SD, x: torch.Tensor, depth: int, num_heads: int, global_attn_indexes: Sequence[int],
                  window_size: int = 14, prefix: str = "image_encoder.", **_unused) -> torch.Tensor:
    """modeling/image_encoder.py:106-116.  x: preprocessed [B,3,1024,1024] -> [B,256,64,64]."""
    p = prefix
    t = F.conv2d(x, sd[p + "patch_embed.proj.weight"], sd[p + "patch_embed.proj.bias"],
                 stride=sd[p + "patch_embed.proj.weight"].shape[-1]).permute(0, 2, 3, 1)
    t = t + sd[p + "pos_embed"]
    for i in range(depth):
        win = 0 if i in global_attn_indexes else window_size
        t = _encoder_block(sd, f"{p}blocks.{i}.", t, num_heads, win)
    t = t.permute(0, 3, 1, 2)
    t = F.conv2d(t, sd[p + "neck.0.weight"])
    t = layer_norm_2d(t, sd[p + "neck.1.weight"], sd[p + "neck.1.bias"])
    t = F.conv2d(t, sd[p + "neck.2.weight"], padding=1)
    return layer_norm_2d(t, sd[p + "neck.3.weight"], sd[p + "neck.3.bias"])


# ----------------------------------------------------------------------------------------------- prompt encoder
def _pe(sd: SD, coords01: torch.Tensor) -> torch.Tensor:
    """modeling/prompt_encoder.py:185-192."""
    c = 2 * coords01 - 1
    c = c @ sd["prompt_encoder.pe_layer.positional_encoding_gaussian_matrix"]
    c = 2 * np.pi * c
    return torch.cat([torch.sin(c), torch.cos(c)], dim=-1)


def dense_pe(sd: SD, size: int = 64) -> torch.Tensor:
    """modeling/prompt_encoder.py:62-71,194-206 -> [1, 256, size, size]."""
    g = torch.ones((size, size), dtype=torch.float32)
    y = (g.cumsum(0) - 0.5) / size
    x = (g.cumsum(1) - 0.5) / size
    return _pe(sd, torch.stack([x, y], dim=-1)).permute(2, 0, 1).unsqueeze(0)


def _pe_coords(sd: SD, coords: torch.Tensor, img: int = 1024) -> torch.Tensor:
    c = coords.clone().to(torch.float)
    c[:, :, 0] = c[:, :, 0] / img
    c[:, :, 1] = c[:, :, 1] / img
    return _pe(sd, c)


@torch.no_grad()
def prompt_encoder(sd: SD, points: Optional[Tuple[torch.Tensor, torch.Tensor]], boxes: Optional[torch.Tensor],
                   masks: Optional[torch.Tensor], emb_size: int = 64, img_size: int = 1024) -> Tuple[torch.Tensor, torch.Tensor]:
    """modeling/prompt_encoder.py:128-168 -> (sparse [B,N,256], dense [B,256,64,64])."""
    pe = "prompt_encoder."
    if points is not None:
        bs = points[0].shape[0]
    elif boxes is not None:
        bs = boxes.shape[0]
    elif masks is not None:
        bs = masks.shape[0]
    else:
        bs = 1
    C = sd[pe + "no_mask_embed.weight"].shape[1]
    sparse = torch.empty((bs, 0, C))
    if points is not None:
        coords, labels = points
        coords = coords + 0.5
        if boxes is None:  # pad point, label -1 (:81-85)
            coords = torch.cat([coords, torch.zeros((coords.shape[0], 1, 2))], dim=1)
            labels = torch.cat([labels, -torch.ones((labels.shape[0], 1))], dim=1)
        e = _pe_coords(sd, coords, img_size)
        e[labels == -1] = 0.0
        e[labels == -1] += sd[pe + "not_a_point_embed.weight"]
        e[labels == 0] += sd[pe + "point_embeddings.0.weight"]
        e[labels == 1] += sd[pe + "point_embeddings.1.weight"]
        sparse = torch.cat([sparse, e], dim=1)
    if boxes is not None:
        c = (boxes + 0.5).reshape(-1, 2, 2)
        e = _pe_coords(sd, c, img_size)
        e[:, 0, :] += sd[pe + "point_embeddings.2.weight"]
        e[:, 1, :] += sd[pe + "point_embeddings.3.weight"]
        sparse = torch.cat([sparse, e], dim=1)
    if masks is not None:
        m = pe + "mask_downscaling."
        d = F.conv2d(masks, sd[m + "0.weight"], sd[m + "0.bias"], stride=2)
        d = F.gelu(layer_norm_2d(d, sd[m + "1.weight"], sd[m + "1.bias"]))
        d = F.conv2d(d, sd[m + "3.weight"], sd[m + "3.bias"], stride=2)
        d = F.gelu(layer_norm_2d(d, sd[m + "4.weight"], sd[m + "4.bias"]))
        dense = F.conv2d(d, sd[m + "6.weight"], sd[m + "6.bias"])
    else:
        dense = sd[pe + "no_mask_embed.weight"].reshape(1, -1, 1, 1).expand(bs, -1, emb_size, emb_size)
    return sparse, dense


# ----------------------------------------------------------------------------------------------- mask decoder
def _dec_attention(sd: SD, p: str, q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, heads: int) -> torch.Tensor:
    """modeling/transformer.py:218-240 (scale applied after q.k)."""
    q = F.linear(q, sd[p + "q_proj.weight"], sd[p + "q_proj.bias"])
    k = F.linear(k, sd[p + "k_proj.weight"], sd[p + "k_proj.bias"])
    v = F.linear(v, sd[p + "v_proj.weight"], sd[p + "v_proj.bias"])

    def split(t):
        b, n, c = t.shape
        return t.reshape(b, n, heads, c // heads).transpose(1, 2)

    q, k, v = split(q), split(k), split(v)
    a = (q @ k.permute(0, 1, 3, 2)) / math.sqrt(q.shape[-1])
    o = torch.softmax(a, dim=-1) @ v
    b, h, n, c = o.shape
    return F.linear(o.transpose(1, 2).reshape(b, n, h * c), sd[p + "out_proj.weight"], sd[p + "out_proj.bias"])


def _ln(sd: SD, p: str, x: torch.Tensor) -> torch.Tensor:
    return F.layer_norm(x, (x.shape[-1],), sd[p + "weight"], sd[p + "bias"], eps=1e-5)


def two_way_transformer(sd: SD, src: torch.Tensor, pos: torch.Tensor, tokens: torch.Tensor, heads: int = 8,
                        depth: int = 2, prefix: str = "mask_decoder.transformer.") -> Tuple[torch.Tensor, torch.Tensor]:
    """modeling/transformer.py:62-106 and :151-182."""
    keys = src.flatten(2).permute(0, 2, 1)
    kpe = pos.flatten(2).permute(0, 2, 1)
    queries, qpe = tokens, tokens
    for i in range(depth):
        L = f"{prefix}layers.{i}."
        if i == 0:
            queries = _dec_attention(sd, L + "self_attn.", queries, queries, queries, heads)
        else:
            q = queries + qpe
            queries = queries + _dec_attention(sd, L + "self_attn.", q, q, queries, heads)
        queries = _ln(sd, L + "norm1.", queries)
        queries = _ln(sd, L + "norm2.", queries + _dec_attention(
            sd, L + "cross_attn_token_to_image.", queries + qpe, keys + kpe, keys, heads))
        mlp = F.linear(F.relu(F.linear(queries, sd[L + "mlp.lin1.weight"], sd[L + "mlp.lin1.bias"])),
                       sd[L + "mlp.lin2.weight"], sd[L + "mlp.lin2.bias"])
        queries = _ln(sd, L + "norm3.", queries + mlp)
        keys = _ln(sd, L + "norm4.", keys + _dec_attention(
            sd, L + "cross_attn_image_to_token.", keys + kpe, queries + qpe, queries, heads))
    queries = _ln(sd, prefix + "norm_final_attn.", queries + _dec_attention(
        sd, prefix + "final_attn_token_to_image.", queries + qpe, keys + kpe, keys, heads))
    return queries, keys


def _mlp(sd: SD, p: str, x: torch.Tensor, n: int = 3) -> torch.Tensor:
    """modeling/mask_decoder.py:154-176."""
    for j in range(n):
        x = F.linear(x, sd[f"{p}layers.{j}.weight"], sd[f"{p}layers.{j}.bias"])
        if j < n - 1:
            x = F.relu(x)
    return x


@torch.no_grad()
def mask_decoder(sd: SD, image_embeddings: torch.Tensor, image_pe: torch.Tensor, sparse: torch.Tensor,
                 dense: torch.Tensor, multimask_output: bool, heads: int = 8) -> Tuple[torch.Tensor, torch.Tensor]:
    """modeling/mask_decoder.py:71-149 -> (masks [B,1|3,4h,4w], iou [B,1|3])."""
    d = "mask_decoder."
    n_mask = sd[d + "mask_tokens.weight"].shape[0]
    out_tokens = torch.cat([sd[d + "iou_token.weight"], sd[d + "mask_tokens.weight"]], dim=0)
    tokens = torch.cat((out_tokens.unsqueeze(0).expand(sparse.size(0), -1, -1), sparse), dim=1)
    src = torch.repeat_interleave(image_embeddings, tokens.shape[0], dim=0) + dense
    pos = torch.repeat_interleave(image_pe, tokens.shape[0], dim=0)
    b, c, h, w = src.shape
    hs, src = two_way_transformer(sd, src, pos, tokens, heads)
    iou_tok, mask_toks = hs[:, 0, :], hs[:, 1:1 + n_mask, :]
    src = src.transpose(1, 2).view(b, c, h, w)
    u = d + "output_upscaling."
    up = F.conv_transpose2d(src, sd[u + "0.weight"], sd[u + "0.bias"], stride=2)
    up = F.gelu(layer_norm_2d(up, sd[u + "1.weight"], sd[u + "1.bias"]))
    up = F.gelu(F.conv_transpose2d(up, sd[u + "3.weight"], sd[u + "3.bias"], stride=2))
    hyper = torch.stack([_mlp(sd, f"{d}output_hypernetworks_mlps.{i}.", mask_toks[:, i, :]) for i in range(n_mask)], 1)
    b, c, h, w = up.shape
    masks = (hyper @ up.view(b, c, h * w)).view(b, -1, h, w)
    iou = _mlp(sd, d + "iou_prediction_head.", iou_tok)
    sl = slice(1, None) if multimask_output else slice(0, 1)
    return masks[:, sl], iou[:, sl]


# ----------------------------------------------------------------------------------------------- post-processing
def postprocess_masks(masks: torch.Tensor, input_size: Sequence[int], original_size: Sequence[int],
                      img_size: int = 1024) -> torch.Tensor:
    """modeling/sam.py:133-162 == sam_mask_decoder_head.py:106-135."""
    m = F.interpolate(masks, (img_size, img_size), mode="bilinear", align_corners=False)
    m = m[..., : input_size[0], : input_size[1]]
    return F.interpolate(m, tuple(original_size), mode="bilinear", align_corners=False)


def get_preprocess_shape(oldh: int, oldw: int, long_side: int = 1024) -> Tuple[int, int]:
    """utils/transforms.py:93-102."""
    scale = long_side * 1.0 / max(oldh, oldw)
    return int(oldh * scale + 0.5), int(oldw * scale + 0.5)


# ----------------------------------------------------------------------------------------------- mask statistics
def stability_score(masks: np.ndarray, mask_threshold: float, threshold_offset: float) -> np.ndarray:
    """utils/amg.py:154-176: count(x > thr + off) / count(x > thr - off) over the last two axes.  torch compares the fp32
    tensor with the Python scalar rounded to fp32 and divides the two int32 counts as fp32 (0 / 0 -> NaN)."""
    x = np.asarray(masks, np.float32)
    hi, lo = np.float32(mask_threshold + threshold_offset), np.float32(mask_threshold - threshold_offset)
    inter = (x > hi).sum(axis=(-2, -1)).astype(np.int32)
    union = (x > lo).sum(axis=(-2, -1)).astype(np.int32)
    with np.errstate(divide="ignore", invalid="ignore"):
        return inter.astype(np.float32) / union.astype(np.float32)


def mask_to_box(masks: np.ndarray) -> np.ndarray:
    """utils/amg.py:303-346: XYXY box of the set pixels of every [..., H, W] mask (int64), zeros for an empty mask."""
    m = np.asarray(masks).astype(bool)
    lead = m.shape[:-2]
    flat = m.reshape((-1,) + m.shape[-2:])
    out = np.zeros((flat.shape[0], 4), np.int64)
    for i, one in enumerate(flat):
        ys, xs = np.nonzero(one)
        if ys.size:
            out[i] = (xs.min(), ys.min(), xs.max(), ys.max())
    return out.reshape(lead + (4,))


# ----------------------------------------------------------------------------------------------- image ingest
_PIL_PRECISION_BITS = 32 - 8 - 2


def pil_resize_coeffs(in_size: int, out_size: int):
    """Pillow src/libImaging/Resample.c `precompute_coeffs` + `normalize_coeffs_8bpc` with the bilinear (triangle)
    filter (support 1.0), full-image box: -> (bounds int32 [out, 2] = (first tap, tap count), kk int32 [out, ksize]).
    Pillow is the un-vendored dependency behind utils/transforms.py:26-31 (torchvision `resize(to_pil_image(img))`);
    this restates its published algorithm in float64 / integers and is pinned against Pillow itself
    (tests/golden/make_golden_resize.py -> tests/golden/resize_golden.npz, tests/test_resize_oracle.py)."""
    scale = float(np.float32(in_size)) / out_size
    filterscale = max(scale, 1.0)
    support = 1.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), np.int32)
    kk = np.zeros((out_size, ksize), np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = 0.0 + (xx + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)  # C (int) cast truncates toward zero, like int()
        xmax = min(int(center + support + 0.5), in_size) - xmin
        w = [0.0] * ksize
        ww = 0.0
        for x in range(xmax):
            a = abs((x + xmin - center + 0.5) * ss)
            w[x] = 1.0 - a if a < 1.0 else 0.0
            ww += w[x]
        for x in range(xmax):
            if ww != 0.0:
                w[x] /= ww
        bounds[xx] = (xmin, xmax)
        for x in range(ksize):
            v = w[x] * (1 << _PIL_PRECISION_BITS)
            kk[xx, x] = int(-0.5 + v) if w[x] < 0 else int(0.5 + v)
    return bounds, kk


def _pil_resample_axis0(img: np.ndarray, bounds: np.ndarray, kk: np.ndarray) -> np.ndarray:
    """One 8-bit pass along axis 0: clip8((2^21 + sum_k px_k * kk_k) >> 22)."""
    out = np.empty((bounds.shape[0],) + img.shape[1:], np.uint8)
    src = img.astype(np.int64)
    for o in range(bounds.shape[0]):
        lo, n = int(bounds[o, 0]), int(bounds[o, 1])
        acc = np.full(img.shape[1:], 1 << (_PIL_PRECISION_BITS - 1), np.int64)
        for t in range(n):
            acc += src[lo + t] * int(kk[o, t])
        out[o] = np.clip(acc >> _PIL_PRECISION_BITS, 0, 255).astype(np.uint8)
    return out


def resize_bilinear_u8(image: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    """`PIL.Image.resize((out_w, out_h), BILINEAR)` of an HxWxC (or HxW) uint8 image: horizontal pass first, its uint8
    result feeds the vertical pass; a pass whose size does not change is skipped (Resample.c ImagingResample)."""
    img = np.ascontiguousarray(image)
    H, W = img.shape[:2]
    if out_w != W:
        b, k = pil_resize_coeffs(W, out_w)
        img = np.swapaxes(_pil_resample_axis0(np.swapaxes(img, 0, 1), b, k), 0, 1)
    if out_h != H:
        b, k = pil_resize_coeffs(H, out_h)
        img = _pil_resample_axis0(img, b, k)
    return np.ascontiguousarray(img)


def apply_image(image: np.ndarray, long_side: int = 1024) -> np.ndarray:
    """utils/transforms.py:26-31 (ResizeLongestSide.apply_image)."""
    newh, neww = get_preprocess_shape(image.shape[0], image.shape[1], long_side)
    return resize_bilinear_u8(image, newh, neww)


# ----------------------------------------------------------------------------------------------- prompt extraction
@dataclass
class OraclePrompt:
    class_idx: int
    img_size: Tuple[int, int]
    pos_seeds: np.ndarray  # int32 [1, 2] (x, y)
    neg_seeds: np.ndarray  # int32 [K-1, 2]
    box: Optional[np.ndarray]  # int32 [4] xmin, ymin, xmax, ymax


def _round_half_even_f32(s: int, n: int) -> int:
    """fp32(sum) / fp32(n) with IEEE division, then round-half-to-even — what torch's CPU
    `coords.float().mean(0).round().int()` evaluates to (prompt_utils.py:41-42)."""
    q = np.float32(s) / np.float32(n)
    return int(np.rint(q))


def extract_seeds_boxes(mask: np.ndarray):
    """utils/prompt_utils.py:34-67: per-class seed on the non-overlap area, box on the full class mask.
    mask: bool [C,H,W] -> seeds int32 [C,2] (x,y), has_seed [C], boxes int32 [C,4], has_box [C]."""
    mask = np.asarray(mask).astype(bool)
    C = mask.shape[0]
    seeds = np.zeros((C, 2), np.int32)
    boxes = np.zeros((C, 4), np.int32)
    has_seed = np.zeros(C, bool)
    has_box = np.zeros(C, bool)
    if C == 0 or mask[0].size == 0:
        return seeds, has_seed, boxes, has_box
    non_overlap = mask.sum(0) < 2
    for c in range(C):
        rows, cols = np.nonzero(mask[c] & non_overlap)
        if rows.size:
            has_seed[c] = True
            seeds[c] = (_round_half_even_f32(int(cols.sum()), cols.size), _round_half_even_f32(int(rows.sum()), rows.size))
        rows, cols = np.nonzero(mask[c])
        if rows.size:
            has_box[c] = True
            boxes[c] = (cols.min(), rows.min(), cols.max(), rows.max())
    return seeds, has_seed, boxes, has_box


def prompt_extract(mask: np.ndarray) -> List[OraclePrompt]:
    """utils/prompt_utils.py:112-143 (seeds=True, boxes=True, mask=False)."""
    seeds, has_seed, boxes, has_box = extract_seeds_boxes(mask)
    idx = [c for c in range(mask.shape[0]) if has_seed[c]]
    out = []
    for c in idx:
        others = [seeds[i] for i in idx if i != c]
        if not others:  # torch.cat of an empty list raises in the reference (:122-123)
            raise ValueError("torch.cat(): expected a non-empty list of Tensors")
        out.append(OraclePrompt(c, tuple(mask.shape[-2:]), seeds[c][None].copy(), np.stack(others).astype(np.int32),
                                boxes[c].copy() if has_box[c] else None))
    return out


def scale_coords(coords: torch.Tensor, original_size: Sequence[int], target_size: Sequence[int]) -> torch.Tensor:
    """utils/prompt_utils.py:146-166 — fp32 (target/original) flipped to (x, y), then multiply."""
    o = torch.tensor(original_size, dtype=torch.float)
    t = torch.tensor(target_size, dtype=torch.float)
    return coords.float() * (t / o).flip(-1)


# ----------------------------------------------------------------------------------------------- U-Net ingest resize
# scripts/save_refined_segmentations.py:63 `cv2.resize(img, (W, H), interpolation=cv2.INTER_LINEAR)` on the uint8 grey
# radiograph.  OpenCV (environment.yml: opencv 4.9) is an un-vendored dependency; its published uint8 linear path
# (modules/imgproc/src/resize.cpp: resizeGeneric_ with HResizeLinear / VResizeLinear<uchar, int, short, FixedPtCast>)
# is restated here: 11-bit fixed-point coefficients `saturate_cast<short>(w * 2048)` from the fp32 fractional position
# ((d + 0.5) * scale - 0.5 in double, rounded to float), horizontal taps clamped with the weight forced to (1, 0) at the
# borders, vertical ROWS clamped but the weights kept (two truncating products of the same row at the borders),
# horizontal pass in exact int32, vertical pass ((b0 * (S0 >> 4)) >> 16) + ((b1 * (S1 >> 4)) >> 16) + 2) >> 2.
# Pinned bit-exactly against cv2 itself by tests/golden/make_golden_cv2resize.py -> tests/golden/cv2resize_golden.npz.
def cv2_linear_coeffs(ssize: int, dsize: int, clamp_weights: bool):
    """-> (i0 [dsize], i1 [dsize], w [dsize, 2] int32) of one axis."""
    d = np.arange(dsize, dtype=np.float64)
    f = ((d + 0.5) * (ssize / dsize) - 0.5).astype(np.float32)
    s = np.floor(f).astype(np.int64)
    f = (f - s.astype(np.float32)).astype(np.float32)
    if clamp_weights:
        lo, hi = s < 0, s >= ssize - 1
        f = np.where(lo | hi, np.float32(0), f).astype(np.float32)
        s = np.where(lo, 0, np.where(hi, ssize - 1, s))
    w1 = np.clip(np.rint(f * np.float32(2048)), -32768, 32767).astype(np.int32)
    w0 = np.clip(np.rint((np.float32(1.0) - f) * np.float32(2048)), -32768, 32767).astype(np.int32)
    i0 = np.clip(s, 0, ssize - 1)
    i1 = np.clip(s + 1, 0, ssize - 1)
    return i0, i1, np.stack([w0, w1], axis=1)


def cv2_resize_linear_u8(img: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    """uint8 [H, W] -> uint8 [out_h, out_w], bit-exact with cv2.resize(..., interpolation=cv2.INTER_LINEAR)."""
    assert img.dtype == np.uint8 and img.ndim == 2
    H, W = img.shape
    x0, x1, xw = cv2_linear_coeffs(W, out_w, True)
    y0, y1, yw = cv2_linear_coeffs(H, out_h, False)
    im = img.astype(np.int64)
    rows = im[:, x0] * xw[:, 0][None, :] + im[:, x1] * xw[:, 1][None, :]
    b0, b1 = yw[:, 0].astype(np.int64)[:, None], yw[:, 1].astype(np.int64)[:, None]
    out = (((b0 * (rows[y0] >> 4)) >> 16) + ((b1 * (rows[y1] >> 4)) >> 16) + 2) >> 2
    return np.clip(out, 0, 255).astype(np.uint8)


# ----------------------------------------------------------------------------------------------- MedSAM ingest
# scripts/generate_img_embeddings.py:49-64 (the reference's DEFAULT sam_type = 'medsam', :16): cv2.resize(RGB uint8,
# (1024, 1024), INTER_CUBIC) -> min-max normalise to [0, 1] in float64 -> float32 [1, 3, 1024, 1024] -> image_encoder
# (Sam.preprocess is bypassed: no mean / std, no padding).  OpenCV's published uint8 cubic path (resize.cpp:
# interpolateCubic with A = -0.75 in fp32, 11-bit weights saturate_cast<short>(c * 2048), HResizeCubic in int32 with the
# tap indices clamped to the image, VResizeCubic's vector body for uchar: fp32 ((S3 b3 + S2 b2) + S1 b1) + S0 b0 with
# b = w / 2^22, round half to even, saturate) is restated; pinned bit-exactly against cv2 with Intel IPP switched off
# (conda's opencv 4.9 of environment.yml has no IPP; pip wheels route cubic through IPP, which differs by 1 LSB in ~4 %
# of the pixels) by tests/golden/make_golden_cv2resize.py.
def cv2_cubic_coeffs(ssize: int, dsize: int):
    d = np.arange(dsize, dtype=np.float64)
    f = ((d + 0.5) * (ssize / dsize) - 0.5).astype(np.float32)
    s = np.floor(f).astype(np.int64)
    x = (f - s.astype(np.float32)).astype(np.float32)
    A, one = np.float32(-0.75), np.float32(1)
    c0 = ((A * (x + one) - np.float32(5) * A) * (x + one) + np.float32(8) * A) * (x + one) - np.float32(4) * A
    c1 = ((A + np.float32(2)) * x - (A + np.float32(3))) * x * x + one
    c2 = ((A + np.float32(2)) * (one - x) - (A + np.float32(3))) * (one - x) * (one - x) + one
    c3 = one - c0 - c1 - c2
    w = np.clip(np.rint(np.stack([c0, c1, c2, c3], 1).astype(np.float32) * np.float32(2048)), -32768, 32767).astype(np.int32)
    idx = np.clip(s[:, None] - 1 + np.arange(4)[None, :], 0, ssize - 1)
    return idx, w


def cv2_resize_cubic_u8(img: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    """uint8 [H, W] -> uint8 [out_h, out_w], bit-exact with cv2.resize(..., INTER_CUBIC) of an IPP-less OpenCV."""
    assert img.dtype == np.uint8 and img.ndim == 2
    xi, xw = cv2_cubic_coeffs(img.shape[1], out_w)
    yi, yw = cv2_cubic_coeffs(img.shape[0], out_h)
    im = img.astype(np.int64)
    rows = sum(im[:, xi[:, k]] * xw[:, k].astype(np.int64)[None, :] for k in range(4))
    scale = np.float32(1.0) / np.float32(2048.0 * 2048.0)
    S = [rows[yi[:, k]].astype(np.float32) for k in range(4)]
    b = [(yw[:, k].astype(np.float32) * scale).astype(np.float32)[:, None] for k in range(4)]
    acc = (S[3] * b[3]).astype(np.float32)
    for k in (2, 1, 0):
        acc = ((S[k] * b[k]).astype(np.float32) + acc).astype(np.float32)
    return np.clip(np.rint(acc), 0, 255).astype(np.uint8)


def medsam_preprocess(gray: np.ndarray, size: int = 1024) -> torch.Tensor:
    """scripts/generate_img_embeddings.py:39-40,49-62: grey uint8 [H, W] -> float32 [1, 3, size, size] in [0, 1]."""
    r = cv2_resize_cubic_u8(gray, size, size)
    lo, hi = r.min(), r.max()
    norm = (r - lo) / np.clip(hi - lo, a_min=1e-8, a_max=None)  # uint8 difference / float64, like the reference
    return torch.tensor(np.repeat(norm[:, :, None], 3, axis=2)).float().permute(2, 0, 1).unsqueeze(0)


# ----------------------------------------------------------------------------------------------- refinement loop
@torch.no_grad()
def predict_mask(sd: SD, features: torch.Tensor, prompt: OraclePrompt, prompt2use: Sequence[str],
                 input_size: Sequence[int], original_size: Sequence[int], mask_prev: Optional[torch.Tensor] = None):
    """sam_mask_decoder_head.py:37-104 for one prompt (B = 1)."""
    pts, labs = [], []
    if "pos_points" in prompt2use:
        p = scale_coords(torch.from_numpy(prompt.pos_seeds), prompt.img_size, input_size)
        pts.append(p); labs.append(torch.ones(p.shape[0]))
    if "neg_points" in prompt2use:
        p = scale_coords(torch.from_numpy(prompt.neg_seeds), prompt.img_size, input_size)
        pts.append(p); labs.append(torch.zeros(p.shape[0]))
    box = None
    if "box" in prompt2use:
        b = torch.from_numpy(prompt.box).unsqueeze(0).reshape(-1, 2)
        box = scale_coords(b, prompt.img_size, input_size).reshape(-1, 4).float()
    points = (torch.cat(pts).float().unsqueeze(0), torch.cat(labs).int().unsqueeze(0)) if pts else None
    sparse, dense = prompt_encoder(sd, points, box, mask_prev)
    low, iou = mask_decoder(sd, features, dense_pe(sd), sparse, dense, multimask_output=False)
    masks = postprocess_masks(low, input_size, original_size) > 0.0
    return masks, iou, low


@torch.no_grad()
def refine(sd: SD, features: torch.Tensor, seg: np.ndarray, input_size: Sequence[int], original_size: Sequence[int],
           prompts1=("box",), prompts2=("pos_points", "neg_points")):
    """utils/seg_refinement.py:99-116 -> (seg bool [C,H,W], est_dice [C], extras for testing)."""
    seg = np.asarray(seg).astype(bool).copy()
    est = np.full(seg.shape[0], np.nan, np.float32)
    prompts = prompt_extract(seg)
    native, lows = {}, {}
    for p in prompts:
        mask, score, low1 = predict_mask(sd, features, p, prompts1, input_size, original_size)
        low = low1
        if prompts2 is not None:
            mask, score, low = predict_mask(sd, features, p, prompts2, input_size, original_size, low1)
        small = F.interpolate(mask.float(), size=seg.shape[-2:], mode="nearest-exact")
        seg[p.class_idx] = small.squeeze().numpy() > 0.5
        s = float(score.reshape(-1)[0])
        est[p.class_idx] = 2 * s / (1 + s)
        native[p.class_idx] = mask[0, 0].numpy()
        lows[p.class_idx] = (low1[0, 0].numpy(), low[0, 0].numpy())
    return seg, est, native, lows


# ---------------------------------------------------------------------------------------------------------------
# Connected-component pre-processing (reference utils/segmentation_preprocessing.py:7-52).
# The labelling itself lives in kornia 0.7.0 (`kornia.contrib.connected_components`, absent from this image): its
# published algorithm is `num_iterations` rounds of 3x3 max-pooling of the batch-global pixel indices inside the
# mask, so after convergence a component's label is the largest global index it contains.  This restatement labels
# with scipy (8-connectivity) and assigns exactly those converged labels.  Pinned against the reference's own
# selection code driven by a torch restatement of the kornia loop (tests/golden/make_golden_ccl.py).
def ccl_labels(bin_mask: np.ndarray) -> np.ndarray:
    """bool [C,H,W] -> int64 labels like converged kornia.contrib.connected_components on a (C,1,H,W) batch: the
    largest batch-global pixel index of the 8-connected component (0 = background and, as in the reference, a lone
    pixel at global index 0)."""
    from scipy import ndimage
    C, H, W = bin_mask.shape
    out = np.zeros((C, H, W), np.int64)
    gidx = np.arange(C * H * W, dtype=np.int64).reshape(C, H, W)
    for c in range(C):
        lab, n = ndimage.label(bin_mask[c], structure=np.ones((3, 3), bool))
        if n == 0:
            continue
        mx = ndimage.maximum(gidx[c], labels=lab, index=np.arange(1, n + 1)).astype(np.int64)
        out[c] = np.where(lab > 0, mx[np.maximum(lab, 1) - 1], 0)
    return out


def remove_all_but_one_connected_component(prob: np.ndarray, selection: str) -> np.ndarray:
    """utils/segmentation_preprocessing.py:7-52 on float32 [C,H,W]."""
    lbl = ccl_labels(prob > np.float32(0.5))
    refined = np.zeros_like(prob)
    for c in range(prob.shape[0]):
        comps = np.unique(lbl[c])
        comps = comps[comps != 0]
        if comps.size == 0:
            continue
        areas = np.array([(lbl[c] == k).sum() for k in comps])
        if selection == "largest":
            win = comps[np.argmax(areas)]
        elif selection == "highest_probability":
            # reference: fp32 sum / int64 area -> fp32; the fp64 sum rounded to fp32 is the closest statement of it
            means = np.array([np.float32(np.float32(prob[c][lbl[c] == k].astype(np.float64).sum()) / np.float32(a))
                              for k, a in zip(comps, areas)], np.float32)
            win = comps[np.argmax(means)]
        else:
            raise NotImplementedError(selection)
        refined[c] = lbl[c] == win
    return refined * prob
