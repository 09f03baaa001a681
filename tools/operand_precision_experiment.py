"""CPU study for VERDICT r1 item 1: which operand rounding lets CUDA-encoder -> refine masks reach Dice >= 0.999?

Emulates the operand roundings of the CUDA encoder (every GEMM / attention operand rounded to a 16-bit format, fp32
accumulate, fp32 residual stream / LN / softmax statistics) on the oracle's functional ViT, then runs the ORACLE's
two-pass refine on the perturbed embedding and on the fp32 embedding and reports embedding rel-L2, per-class Dice
(utils/dice_coefficient.py:30-53 definition) and the mismatched native pixels.

    python tools/operand_precision_experiment.py [vit_b] [formats...]     formats: bf16 fp16 bf16w (weights only) ...

Test infrastructure (imports the oracle).  Random-init weights."""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

from oracle import sam_oracle as O  # noqa: E402


def rounder(fmt):
    if fmt == "fp32":
        return lambda t: t
    if fmt == "bf16":
        return lambda t: t.bfloat16().float()
    if fmt == "fp16":
        return lambda t: t.half().float()
    raise ValueError(fmt)


class Cfg:
    """which tensors are rounded, and to what"""

    def __init__(self, act="fp32", wgt="fp32", p="fp32", fold=False, res="fp32"):
        self.a, self.w, self.p, self.fold = rounder(act), rounder(wgt), rounder(p), fold
        self.r = f24_round if res == "f24" else (lambda t: t)


def f24_round(x):
    """Residual stream stored as fp16 hi + 8-bit extension of the mantissa (csrc/gemm_epilogue.cuh, `f24`): x ~ hi + q * ulp(hi) / 256,
    q = round-to-nearest of the remainder in 1/256 ulp, clamped to [-128, 127]."""
    h = x.half().float()
    e = torch.floor(torch.log2(h.abs().clamp_min(2.0 ** -14)))   # exponent of hi (subnormals share 2^-14)
    step = torch.exp2(e - 10 - 8)                                   # ulp(hi) / 256
    q = torch.clamp(torch.round((x - h) / step), -128, 127)
    return h + q * step


def ln_linear(x, g, b, W, bias, c):
    """LN + linear.  fold: the LayerNorm-folded form of the CUDA encoder (A operand = round16(x), weights
    round16(gamma * W), epilogue rstd * (acc - mu * colsum) + (beta W^T + bias), statistics from sum / sum of squares)."""
    D = x.shape[-1]
    if not c.fold:
        return F.linear(c.a(F.layer_norm(x, (D,), g, b, eps=1e-6)), c.w(W), bias)
    s1, s2 = x.sum(-1, keepdim=True), (x * x).sum(-1, keepdim=True)
    mu = s1 / D
    rstd = torch.rsqrt(s2 / D - mu * mu + 1e-6)
    Wf = c.w(W * g[None, :])
    acc = F.linear(c.a(x), Wf)
    return acc * rstd + ((-rstd * mu) * Wf.sum(1)[None, :] + (F.linear(b[None, :], W)[0] + bias))


def block(sd, p, x, heads, window, c):
    B, H, W_, D = x.shape
    hd = D // heads
    qkv = c.a(ln_linear(x, sd[p + "norm1.weight"], sd[p + "norm1.bias"], sd[p + "attn.qkv.weight"], sd[p + "attn.qkv.bias"], c))
    if window > 0:
        ph, pw = (-H) % window, (-W_) % window
        full = c.a(sd[p + "attn.qkv.bias"]).expand(B, H + ph, W_ + pw, 3 * D).clone()
        full[:, :H, :W_] = qkv
        Hp, Wp = H + ph, W_ + pw
        y = full.view(B, Hp // window, window, Wp // window, window, 3 * D).permute(0, 1, 3, 2, 4, 5).reshape(-1, window, window, 3 * D)
        S = window
    else:
        y, S = qkv, H
    Bn = y.shape[0]
    q, k, v = y.reshape(Bn, S * S, 3, heads, hd).permute(2, 0, 3, 1, 4).reshape(3, Bn * heads, S * S, hd)
    attn = (q * hd ** -0.5) @ k.transpose(-2, -1) + O._rel_pos_bias(q, c.w(sd[p + "attn.rel_pos_h"]), c.w(sd[p + "attn.rel_pos_w"]), S)
    m = attn.amax(-1, keepdim=True)
    e = torch.exp(attn - m)
    out = (c.p(e) @ v) / e.sum(-1, keepdim=True)   # unnormalised P rounded as the MMA operand, fp32 row sum
    out = out.view(Bn, heads, S, S, hd).permute(0, 2, 3, 1, 4).reshape(Bn, S, S, D)
    if window > 0:
        out = out.view(B, Hp // window, Wp // window, window, window, D).permute(0, 1, 3, 2, 4, 5).reshape(B, Hp, Wp, D)[:, :H, :W_]
    x = c.r(x + F.linear(c.a(out), c.w(sd[p + "attn.proj.weight"]), sd[p + "attn.proj.bias"]))
    h = F.gelu(ln_linear(x, sd[p + "norm2.weight"], sd[p + "norm2.bias"], sd[p + "mlp.lin1.weight"], sd[p + "mlp.lin1.bias"], c))
    return c.r(x + F.linear(c.a(h), c.w(sd[p + "mlp.lin2.weight"]), sd[p + "mlp.lin2.bias"]))


@torch.no_grad()
def encoder(sd, cfg, x0, c):
    p = "image_encoder."
    D = cfg["embed_dim"]
    patches = F.unfold(x0, 16, stride=16).transpose(1, 2)  # [B, 4096, 768]
    t = F.linear(c.a(patches), c.w(sd[p + "patch_embed.proj.weight"].reshape(D, -1)), sd[p + "patch_embed.proj.bias"])
    t = c.r(t.view(-1, 64, 64, D) + sd[p + "pos_embed"])
    for i in range(cfg["depth"]):
        win = 0 if i in cfg["global_attn_indexes"] else 14
        t = block(sd, f"{p}blocks.{i}.", t, cfg["num_heads"], win, c)
    t = c.a(t).permute(0, 3, 1, 2)
    t = F.conv2d(t, c.w(sd[p + "neck.0.weight"]))
    t = O.layer_norm_2d(t, sd[p + "neck.1.weight"], sd[p + "neck.1.bias"])
    t = F.conv2d(c.a(t), c.w(sd[p + "neck.2.weight"]), padding=1)
    return O.layer_norm_2d(t, sd[p + "neck.3.weight"], sd[p + "neck.3.bias"])


def dice_rows(a, b):
    out = []
    for x, y in zip(a, b):
        s = x.sum() + y.sum()
        out.append(float("nan") if s == 0 else 2.0 * float((x & y).sum()) / float(s))
    return out


@torch.no_grad()
def run(model="vit_b", variants=None, seeds=(3,), native=(1182, 754)):
    torch.set_num_threads(8)
    sd = O.random_state_dict(model, seed=0)
    cfg = O.VIT_CONFIGS[model]
    variants = variants or ["bf16", "fp16"]
    table = {
        "bf16": Cfg("bf16", "bf16", "bf16"), "fp16": Cfg("fp16", "fp16", "fp16"),
        "bf16w": Cfg("fp32", "bf16", "fp32"), "bf16a": Cfg("bf16", "fp32", "bf16"),
        "fp16a_bf16w": Cfg("fp16", "bf16", "fp16"),
        "fp16fold": Cfg("fp16", "fp16", "fp16", fold=True), "bf16fold": Cfg("bf16", "bf16", "bf16", fold=True),
        "fp16fold_f24": Cfg("fp16", "fp16", "fp16", fold=True, res="f24"),
        "f24only": Cfg("fp32", "fp32", "fp32", fold=False, res="f24"),
    }
    for seed in seeds:
        img = torch.from_numpy(O.synthetic_radiograph(seed)).permute(2, 0, 1).float()
        x0 = O.preprocess(img)[None]
        ref = encoder(sd, cfg, x0, Cfg())
        seg = O.synthetic_unet_masks(seed)
        seg_r, est_r, nat_r, low_r = O.refine(sd, ref, seg, (1024, 1024), native)
        for name in variants:
            emb = encoder(sd, cfg, x0, table[name])
            rel = float((emb - ref).norm() / ref.norm())
            seg_t, est_t, nat_t, low_t = O.refine(sd, emb, seg, (1024, 1024), native)
            ks = sorted(nat_r)
            d = dice_rows([nat_r[k] for k in ks], [nat_t[k] for k in ks])
            mism = sum(int((nat_r[k] != nat_t[k]).sum()) for k in ks)
            tot = sum(nat_r[k].size for k in ks)
            lowerr = max(float(np.abs(low_r[k][1] - low_t[k][1]).max()) for k in ks)
            print(f"{model} seed {seed} {name:12s} emb rel-L2 {rel:.3e}  min Dice {np.nanmin(d):.5f}  mean Dice {np.nanmean(d):.5f}  "
                  f"mismatched px {mism} / {tot} ({mism / tot:.2e})  max|dlow| {lowerr:.2e}", flush=True)


if __name__ == "__main__":
    a = sys.argv[1:]
    run(a[0] if a else "vit_b", a[1:] or None)
