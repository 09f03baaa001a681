"""CPU experiment for DESIGN section 6 gap 3: fold LayerNorm into the consumer GEMM,
    LN(x) W^T + b = rstd * (x (gamma*W)^T - mean * rowsum(gamma*W)) + (beta W^T + b),
with the A operand = bf16(x) instead of bf16(LN(x)).  Emulates the bf16 operand roundings of the CUDA encoder on the
oracle's functional ViT and reports the embedding error of (a) the current order of operations and (b) the folded form
against the fp32 oracle.  Test infrastructure (imports the oracle); random-init weights, so the residual-stream statistics
are NOT those of a trained SAM (no checkpoint is available offline): indicative only."""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

from oracle import sam_oracle as O  # noqa: E402


def r16(t):
    return t.bfloat16().float()


def ln_linear(x, g, b, W, bias, mode):
    """LN over the last dim followed by a linear layer, in one of three arithmetic models."""
    D = x.shape[-1]
    if mode == "fp32":
        return F.linear(F.layer_norm(x, (D,), g, b, eps=1e-6), W, bias)
    if mode == "bf16":  # current kernels: bf16(LN(x)) x bf16(W), fp32 accumulate
        return F.linear(r16(F.layer_norm(x, (D,), g, b, eps=1e-6)), r16(W), bias)
    mu = x.mean(-1, keepdim=True)
    rstd = torch.rsqrt((x - mu).pow(2).mean(-1, keepdim=True) + 1e-6)
    Wf = r16(W * g[None, :])
    acc = F.linear(r16(x), Wf)
    return rstd * (acc - mu * Wf.sum(1)[None, :]) + (F.linear(b[None, :], W)[0] + bias)


def block(sd, p, x, heads, window, mode):
    B, H, W_, D = x.shape
    hd = D // heads
    qkv = ln_linear(x, sd[p + "norm1.weight"], sd[p + "norm1.bias"], sd[p + "attn.qkv.weight"], sd[p + "attn.qkv.bias"], mode)
    if mode != "fp32":
        qkv = r16(qkv)
    if window > 0:  # pad AFTER norm1: the pad tokens' qkv is the bias (linear of zeros)
        ph, pw = (-H) % window, (-W_) % window
        pad = sd[p + "attn.qkv.bias"] if mode == "fp32" else r16(sd[p + "attn.qkv.bias"])
        full = pad.expand(B, H + ph, W_ + pw, 3 * D).clone()
        full[:, :H, :W_] = qkv
        Hp, Wp = H + ph, W_ + pw
        y = full.view(B, Hp // window, window, Wp // window, window, 3 * D).permute(0, 1, 3, 2, 4, 5).reshape(-1, window, window, 3 * D)
        S = window
    else:
        y, S = qkv, H
    Bn = y.shape[0]
    q, k, v = y.reshape(Bn, S * S, 3, heads, hd).permute(2, 0, 3, 1, 4).reshape(3, Bn * heads, S * S, hd)
    attn = (q * hd ** -0.5) @ k.transpose(-2, -1) + O._rel_pos_bias(q, sd[p + "attn.rel_pos_h"], sd[p + "attn.rel_pos_w"], S)
    out = (attn.softmax(-1) @ v).view(Bn, heads, S, S, hd).permute(0, 2, 3, 1, 4).reshape(Bn, S, S, D)
    if window > 0:
        out = out.view(B, Hp // window, Wp // window, window, window, D).permute(0, 1, 3, 2, 4, 5).reshape(B, Hp, Wp, D)[:, :H, :W_]
    if mode == "fp32":
        x = x + F.linear(out, sd[p + "attn.proj.weight"], sd[p + "attn.proj.bias"])
    else:
        x = x + F.linear(r16(out), r16(sd[p + "attn.proj.weight"]), sd[p + "attn.proj.bias"])
    h = ln_linear(x, sd[p + "norm2.weight"], sd[p + "norm2.bias"], sd[p + "mlp.lin1.weight"], sd[p + "mlp.lin1.bias"], mode)
    h = F.gelu(h)
    if mode == "fp32":
        return x + F.linear(h, sd[p + "mlp.lin2.weight"], sd[p + "mlp.lin2.bias"])
    return x + F.linear(r16(h), r16(sd[p + "mlp.lin2.weight"]), sd[p + "mlp.lin2.bias"])


@torch.no_grad()
def run(model="vit_b", seed=0):
    torch.set_num_threads(8)
    sd = O.random_state_dict(model, seed=seed)
    cfg = O.VIT_CONFIGS[model]
    img = torch.from_numpy(O.synthetic_radiograph(3)).permute(2, 0, 1).float()
    x0 = O.preprocess(img)[None]
    p = "image_encoder."
    res = {}
    for mode in ("fp32", "bf16", "fold"):
        t = F.conv2d(x0, sd[p + "patch_embed.proj.weight"], sd[p + "patch_embed.proj.bias"], stride=16).permute(0, 2, 3, 1)
        t = t + sd[p + "pos_embed"]
        stats = []
        for i in range(cfg["depth"]):
            win = 0 if i in cfg["global_attn_indexes"] else 14
            stats.append(float((t.mean(-1).abs() / t.std(-1)).mean()))
            t = block(sd, f"{p}blocks.{i}.", t, cfg["num_heads"], win, mode)
        res[mode] = t
        if mode == "fp32":
            print("mean |mu| / sigma of the residual stream per block:", [round(s, 3) for s in stats])
    ref = res["fp32"]
    for mode in ("bf16", "fold"):
        rel = float((res[mode] - ref).norm() / ref.norm())
        print(f"{model} residual stream after {cfg['depth']} blocks, {mode:5s} vs fp32: rel-L2 {rel:.3e}")


if __name__ == "__main__":
    run(sys.argv[1] if len(sys.argv) > 1 else "vit_b")
