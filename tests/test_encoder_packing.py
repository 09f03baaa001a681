"""CPU: the LayerNorm-folding weight preparation of the encoder engine (image_encoder._pack) reproduces
LN(x) W^T + b as rstd * (x (gamma W)^T - mean * colsum) + (beta W^T + b) (csrc/gemm_epilogue.cuh)."""
import torch
import torch.nn.functional as F

from samcarriestheburden_b200.segment_anything.modeling.image_encoder import _pack


def test_fold_packing_matches_layernorm_linear():
    g = torch.Generator().manual_seed(0)
    D, N = 256, 384
    sd = {"blk.norm1.weight": 1 + 0.3 * torch.randn(D, generator=g), "blk.norm1.bias": 0.2 * torch.randn(D, generator=g),
          "blk.attn.qkv.weight": torch.randn((N, D), generator=g) / D ** 0.5, "blk.attn.qkv.bias": torch.randn(N, generator=g)}
    cache = {}
    Wf = _pack(sd, "blk.attn.qkv|fold_w|blk.norm1", torch.float32, "cpu", cache)   # fp32 "operand": exact algebra
    s = _pack(sd, "blk.attn.qkv|fold_s|blk.norm1", torch.float32, "cpu", cache)
    c = _pack(sd, "blk.attn.qkv|fold_c|blk.norm1", torch.float32, "cpu", cache)
    assert len(cache) == 1 and Wf.shape == (N, D) and s.shape == (N,) and c.shape == (N,)
    x = 2.0 * torch.randn((64, D), generator=g) + 0.5
    mean = x.mean(-1, keepdim=True)
    rstd = torch.rsqrt(x.var(-1, unbiased=False, keepdim=True) + 1e-6)
    got = rstd * (x @ Wf.T - mean * s[None, :]) + c[None, :]
    ref = F.linear(F.layer_norm(x, (D,), sd["blk.norm1.weight"], sd["blk.norm1.bias"], eps=1e-6),
                   sd["blk.attn.qkv.weight"], sd["blk.attn.qkv.bias"])
    assert torch.allclose(got, ref, atol=2e-5, rtol=1e-5), float((got - ref).abs().max())
    # 16-bit operands: the column sums are those of the ROUNDED weights
    Wh = _pack(sd, "blk.attn.qkv|fold_w|blk.norm1", torch.float16, "cpu", {})
    sh = _pack(sd, "blk.attn.qkv|fold_s|blk.norm1", torch.float16, "cpu", {})
    assert Wh.dtype == torch.float16 and torch.equal(sh, Wh.double().sum(1).float())
