"""Op-level parity of the CUDA kernels (through the C ABI) against plain torch fp32 restatements."""
import ctypes as C

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import sam_oracle as O
from samcarriestheburden_b200 import _lib

pytestmark = pytest.mark.gpu
DEV = "cuda"


FMT = {"bf16": (torch.bfloat16, 0), "fp16": (torch.float16, 1)}  # name -> (torch dtype, b200sam operand_format)


def _gemm(A, W, bias=None, residual=None, res_row_mod=0, gelu=False, out_bf16=True, max_ctas=0):
    """out_bf16: 16-bit output in the operands' format (bf16 or fp16, taken from A.dtype), else fp32."""
    lib = _lib.load()
    M, K = A.shape
    N = W.shape[0]
    assert A.dtype == W.dtype and A.dtype in (torch.bfloat16, torch.float16)
    fn = lib.b200sam_gemm_f16 if A.dtype == torch.float16 else lib.b200sam_gemm_bf16
    out = torch.empty((M, N), dtype=A.dtype if out_bf16 else torch.float32, device=DEV)
    ldr = residual.shape[1] if residual is not None else 0
    _lib.check(fn(A.data_ptr(), W.data_ptr(), out.data_ptr(), _lib.ptr(bias), _lib.ptr(residual), M,
                  N, K, A.stride(0), W.stride(0), N, ldr, res_row_mod, int(gelu), int(out_bf16),
                  max_ctas, _lib.current_stream()), "gemm")
    torch.cuda.synchronize()
    return out


@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (256, 512, 128), (4096, 3840, 1280), (4096, 1280, 5120),
                                   (200, 384, 192), (4096, 768, 768), (1000, 264, 72),
                                   # N <= 128: tall 256 x 128 tiles (two accumulators per weight k-block), ragged M
                                   (300, 128, 192), (513, 64, 64), (70000, 128, 768), (131, 8, 64)])
@pytest.mark.parametrize("fmt", ["bf16", "fp16"])
def test_gemm_bf16_bias(M, N, K, fmt):
    g = torch.Generator(device="cpu").manual_seed(M + N + K)
    dt = FMT[fmt][0]
    A = torch.randn((M, K), generator=g).to(DEV).to(dt)
    W = (torch.randn((N, K), generator=g) / K ** 0.5).to(DEV).to(dt)
    b = torch.randn((N,), generator=g).to(DEV)
    ref = A.float() @ W.float().T + b
    out = _gemm(A, W, b)
    err = (out.float() - ref).abs().max().item()
    # the 16-bit output rounding dominates: 2^-8 relative for bf16, 2^-11 for fp16
    assert err <= (2e-2 if fmt == "bf16" else 2.5e-3) * ref.abs().max().item() + 1e-3, err
    # fp32 output isolates accumulation error (inputs are exact bf16): tight tolerance
    out32 = _gemm(A, W, b, out_bf16=False)
    assert torch.allclose(out32, ref, atol=2e-3, rtol=2e-3), (out32 - ref).abs().max().item()


def test_gemm_epilogues():
    g = torch.Generator(device="cpu").manual_seed(1)
    M, N, K = 8192, 1280, 1280
    A = torch.randn((M, K), generator=g).to(DEV).bfloat16()
    W = (torch.randn((N, K), generator=g) / K ** 0.5).to(DEV).bfloat16()
    b = torch.randn((N,), generator=g).to(DEV)
    lin = A.float() @ W.float().T + b
    out = _gemm(A, W, b, gelu=True)
    assert torch.allclose(out.float(), F.gelu(lin), atol=3e-2, rtol=2e-2)
    res = torch.randn((M, N), generator=g).to(DEV)
    got = res.clone()  # in-place residual, as used for the residual stream
    lib = _lib.load()
    _lib.check(lib.b200sam_gemm_bf16(A.data_ptr(), W.data_ptr(), got.data_ptr(), b.data_ptr(), got.data_ptr(), M, N, K,
                                     K, K, N, N, 0, 0, 0, 0, _lib.current_stream()))
    torch.cuda.synchronize()
    assert torch.allclose(got, lin + res, atol=2e-3, rtol=2e-3)
    pos = torch.randn((4096, N), generator=g).to(DEV)  # broadcast table (pos_embed) over 2 images
    out = _gemm(A, W, b, residual=pos, res_row_mod=4096, out_bf16=False)
    assert torch.allclose(out, lin + pos.repeat(2, 1), atol=2e-3, rtol=2e-3)
    # a persistent grid smaller than the tile count must give identical results
    assert torch.equal(_gemm(A, W, b, max_ctas=7), _gemm(A, W, b))
    # fp16 operands: same epilogues, finite saturation instead of inf on overflow
    Ah, Wh = A.half(), W.half()
    linh = Ah.float() @ Wh.float().T + b
    assert torch.allclose(_gemm(Ah, Wh, b, gelu=True).float(), F.gelu(linh), atol=4e-3, rtol=2e-3)
    big = _gemm((Ah * 300).half(), (Wh * 300).half(), None)
    assert bool(torch.isfinite(big.float()).all()) and float(big.float().abs().max()) == 65504.0


@pytest.fixture(params=["single", "pair"])
def gemm_kernel(request):
    """Run a test on both GEMM kernels: single-CTA and CTA pair (tcgen05 cta_group::2)."""
    lib = _lib.load()
    lib.b200sam_set_gemm_pair({"single": 0, "pair": 1}[request.param])
    yield request.param
    lib.b200sam_set_gemm_pair(-1)


@pytest.mark.parametrize("M,N,K", [(256, 256, 64), (4096, 3840, 1280), (8192, 1280, 5120), (1000, 264, 72), (131, 384, 192),
                                   (300, 136, 64), (32768, 1280, 1280)])
def test_gemm_pair_kernel_is_bit_identical_to_single(M, N, K):
    """The CTA-pair kernel accumulates every output element over the same K sequence as the single-CTA kernel and shares
    its epilogue: the results must be bit-identical (16-bit + GELU, fp32 + residual, ragged M / N tails)."""
    lib = _lib.load()
    assert lib.b200sam_gemm_pair_max_clusters() >= 32
    g = torch.Generator(device="cpu").manual_seed(M + N + K)
    A = torch.randn((M, K), generator=g).to(DEV).half()
    W = (torch.randn((N, K), generator=g) / K ** 0.5).to(DEV).half()
    b = torch.randn((N,), generator=g).to(DEV)
    res = torch.randn((M, N), generator=g).to(DEV)
    outs = {}
    try:
        for mode in (0, 1):  # single-CTA, CTA pair
            lib.b200sam_set_gemm_pair(mode)
            outs[mode] = (_gemm(A, W, b, gelu=True), _gemm(A, W, b, residual=res, out_bf16=False))
    finally:
        lib.b200sam_set_gemm_pair(-1)
    assert torch.equal(outs[0][0], outs[1][0])
    assert torch.equal(outs[0][1], outs[1][1])
    ref = A.float() @ W.float().T + b
    assert torch.allclose(outs[1][1], ref + res, atol=2e-3, rtol=2e-3)


@pytest.mark.parametrize("fmt", ["bf16", "fp16"])
@pytest.mark.parametrize("M,D,N", [(8192, 1280, 3840), (4096, 768, 3072), (1000, 1024, 1024)])
def test_gemm_layernorm_folding(fmt, M, D, N, gemm_kernel):
    """The two halves of a LayerNorm folded into the GEMMs around it (Block.forward, image_encoder.py:166-182):
    producer x = A W^T + b + residual with the 16-bit copy of x and per-row partial sums, consumer
    y = act(LN(x) W2^T + b2) computed as rstd * (x16 (gamma * W2)^T - mean * colsum) + (beta W2^T + b2)."""
    lib = _lib.load()
    dt, of = FMT[fmt]
    g = torch.Generator(device="cpu").manual_seed(M + D + N)
    A = torch.randn((M, D), generator=g).to(DEV).to(dt)
    W = (torch.randn((D, D), generator=g) / D ** 0.5).to(DEV).to(dt)
    b = torch.randn((D,), generator=g).to(DEV)
    res = (2.0 * torch.randn((M, D), generator=g) + 0.7).to(DEV)  # mean / sigma ~ 0.3
    x_ref = A.float() @ W.float().T + b + res
    x = res.clone()
    x16 = torch.empty((M, D), dtype=dt, device=DEV)
    nparts = D // 64
    stat = torch.full((M, nparts, 2), float("nan"), device=DEV)
    _lib.check(lib.b200sam_gemm_ln_residual(A.data_ptr(), W.data_ptr(), b.data_ptr(), x.data_ptr(), x.data_ptr(),
                                            x16.data_ptr(), stat.data_ptr(), M, D, D, of, _lib.current_stream()))
    torch.cuda.synchronize()
    assert torch.allclose(x, x_ref, atol=2e-3, rtol=2e-3)
    assert torch.equal(x16, x.to(dt))
    sums = stat.sum(1)
    assert torch.allclose(sums[:, 0], x.sum(1), rtol=1e-4, atol=1e-2)
    assert torch.allclose(sums[:, 1], (x * x).sum(1), rtol=1e-4, atol=1e-2)
    gamma = (1.0 + 0.2 * torch.randn((D,), generator=g)).to(DEV)
    beta = (0.2 * torch.randn((D,), generator=g)).to(DEV)
    W2 = (torch.randn((N, D), generator=g) / D ** 0.5).to(DEV)
    b2 = torch.randn((N,), generator=g).to(DEV)
    Wf = (W2 * gamma[None, :]).to(dt).contiguous()
    colsum = Wf.double().sum(1).float().contiguous()
    cfold = (W2.double() @ beta.double() + b2.double()).float().contiguous()
    for gelu in (0, 1):
        y = torch.empty((M, N), dtype=dt, device=DEV)
        _lib.check(lib.b200sam_gemm_ln_folded(x16.data_ptr(), Wf.data_ptr(), cfold.data_ptr(), colsum.data_ptr(),
                                              stat.data_ptr(), nparts, 1e-6, y.data_ptr(), M, N, D, gelu, of,
                                              _lib.current_stream()))
        torch.cuda.synchronize()
        ref = F.layer_norm(x, (D,), gamma, beta, eps=1e-6) @ W2.T + b2
        ref = F.gelu(ref) if gelu else ref
        err = (y.float() - ref).abs()
        tol = 6e-2 if fmt == "bf16" else 8e-3
        assert err.max().item() < tol and err.mean().item() < tol / 8, (fmt, gelu, err.max().item(), err.mean().item())


@pytest.mark.parametrize("D", [256, 768, 1024, 1280])
def test_layernorm(D):
    g = torch.Generator(device="cpu").manual_seed(D)
    x = (torch.randn((777, D), generator=g) * 3 + 1.5).to(DEV)
    w, b = torch.randn((D,), generator=g).to(DEV), torch.randn((D,), generator=g).to(DEV)
    lib = _lib.load()
    ref = F.layer_norm(x, (D,), w, b, eps=1e-6)
    y32 = torch.empty_like(x)
    _lib.check(lib.b200sam_layernorm(x.data_ptr(), w.data_ptr(), b.data_ptr(), 1e-6, 777, D, y32.data_ptr(), 0,
                                     _lib.current_stream()))
    y16 = torch.empty((777, D), dtype=torch.bfloat16, device=DEV)
    _lib.check(lib.b200sam_layernorm(x.data_ptr(), w.data_ptr(), b.data_ptr(), 1e-6, 777, D, y16.data_ptr(), 1,
                                     _lib.current_stream()))
    yh = torch.empty((777, D), dtype=torch.float16, device=DEV)
    _lib.check(lib.b200sam_layernorm(x.data_ptr(), w.data_ptr(), b.data_ptr(), 1e-6, 777, D, yh.data_ptr(), 2,
                                     _lib.current_stream()))
    torch.cuda.synchronize()
    assert torch.allclose(y32, ref, atol=1e-5, rtol=1e-5)
    assert torch.equal(y16, ref.bfloat16()) or (y16.float() - ref).abs().max() < 4e-2
    assert torch.equal(yh, ref.half()) or (yh.float() - ref).abs().max() < 5e-3


def _attention_reference(qkv, bias16, rel_h, rel_w, heads, hd, window):
    """fp32 restatement of Attention.forward + window partition on bf16-rounded inputs (image_encoder.py:166-240)."""
    D = heads * hd
    B = qkv.shape[0] // 4096
    x = qkv.float().view(B, 64, 64, 3 * D)
    if window:
        pad = (-64) % window
        grid = bias16.float().view(1, 1, 1, 3 * D).expand(B, 64 + pad, 64 + pad, 3 * D).clone()
        grid[:, :64, :64] = x
        S = window
        nw = (64 + pad) // window
        t = grid.view(B, nw, S, nw, S, 3 * D).permute(0, 1, 3, 2, 4, 5).reshape(-1, S * S, 3, heads, hd)
    else:
        S = 64
        t = x.reshape(B, S * S, 3, heads, hd)
    t = t.permute(2, 0, 3, 1, 4).reshape(3, -1, S * S, hd)
    q, k, v = t[0], t[1], t[2]
    attn = (q * hd ** -0.5) @ k.transpose(-2, -1) + O._rel_pos_bias(q, rel_h.float(), rel_w.float(), S)
    o = attn.softmax(-1) @ v
    o = o.view(-1, heads, S, S, hd).permute(0, 2, 3, 1, 4).reshape(-1, S, S, D)
    if window:
        o = o.view(B, nw, nw, S, S, D).permute(0, 1, 3, 2, 4, 5).reshape(B, nw * S, nw * S, D)[:, :64, :64]
    return o.reshape(B * 4096, D)


@pytest.mark.parametrize("fmt", ["bf16", "fp16"])
@pytest.mark.parametrize("heads,hd,glob", [(16, 80, 0), (12, 64, 0), (4, 80, 1), (3, 64, 1)])
def test_encoder_attention(heads, hd, glob, fmt):
    g = torch.Generator(device="cpu").manual_seed(heads * 100 + hd + glob)
    dt, of = FMT[fmt]
    D = heads * hd
    B = 2
    S = 64 if glob == 1 else 14
    qkv = torch.randn((B * 4096, 3 * D), generator=g).to(DEV).to(dt)
    bias = torch.randn((3 * D,), generator=g).to(DEV).to(dt)
    rel_h = (0.3 * torch.randn((2 * S - 1, hd), generator=g)).to(DEV).to(dt)
    rel_w = (0.3 * torch.randn((2 * S - 1, hd), generator=g)).to(DEV).to(dt)
    out = torch.empty((B * 4096, D), dtype=dt, device=DEV)
    lib = _lib.load()
    _lib.check(lib.b200sam_encoder_attention(qkv.data_ptr(), bias.data_ptr(), rel_h.data_ptr(), rel_w.data_ptr(),
                                             out.data_ptr(), B, heads, hd, glob, of, _lib.current_stream()))
    torch.cuda.synchronize()
    ref = _attention_reference(qkv, bias, rel_h, rel_w, heads, hd, 0 if glob == 1 else 14)
    err = (out.float() - ref).abs()
    # outputs are O(1): the bounds are the P / output roundings of the format (bf16 2^-8, fp16 2^-11) with margin
    # (measured: bf16 max 1.6e-2 / mean 1.3e-3)
    tmax, tmean = (2.5e-2, 2.5e-3) if fmt == "bf16" else (5e-3, 5e-4)
    print(f"attention {fmt} heads={heads} hd={hd} glob={glob}: max {err.max().item():.2e} mean {err.mean().item():.2e}")
    assert err.max().item() < tmax, (err.max().item(), err.mean().item())
    assert err.mean().item() < tmean


@pytest.mark.parametrize("heads,B", [(1, 1), (2, 3), (16, 1)])
def test_windowed_attention_persistent_grid_shapes(heads, B):
    """The windowed kernel is persistent (2 CTAs per SM walk over the 25 * heads * B windows): fewer items than CTAs (25, 150),
    and a count that is not a multiple of the grid (400 on 296 CTAs: some CTAs take two windows, most one)."""
    hd, S = 64, 14
    g = torch.Generator(device="cpu").manual_seed(7 + heads + B)
    D = heads * hd
    qkv = torch.randn((B * 4096, 3 * D), generator=g).to(DEV).half()
    bias = torch.randn((3 * D,), generator=g).to(DEV).half()
    rel_h = (0.3 * torch.randn((2 * S - 1, hd), generator=g)).to(DEV).half()
    rel_w = (0.3 * torch.randn((2 * S - 1, hd), generator=g)).to(DEV).half()
    out = torch.full((B * 4096, D), float("nan"), dtype=torch.float16, device=DEV)
    lib = _lib.load()
    for _ in range(2):  # twice: the second launch starts from whatever the first left in shared memory / TMEM
        _lib.check(lib.b200sam_encoder_attention(qkv.data_ptr(), bias.data_ptr(), rel_h.data_ptr(), rel_w.data_ptr(),
                                                 out.data_ptr(), B, heads, hd, 0, 1, _lib.current_stream()))
    torch.cuda.synchronize()
    ref = _attention_reference(qkv, bias, rel_h, rel_w, heads, hd, 14)
    err = (out.float() - ref).abs()
    assert not torch.isnan(out).any()
    assert err.max().item() < 5e-3 and err.mean().item() < 5e-4, (err.max().item(), err.mean().item())


def test_encoder_attention_rejects_unknown_modes():
    lib = _lib.load()
    t = torch.zeros((4096, 3 * 64), dtype=torch.float16, device=DEV)
    o = torch.zeros((4096, 64), dtype=torch.float16, device=DEV)
    for glob, of in ((2, 1), (0, 7)):
        rc = lib.b200sam_encoder_attention(t.data_ptr(), t.data_ptr(), t.data_ptr(), t.data_ptr(), o.data_ptr(), 1, 1, 64,
                                           glob, of, _lib.current_stream())
        assert rc != 0 and lib.b200sam_last_error()


@pytest.mark.parametrize("seed", range(4))
def test_prompt_extraction_bit_exact(seed):
    from samcarriestheburden_b200.segment_anything.utils.prompt_utils import extract_seeds_boxes
    masks = np.stack([O.synthetic_unet_masks(10 * seed + i) for i in range(3)])
    if seed == 1:
        masks[0, 3] = masks[0, 5]
    if seed == 2:
        rng = np.random.default_rng(5)
        masks &= rng.random(masks.shape) < 0.5
    if seed == 3:
        masks[1] = False  # image without any class
    seeds, boxes, has_seed, has_box = extract_seeds_boxes(torch.from_numpy(masks).to(DEV))
    torch.cuda.synchronize()
    for i in range(masks.shape[0]):
        s, hs, b, hb = O.extract_seeds_boxes(masks[i])
        assert np.array_equal(has_seed[i].cpu().numpy().astype(bool), hs)
        assert np.array_equal(has_box[i].cpu().numpy().astype(bool), hb)
        assert np.array_equal(seeds[i].cpu().numpy()[hs], s[hs])
        assert np.array_equal(boxes[i].cpu().numpy()[hb], b[hb])


def test_prompt_extraction_ragged_shapes():
    from samcarriestheburden_b200.segment_anything.utils.prompt_utils import extract_seeds_boxes
    rng = np.random.default_rng(0)
    # (40, 48, 64) / (64, 32, 32): the 16-pixel fast path with more classes than one load batch and than 32 (two masks)
    for (C_, H, W) in [(1, 1, 1), (5, 37, 53), (17, 384, 224), (3, 1, 1000), (64, 33, 31), (40, 48, 64), (64, 32, 32)]:
        masks = rng.random((2, C_, H, W)) < 0.3
        seeds, boxes, has_seed, has_box = extract_seeds_boxes(torch.from_numpy(masks).to(DEV))
        for i in range(2):
            s, hs, b, hb = O.extract_seeds_boxes(masks[i])
            assert np.array_equal(has_seed[i].cpu().numpy().astype(bool), hs)
            assert np.array_equal(seeds[i].cpu().numpy()[hs], s[hs])
            assert np.array_equal(boxes[i].cpu().numpy()[hb], b[hb])


def test_prompt_extraction_nonbool_bytes_and_big_batch():
    """C ABI contract: any non-zero byte is a set pixel (the fast path adds 0/1 bytes and recounts otherwise);
    and a batch large enough that every CTA of the persistent grid walks several (image, chunk) items."""
    from samcarriestheburden_b200.segment_anything.utils.prompt_utils import extract_seeds_boxes
    lib = _lib.load()
    masks = np.stack([O.synthetic_unet_masks(60 + i) for i in range(40)])
    t = torch.from_numpy(masks).to(DEV)
    seeds, boxes, has_seed, has_box = extract_seeds_boxes(t)
    for i in (0, 17, 39):
        s, hs, b, hb = O.extract_seeds_boxes(masks[i])
        assert np.array_equal(has_seed[i].cpu().numpy().astype(bool), hs)
        assert np.array_equal(seeds[i].cpu().numpy()[hs], s[hs])
        assert np.array_equal(boxes[i].cpu().numpy()[hb], b[hb])
    rng = np.random.default_rng(1)
    raw = (masks[:3].astype(np.uint8) * rng.integers(1, 256, size=masks[:3].shape).astype(np.uint8))
    u8 = torch.from_numpy(raw).to(DEV)
    N, Cn, H, W = u8.shape
    s2 = torch.empty((N, Cn, 2), dtype=torch.int32, device=DEV)
    b2 = torch.empty((N, Cn, 4), dtype=torch.int32, device=DEV)
    hs2 = torch.empty((N, Cn), dtype=torch.uint8, device=DEV)
    hb2 = torch.empty((N, Cn), dtype=torch.uint8, device=DEV)
    scratch = torch.empty(lib.b200sam_prompt_extract_scratch_bytes(N, Cn) // 8 + 1, dtype=torch.int64, device=DEV)
    _lib.check(lib.b200sam_prompt_extract(u8.data_ptr(), N, Cn, H, W, s2.data_ptr(), b2.data_ptr(), hs2.data_ptr(),
                                          hb2.data_ptr(), scratch.data_ptr(), _lib.current_stream()))
    torch.cuda.synchronize()
    assert torch.equal(s2, seeds[:3]) and torch.equal(b2, boxes[:3])
    assert torch.equal(hs2, has_seed[:3]) and torch.equal(hb2, has_box[:3])


@pytest.mark.parametrize("orig", [(1024, 1024), (1182, 754), (578, 881), (2570, 2040), (301, 299)])
def test_upscale_threshold(orig):
    from samcarriestheburden_b200.segment_anything.modeling.sam import upscale_masks
    g = torch.Generator(device="cpu").manual_seed(orig[0])
    low = torch.randn((3, 1, 256, 256), generator=g)
    inp = O.get_preprocess_shape(*orig)
    ref = O.postprocess_masks(low, inp, orig)
    mask, small = upscale_masks(low.to(DEV), inp, orig, small_size=(384, 224))
    logits = upscale_masks(low.to(DEV), inp, orig, return_logits=True)
    torch.cuda.synchronize()
    assert (logits.cpu() - ref).abs().max().item() < 1e-5
    ref_mask = ref > 0
    mism = int((mask.cpu() != ref_mask).sum())
    inter = float((mask.cpu() & ref_mask).sum())
    dice = 2 * inter / float(mask.sum().cpu() + ref_mask.sum())
    # the mask path evaluates the same bilinear composition in its separable three-tap form (rounding differs by
    # ~1 ulp of the value): only pixels whose logit is within the restatement's own 2e-6 band of 0 may flip
    flipped = mask.cpu() != ref_mask
    band = float(ref.abs()[flipped].max()) if mism else 0.0
    print(f"upscale {orig}: {mism} flipped px of {ref_mask.numel()}, max |logit| at a flip {band:.2e}")
    assert mism <= 6 and band < 2e-6 and dice >= 0.9999, (mism, band, dice)
    ref_small = F.interpolate(ref_mask.float(), size=(384, 224), mode="nearest-exact") > 0.5
    assert int((small.cpu() != ref_small).sum()) <= 2
    # the nearest-exact tap is taken from the native mask the same launch produced
    own_small = F.interpolate(mask.float().cpu(), size=(384, 224), mode="nearest-exact") > 0.5
    assert torch.equal(small.cpu(), own_small)


def test_upscale_nonzero_threshold_and_exact_zero_logits():
    from samcarriestheburden_b200.segment_anything.modeling.sam import upscale_masks
    g = torch.Generator(device="cpu").manual_seed(4)
    low = torch.randn((2, 1, 256, 256), generator=g)
    low[1, 0, 100:140] = 0.0      # exact zeros: 0 > 0 is False
    low[1, 0, 140:160] = -0.0
    for orig in [(1024, 1024), (200, 120)]:
        inp = O.get_preprocess_shape(*orig)
        ref = O.postprocess_masks(low, inp, orig)
        for thr in (0.0, 0.25):
            m, small = upscale_masks(low.to(DEV), inp, orig, threshold=thr, small_size=(96, 56))
            torch.cuda.synchronize()
            refm = ref > thr
            flipped = m.cpu() != refm
            band = float((ref - thr).abs()[flipped].max()) if bool(flipped.any()) else 0.0
            assert int(flipped.sum()) <= 6 and band < 2e-6, (orig, thr, int(flipped.sum()), band)
            own = F.interpolate(m.float().cpu(), size=(96, 56), mode="nearest-exact") > 0.5
            assert torch.equal(small.cpu(), own), (orig, thr)
    zero_rows = upscale_masks(low[1:].to(DEV), (1024, 1024), (1024, 1024))[0, 0, 410:550]
    assert not bool(zero_rows.any())


def test_upscale_mask_only_small_and_odd_pitch():
    """mask-only / small-only calls and output pitches that are not multiples of 8 (staged write-out path)."""
    from samcarriestheburden_b200.segment_anything.modeling.sam import upscale_masks
    g = torch.Generator(device="cpu").manual_seed(9)
    for orig in [(57, 1027), (1025, 8 * 131 + 5), (33, 9)]:
        low = torch.randn((2, 1, 256, 256), generator=g)
        inp = O.get_preprocess_shape(*orig)
        ref_mask = O.postprocess_masks(low, inp, orig) > 0
        canary = torch.full((2 * orig[0] * orig[1] + 64,), 7, dtype=torch.uint8, device=DEV)
        mask = upscale_masks(low.to(DEV), inp, orig)
        torch.cuda.synchronize()
        assert int((mask.cpu() != ref_mask).sum()) <= 2, orig
        assert bool((canary == 7).all())


def test_linear_f32():
    g = torch.Generator(device="cpu").manual_seed(3)
    lib = _lib.load()
    for (M, N, K, act) in [(35, 256, 256, 0), (4096, 128, 256, 0), (391, 2048, 256, 1), (391, 256, 2048, 0),
                           (16384, 128, 64, 2)]:
        A = torch.randn((M, K), generator=g).to(DEV)
        A2 = torch.randn((64, K), generator=g).to(DEV)
        W = (torch.randn((N, K), generator=g) / K ** 0.5).to(DEV)
        b = torch.randn((N,), generator=g).to(DEV)
        res = torch.randn((M, N), generator=g).to(DEV)
        out = torch.empty((M, N), device=DEV)
        _lib.check(lib.b200sam_linear_f32(A.data_ptr(), A2.data_ptr(), 64, W.data_ptr(), b.data_ptr(), res.data_ptr(),
                                          out.data_ptr(), M, N, K, act, _lib.current_stream()))
        torch.cuda.synchronize()
        idx = torch.arange(M, device=DEV) % 64
        ref = (A.double() + A2.double()[idx]) @ W.double().T + b.double()
        ref = F.relu(ref) if act == 1 else (F.gelu(ref) if act == 2 else ref)
        ref = (ref + res.double()).float()
        assert torch.allclose(out, ref, atol=2e-5, rtol=1e-5), (M, N, K, (out - ref).abs().max().item())
