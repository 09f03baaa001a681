"""ViT image encoder (reference: segment_anything/modeling/image_encoder.py).

Same constructor, attributes and state_dict as the reference `ImageEncoderViT`; `forward` runs the whole
encoder through `b200sam_encoder_forward` (tcgen05 GEMMs + fused attention kernels, see csrc/encoder.cu).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Tuple, Type

import torch
import torch.nn as nn

from ... import _lib
from .common import FusedAway, LayerNorm2d, MLPBlock


class PatchEmbed(FusedAway):
    def __init__(self, kernel_size=(16, 16), stride=(16, 16), padding=(0, 0), in_chans=3, embed_dim=768) -> None:
        super().__init__()
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=kernel_size, stride=stride, padding=padding)


class Attention(FusedAway):
    def __init__(self, dim, num_heads=8, qkv_bias=True, use_rel_pos=False, rel_pos_zero_init=True,
                 input_size: Optional[Tuple[int, int]] = None) -> None:
        super().__init__()
        self.num_heads = num_heads
        head_dim = dim // num_heads
        self.scale = head_dim ** -0.5
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.proj = nn.Linear(dim, dim)
        self.use_rel_pos = use_rel_pos
        if use_rel_pos:
            assert input_size is not None, "Input size must be provided if using relative positional encoding."
            self.rel_pos_h = nn.Parameter(torch.zeros(2 * input_size[0] - 1, head_dim))
            self.rel_pos_w = nn.Parameter(torch.zeros(2 * input_size[1] - 1, head_dim))


class Block(FusedAway):
    def __init__(self, dim, num_heads, mlp_ratio=4.0, qkv_bias=True, norm_layer=nn.LayerNorm, act_layer=nn.GELU,
                 use_rel_pos=False, rel_pos_zero_init=True, window_size=0, input_size=None) -> None:
        super().__init__()
        self.norm1 = norm_layer(dim)
        self.attn = Attention(dim, num_heads=num_heads, qkv_bias=qkv_bias, use_rel_pos=use_rel_pos,
                              rel_pos_zero_init=rel_pos_zero_init,
                              input_size=input_size if window_size == 0 else (window_size, window_size))
        self.norm2 = norm_layer(dim)
        self.mlp = MLPBlock(embedding_dim=dim, mlp_dim=int(dim * mlp_ratio), act=act_layer)
        self.window_size = window_size


_OP_DTYPE = {"bf16": torch.bfloat16, "fp16": torch.float16}


def _own(t: torch.Tensor, src: torch.Tensor) -> torch.Tensor:
    """A packed tensor never aliases the live parameter it was made from (an in-place weight edit must not reach a
    half-updated engine)."""
    return t.clone() if t.data_ptr() == src.data_ptr() else t


def _pack(sd, spec: str, op_dtype: torch.dtype, device, cache: dict) -> torch.Tensor:
    """Pack one weight slot named by the library as "key|packing[|norm prefix]" (csrc/encoder.cu)."""
    key, packing, *rest = spec.split("|")
    if packing.startswith("fold_"):
        # LayerNorm `rest[0]` folded into the linear `key` (csrc/gemm_epilogue.cuh):
        #   LN(x) W^T + b = rstd * (x (gamma*W)^T - mean * rowsum(gamma*W)) + (beta W^T + b)
        # One-time weight preparation in float64 with elementwise / reduction ops only (no library GEMV / GEMM anywhere on
        # the product path); the three slots of a linear share one computation.
        hit = cache.get((key, rest[0]))
        if hit is None:
            W, b = sd[key + ".weight"].detach().to(device).double(), sd[key + ".bias"].detach().to(device).double()
            gamma = sd[rest[0] + ".weight"].detach().to(device).double()
            beta = sd[rest[0] + ".bias"].detach().to(device).double()
            Wf = (W * gamma[None, :]).float().to(op_dtype).contiguous()
            hit = {"fold_w": Wf,
                   # sums of the ROUNDED weights: the mean term then cancels exactly against the MMA
                   "fold_s": Wf.double().sum(1).float().contiguous(),
                   "fold_c": ((W * beta[None, :]).sum(1) + b).float().contiguous()}
            cache[(key, rest[0])] = hit
        if packing not in hit:
            raise ValueError(f"unknown packing {packing}")
        return hit[packing]
    if key not in sd:
        raise _lib.B200SamError(f"encoder weight {key} missing from the module state_dict")
    src = sd[key].detach()
    t = src.to(device)
    if packing == "none":
        return torch.zeros(4, dtype=torch.float32, device=device)
    if packing == "f32":
        return _own(t.float().contiguous(), src)
    if packing == "op16":
        return _own(t.to(op_dtype).contiguous(), src)
    if packing == "op16_flat":  # conv weight [out, in, kh, kw] -> [out, in*kh*kw]
        return _own(t.reshape(t.shape[0], -1).to(op_dtype).contiguous(), src)
    if packing == "op16_tap":  # 3x3 conv weight -> [out, (ky*3+kx)*Cin + c]
        return t.permute(0, 2, 3, 1).reshape(t.shape[0], -1).to(op_dtype).contiguous()
    if packing == "f32_tokens":  # pos_embed [1, 64, 64, D] -> [4096, D]
        return _own(t.reshape(-1, t.shape[-1]).float().contiguous(), src)
    raise ValueError(f"unknown packing {packing}")


class _EncoderEngine:
    """Owns the packed weights + the C-side encoder handle for one device."""

    def __init__(self, module: "ImageEncoderViT", device: torch.device) -> None:
        lib = _lib.load()
        self.lib = lib
        self.device = device
        gmask = 0
        for i, blk in enumerate(module.blocks):
            if blk.window_size == 0:
                gmask |= 1 << i
        self.operand_format, self.ln_fused = module.operand_format, module.ln_fused
        op_dtype = _OP_DTYPE[self.operand_format]
        self.cfg = _lib.EncoderConfig(module.embed_dim, len(module.blocks), module.num_heads, gmask, module.out_chans,
                                      _lib.OPERAND_FP16 if self.operand_format == "fp16" else _lib.OPERAND_BF16,
                                      _lib.ENC_LN_FUSED if self.ln_fused else 0)
        sd = {"image_encoder." + k: v for k, v in module.state_dict().items()}
        n = lib.b200sam_encoder_weight_count(C.byref(self.cfg))
        cache: dict = {}
        self.packed = [_pack(sd, lib.b200sam_encoder_weight_name(C.byref(self.cfg), i).decode(), op_dtype, device, cache)
                       for i in range(n)]
        arr = (C.c_void_p * n)(*[t.data_ptr() for t in self.packed])
        handle = C.c_void_p()
        with _lib.on_device(device):
            _lib.check(lib.b200sam_encoder_create(C.byref(self.cfg), arr, n, C.byref(handle)), "b200sam_encoder_create")
        self.handle = handle
        self._ws: Optional[torch.Tensor] = None
        self._ws_batch = 0

    def workspace(self, batch: int) -> torch.Tensor:
        """One workspace sized for the largest batch seen so far (a ragged last batch reuses it)."""
        if self._ws is None or batch > self._ws_batch:
            nbytes = self.lib.b200sam_encoder_workspace_bytes(C.byref(self.cfg), batch)
            self._ws = None  # release before growing
            self._ws = torch.empty(nbytes + 1024, dtype=torch.uint8, device=self.device)
            self._ws_batch = batch
        return self._ws

    def forward(self, image: torch.Tensor, mean, std, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        assert image.dim() == 4 and image.shape[1] == 3, "expected [B,3,h,w]"
        B, _, h, w = image.shape
        if image.dtype != torch.uint8:
            image = image.float()
        image = image.contiguous()
        if out is None:
            out = torch.empty((B, self.cfg.out_chans, 64, 64), dtype=torch.float32, device=self.device)
        ws = self.workspace(B)
        base = (ws.data_ptr() + 1023) & ~1023
        mean3 = (C.c_float * 3)(*mean)
        std3 = (C.c_float * 3)(*std)
        _lib.run(self.device, self.lib.b200sam_encoder_forward, self.handle, image.data_ptr(),
                 int(image.dtype == torch.uint8), B, h, w, mean3, std3, out.data_ptr(), base,
                 ws.numel() - (base - ws.data_ptr()), what="b200sam_encoder_forward")
        return out

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                self.lib.b200sam_encoder_destroy(self.handle)
        except Exception:
            pass


def _default_operands() -> str:
    v = os.environ.get("B200SAM_ENCODER_OPERANDS", "fp16").lower()
    if v not in _OP_DTYPE:
        raise _lib.B200SamError(f"B200SAM_ENCODER_OPERANDS={v!r}: expected 'fp16' or 'bf16'")
    return v


class ImageEncoderViT(nn.Module):
    def __init__(self, img_size: int = 1024, patch_size: int = 16, in_chans: int = 3, embed_dim: int = 768,
                 depth: int = 12, num_heads: int = 12, mlp_ratio: float = 4.0, out_chans: int = 256,
                 qkv_bias: bool = True, norm_layer: Type[nn.Module] = nn.LayerNorm,
                 act_layer: Type[nn.Module] = nn.GELU, use_abs_pos: bool = True, use_rel_pos: bool = False,
                 rel_pos_zero_init: bool = True, window_size: int = 0,
                 global_attn_indexes: Tuple[int, ...] = ()) -> None:
        super().__init__()
        if img_size != 1024 or patch_size != 16 or in_chans != 3 or not use_abs_pos or not use_rel_pos \
                or window_size != 14 or mlp_ratio != 4:
            raise NotImplementedError("b200sam implements the SAM configuration only: 1024 px, 16 px patches, "
                                      "abs+rel pos, 14x14 windows, mlp_ratio 4 (build_sam.py:62-80)")
        self.img_size = img_size
        self.embed_dim, self.num_heads, self.out_chans = embed_dim, num_heads, out_chans
        self.patch_embed = PatchEmbed((patch_size, patch_size), (patch_size, patch_size), in_chans=in_chans,
                                      embed_dim=embed_dim)
        self.pos_embed = nn.Parameter(torch.zeros(1, img_size // patch_size, img_size // patch_size, embed_dim))
        self.blocks = nn.ModuleList()
        for i in range(depth):
            self.blocks.append(Block(dim=embed_dim, num_heads=num_heads, mlp_ratio=mlp_ratio, qkv_bias=qkv_bias,
                                     norm_layer=norm_layer, act_layer=act_layer, use_rel_pos=use_rel_pos,
                                     rel_pos_zero_init=rel_pos_zero_init,
                                     window_size=window_size if i not in global_attn_indexes else 0,
                                     input_size=(img_size // patch_size, img_size // patch_size)))
        self.neck = nn.Sequential(nn.Conv2d(embed_dim, out_chans, kernel_size=1, bias=False), LayerNorm2d(out_chans),
                                  nn.Conv2d(out_chans, out_chans, kernel_size=3, padding=1, bias=False),
                                  LayerNorm2d(out_chans))
        self._engine: Optional[_EncoderEngine] = None
        self._engine_versions: Optional[tuple] = None
        # 16-bit tensor-core operand format: fp16 (11 significand bits) by default, because refined masks at Dice >= 0.999
        # against the fp32 reference need its precision (bf16 operands: 0.997 at random init); "bf16"
        # trades that for the fp32 exponent range.  LayerNorm folding: norm1 / norm2 run inside the GEMM epilogues.
        self.operand_format = _default_operands()
        self.ln_fused = os.environ.get("B200SAM_LN_FUSED", "1") != "0"
        # a parent's load_state_dict never calls the child's override, but it does run the child's post hooks
        self.register_load_state_dict_post_hook(lambda module, incompatible: module._invalidate())

    def _invalidate(self) -> None:
        self._engine = None

    def set_precision(self, operand_format: Optional[str] = None, ln_fused: Optional[bool] = None) -> "ImageEncoderViT":
        """Select the MMA operand format ("fp16" | "bf16") and whether LayerNorm is folded into the GEMMs."""
        if operand_format is not None:
            if operand_format not in _OP_DTYPE:
                raise ValueError(f"operand_format must be 'fp16' or 'bf16', got {operand_format!r}")
            self.operand_format = operand_format
        if ln_fused is not None:
            self.ln_fused = bool(ln_fused)
        self._invalidate()
        return self

    # weights changed / moved -> re-pack lazily
    def _apply(self, fn, *a, **k):
        self._invalidate()
        return super()._apply(fn, *a, **k)

    def _versions(self) -> tuple:
        return tuple(p._version for p in self.parameters())

    def engine(self) -> _EncoderEngine:
        dev = self.pos_embed.device
        if dev.type != "cuda":
            raise _lib.B200SamError("b200sam ImageEncoderViT has no CPU path: move the model to a CUDA device")
        versions = self._versions()  # in-place edits of a parameter bump its version counter: re-pack
        if self._engine is None or self._engine.device != dev or versions != self._engine_versions:
            self._engine = None
            self._engine = _EncoderEngine(self, dev)
            self._engine_versions = versions
        return self._engine

    @torch.no_grad()
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """x: already normalised + padded [B,3,1024,1024] (reference image_encoder.py:106-116)."""
        return self.engine().forward(x, (0.0, 0.0, 0.0), (1.0, 1.0, 1.0))

    @torch.no_grad()
    def forward_raw(self, image: torch.Tensor, pixel_mean, pixel_std) -> torch.Tensor:
        """Fused Sam.preprocess + forward: un-normalised uint8/float [B,3,h,w] with long side 1024."""
        return self.engine().forward(image, pixel_mean, pixel_std)
