"""CPU-side checks of the C-ABI library: it loads without a GPU and exports every symbol the header declares."""
import ctypes as C
import re
from pathlib import Path

import pytest

from samcarriestheburden_b200 import _lib

ROOT = Path(__file__).resolve().parents[1]


def _header_symbols():
    text = (ROOT / "include" / "b200sam.h").read_text()
    return sorted(set(re.findall(r"B200SAM_API[^;(]*?\b(b200sam_\w+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    syms = _header_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/b200sam.h but not exported"
    assert set(syms) == set(_lib.EXPORTED_SYMBOLS), set(syms) ^ set(_lib.EXPORTED_SYMBOLS)
    assert lib.b200sam_abi_version() == _lib.ABI_VERSION == 2


def test_weight_tables_are_self_describing():
    lib = _lib.load()
    gmask = (1 << 7) | (1 << 15) | (1 << 23) | (1 << 31) - (1 << 32)
    cfg = _lib.EncoderConfig(1280, 32, 16, gmask, 256, _lib.OPERAND_FP16, 0)
    n = lib.b200sam_encoder_weight_count(C.byref(cfg))
    assert n == 3 + 32 * 17 + 6
    names = [lib.b200sam_encoder_weight_name(C.byref(cfg), i).decode() for i in range(n)]
    assert names[0] == "image_encoder.patch_embed.proj.weight|op16_flat"
    assert "image_encoder.blocks.31.mlp.lin2.bias|f32" in names
    assert "image_encoder.blocks.3.attn.qkv.weight|op16" in names
    assert names[-3] == "image_encoder.neck.2.weight|op16_tap"
    assert lib.b200sam_encoder_weight_name(C.byref(cfg), n) is None
    # LayerNorm folding: the linears after norm1 / norm2 are described by (linear prefix, fold_*, norm prefix)
    fcfg = _lib.EncoderConfig(1280, 32, 16, gmask, 256, _lib.OPERAND_FP16, _lib.ENC_LN_FUSED)
    assert lib.b200sam_encoder_weight_count(C.byref(fcfg)) == n
    fnames = [lib.b200sam_encoder_weight_name(C.byref(fcfg), i).decode() for i in range(n)]
    for tag in ("fold_w", "fold_s", "fold_c"):
        assert f"image_encoder.blocks.5.attn.qkv|{tag}|image_encoder.blocks.5.norm1" in fnames
        assert f"image_encoder.blocks.5.mlp.lin1|{tag}|image_encoder.blocks.5.norm2" in fnames
    assert "image_encoder.blocks.5.attn.qkv.bias|op16" in fnames  # pad tokens keep the ORIGINAL qkv bias
    nd = lib.b200sam_decoder_weight_count()
    dn = [lib.b200sam_decoder_weight_name(i).decode() for i in range(nd)]
    assert nd == 134 and len(set(dn)) == nd
    assert "mask_decoder.output_upscaling.0.weight|convT" in dn
    # every decoder/encoder key exists in the reference-compatible state_dict
    from samcarriestheburden_b200.segment_anything import sam_model_registry
    sd = sam_model_registry["vit_h"]().state_dict()
    for nm in names + fnames + dn:
        key, packing, *rest = (nm.split("|") + [""])[:3] if "|" in nm else (nm, "", "")
        if packing == "cat4":
            assert all(f"{key}.{i}.weight" in sd for i in range(4))
        elif packing.startswith("fold_"):
            assert all(k in sd for k in (key + ".weight", key + ".bias", rest[0] + ".weight", rest[0] + ".bias")), nm
        else:
            assert key in sd, key


def test_workspace_queries_and_argument_errors():
    lib = _lib.load()
    cfg = _lib.EncoderConfig(768, 12, 12, 0b100100100100, 256, _lib.OPERAND_FP16, _lib.ENC_LN_FUSED)
    b1 = lib.b200sam_encoder_workspace_bytes(C.byref(cfg), 1)
    b4 = lib.b200sam_encoder_workspace_bytes(C.byref(cfg), 4)
    assert b1 > 4096 * 768 * (4 + 2 + 6 + 2 + 8) and 3.9 * b1 < b4 < 4.1 * b1
    assert lib.b200sam_decoder_workspace_bytes(17, 18) > 17 * 4096 * 256 * 4
    assert lib.b200sam_decoder_workspace_bytes(0, 2) == 0
    # bad arguments fail loudly with a message (no GPU work is issued)
    handle = C.c_void_p()
    bad = _lib.EncoderConfig(700, 12, 12, 0, 256, 0, 0)
    arr = (C.c_void_p * 1)(None)
    assert lib.b200sam_encoder_create(C.byref(bad), arr, 1, C.byref(handle)) != 0
    assert b"embed_dim" in lib.b200sam_last_error()
    bad_fmt = _lib.EncoderConfig(768, 12, 12, 0, 256, 5, 0)
    assert lib.b200sam_encoder_create(C.byref(bad_fmt), arr, 1, C.byref(handle)) != 0
    assert b"operand_format" in lib.b200sam_last_error()
    assert lib.b200sam_decode(None, None, 1, 0, None, None, None, 0, None, None, None, 0, None) != 0
    with pytest.raises(_lib.B200SamError):
        _lib.check(2, "demo")


def test_product_path_has_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CPU-only behaviour")
    from samcarriestheburden_b200.segment_anything import SamPredictor, sam_model_registry
    from samcarriestheburden_b200.segment_anything.utils.prompt_utils import PromptExtractor
    import numpy as np
    sam = sam_model_registry["vit_b"]()
    pred = SamPredictor(sam)
    with pytest.raises(_lib.B200SamError):
        pred.set_image(np.zeros((64, 64, 3), np.uint8))
    with pytest.raises(RuntimeError):
        pred.predict(box=np.array([0, 0, 5, 5]))
    with pytest.raises(_lib.B200SamError):
        PromptExtractor(torch.zeros((2, 8, 8), dtype=torch.bool)).extract()
