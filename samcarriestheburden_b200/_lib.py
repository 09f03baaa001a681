"""ctypes binding of libb200sam.so (the C ABI declared in include/b200sam.h).

There is no CPU fallback: if the library cannot be loaded (or, later, no CUDA device is present) the
product path raises.  Tensors cross the boundary as raw device pointers (`tensor.data_ptr()`) plus the
current torch CUDA stream handle — no torch types in any signature.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

_HERE = Path(__file__).resolve().parent
LIB_PATH = _HERE / "libb200sam.so"

_lib = None


class B200SamError(RuntimeError):
    pass


class EncoderConfig(C.Structure):
    _fields_ = [("embed_dim", C.c_int), ("depth", C.c_int), ("num_heads", C.c_int),
                ("global_attn_mask", C.c_int), ("out_chans", C.c_int), ("operand_format", C.c_int),
                ("flags", C.c_int)]


OPERAND_BF16, OPERAND_FP16 = 0, 1
ENC_LN_FUSED = 1
ABI_VERSION = 2


_vp, _i, _f, _sz = C.c_void_p, C.c_int, C.c_float, C.c_size_t
_PROTOTYPES = {
    "b200sam_last_error": (C.c_char_p, []),
    "b200sam_abi_version": (_i, []),
    "b200sam_encoder_weight_count": (_i, [C.POINTER(EncoderConfig)]),
    "b200sam_encoder_weight_name": (C.c_char_p, [C.POINTER(EncoderConfig), _i]),
    "b200sam_encoder_workspace_bytes": (_sz, [C.POINTER(EncoderConfig), _i]),
    "b200sam_encoder_create": (_i, [C.POINTER(EncoderConfig), C.POINTER(_vp), _i, C.POINTER(_vp)]),
    "b200sam_encoder_destroy": (None, [_vp]),
    "b200sam_encoder_forward": (_i, [_vp, _vp, _i, _i, _i, _i, C.POINTER(_f), C.POINTER(_f), _vp, _vp, _sz, _vp]),
    "b200sam_prompt_extract_scratch_bytes": (_sz, [_i, _i]),
    "b200sam_prompt_extract": (_i, [_vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    "b200sam_decoder_weight_count": (_i, []),
    "b200sam_decoder_weight_name": (C.c_char_p, [_i]),
    "b200sam_decoder_workspace_bytes": (_sz, [_i, _i]),
    "b200sam_decoder_create": (_i, [C.POINTER(_vp), _i, C.POINTER(_vp), _vp]),
    "b200sam_decoder_destroy": (None, [_vp]),
    "b200sam_decoder_copy_dense_pe": (_i, [_vp, _vp, _vp]),
    "b200sam_decode": (_i, [_vp, _vp, _i, _i, _vp, _vp, _vp, _i, _vp, _vp, _vp, _sz, _vp]),
    "b200sam_decoder_workspace_bytes_batch": (_sz, [_i, _i, _i]),
    "b200sam_decode_batch": (_i, [_vp, _vp, _i, _vp, _i, _i, _vp, _vp, _vp, _i, _vp, _vp, _vp, _sz, _vp]),
    "b200sam_prompt_encode": (_i, [_vp, _vp, _vp, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    "b200sam_decode_embedded": (_i, [_vp, _vp, _i, _vp, _i, _i, _vp, _vp, _i, _vp, _vp, _vp, _sz, _vp]),
    "b200sam_upscale_threshold": (_i, [_vp, _i, _i, _i, _i, _i, _i, _i, _f, _vp, _vp, _vp, _i, _i, _vp]),
    "b200sam_unet_weight_count": (_i, []),
    "b200sam_unet_weight_name": (C.c_char_p, [_i]),
    "b200sam_unet_conv_kp": (_i, [_i]),
    "b200sam_unet_create": (_i, [_i, _i, _i, C.POINTER(_vp), _i, C.POINTER(_vp), _vp]),
    "b200sam_unet_destroy": (None, [_vp]),
    "b200sam_unet_workspace_bytes": (_sz, [_vp, _i, _i, _i]),
    "b200sam_unet_forward": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp, _vp, _sz, _vp]),
    "b200sam_resize_ksize": (_i, [_i, _i]),
    "b200sam_resize_coeffs_host": (_i, [_i, _i, _vp, _vp]),
    "b200sam_resize_u8": (_i, [_vp, _i, _i, _i, _vp, _vp, _i, _vp, _vp, _i, _i, _i, _vp, _vp, _i, _vp]),
    "b200sam_cvresize_coeffs_host": (_i, [_i, _i, _i, _vp, _vp]),
    "b200sam_cvresize_linear_u8": (_i, [_vp, _i, _i, _i, _vp, _vp, _vp, _vp, _i, _i, _vp, _vp, _f, _f, _vp]),
    "b200sam_cvresize_cubic_coeffs_host": (_i, [_i, _i, _vp, _vp]),
    "b200sam_medsam_preprocess": (_i, [_vp, _i, _i, _vp, _vp, _vp, _vp, _i, _vp, _vp, _vp, _vp]),
    "b200sam_stability_score": (_i, [_vp, _i, _i, _i, _f, _f, _vp, _vp, _vp]),
    "b200sam_mask_to_box": (_i, [_vp, _i, _i, _i, _vp, _vp, _vp]),
    "b200sam_ccl_scratch_bytes": (_sz, [_i, _i, _i]),
    "b200sam_ccl_select": (_i, [_vp, _i, _i, _i, _i, _f, _i, _vp, _vp, _vp]),
    "b200sam_morph_flat": (_i, [_vp, _i, _i, _i, _vp, _i, _i, _i, _i, _i, _vp, _vp]),
    "b200sam_gemm_bf16": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "b200sam_gemm_f16": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "b200sam_gemm_ln_residual": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "b200sam_gemm_ln_folded": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _f, _vp, _i, _i, _i, _i, _i, _vp]),
    "b200sam_layernorm": (_i, [_vp, _vp, _vp, _f, _i, _i, _vp, _i, _vp]),
    "b200sam_encoder_attention": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "b200sam_preprocess_patchify": (_i, [_vp, _i, _i, _i, _i, C.POINTER(_f), C.POINTER(_f), _vp, _i, _vp]),
    "b200sam_set_gemm_pair": (_i, [_i]),
    "b200sam_gemm_pair_max_clusters": (_i, []),
    "b200sam_timing_start": (_i, [_i]),
    "b200sam_timing_stop": (_i, [_vp, _vp, _vp, _vp, _i, _vp]),
    "b200sam_linear_f32": (_i, [_vp, _vp, _i, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp]),
}
EXPORTED_SYMBOLS = tuple(_PROTOTYPES)


def load(build_if_missing: bool = True):
    """dlopen the in-tree library (building it with nvcc if absent) and attach prototypes."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        if not build_if_missing:
            raise B200SamError(f"{LIB_PATH} is missing; run `python -m samcarriestheburden_b200.build`")
        from . import build as _build
        _build.build()
    lib = C.CDLL(str(LIB_PATH))
    for name, (res, args) in _PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError here == ABI drift between header and library
        fn.restype = res
        fn.argtypes = args
    if lib.b200sam_abi_version() != ABI_VERSION:
        raise B200SamError(f"{LIB_PATH} has ABI version {lib.b200sam_abi_version()}, this package needs {ABI_VERSION}: "
                           "rebuild with `python -m samcarriestheburden_b200.build --force`")
    _lib = lib
    return lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().b200sam_last_error()
        raise B200SamError(f"{what or 'b200sam call'} failed (rc={rc}): {msg.decode() if msg else '?'}")


def run(device, fn, *args, what: str = "") -> None:
    """Call the stream-taking C-ABI entry point `fn(*args, stream)` with `device` current and torch's current stream
    of THAT device as the last argument; raises B200SamError on a non-zero status."""
    import torch
    with torch.cuda.device(device):
        check(fn(*args, torch.cuda.current_stream(device).cuda_stream), what)


class KernelTiming:
    """Context manager around b200sam_timing_start / _stop: CUDA-event durations of every tcgen05 GEMM / attention launch
    issued inside the block, as a list of dicts {kind, work (FLOPs), dims, ms}."""
    KINDS = {0: "gemm", 1: "window_attention", 2: "global_attention"}

    def __init__(self, capacity: int = 1 << 16):
        self.capacity = capacity
        self.records = []

    def __enter__(self):
        check(load().b200sam_timing_start(self.capacity), "b200sam_timing_start")
        return self

    def __exit__(self, *exc):
        n = self.capacity
        kinds, work, dims, ms = (C.c_int * n)(), (C.c_double * n)(), (C.c_int * (3 * n))(), (C.c_float * n)()
        cnt = C.c_int(0)
        check(load().b200sam_timing_stop(kinds, work, dims, ms, n, C.byref(cnt)), "b200sam_timing_stop")
        self.records = [{"kind": self.KINDS.get(kinds[i], str(kinds[i])), "work": work[i],
                         "dims": (dims[3 * i], dims[3 * i + 1], dims[3 * i + 2]), "ms": ms[i]} for i in range(cnt.value)]
        return False


def ptr(t) -> int | None:
    """Device pointer of a torch tensor (None -> NULL)."""
    if t is None:
        return None
    return t.data_ptr()


def current_stream(device=None) -> int:
    """Handle of torch's current stream on `device` (default: the current device)."""
    import torch
    return torch.cuda.current_stream(device).cuda_stream


def on_device(device):
    """Context manager: make `device` the current CUDA device for the C-ABI calls inside it.  The library launches on
    the current device (kernel attributes, SM count and descriptor caches are per device), so every wrapper enters this
    with the device that owns its tensors (a model on cuda:1 works while cuda:0 is current)."""
    import torch
    return torch.cuda.device(device)


def require_cuda(device=None):
    import torch
    if not torch.cuda.is_available():
        raise B200SamError("b200sam has no CPU fallback: a CUDA (sm_100a) device is required")
    return torch.device(device if device is not None else "cuda")
