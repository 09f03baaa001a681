"""Connected-component / morphology pre-processing of U-Net probability maps
(reference: utils/segmentation_preprocessing.py).

`remove_all_but_one_connected_component` keeps the reference signature; the 384 max-pool passes of
kornia.contrib.connected_components plus the per-class unique / argmax host round trips are ONE union-find launch
sequence over all planes (csrc/ccl.cu).  `num_iter` is accepted for signature parity: the result equals the
reference's whenever its label propagation has converged within `num_iter` steps (every component's geodesic radius
around its largest-index pixel <= num_iter, which max(H, W) guarantees for anything but spiral-shaped components)."""
from __future__ import annotations

import numpy as np
import torch

from .. import _lib


def _ccl(prob: torch.Tensor, selection: str) -> torch.Tensor:
    if selection not in ("largest", "highest_probability"):
        raise NotImplementedError(f"Invalid selection: {selection}")
    if prob.device.type != "cuda":
        raise _lib.B200SamError("b200sam connected-component selection has no CPU path: pass a CUDA tensor")
    if selection == "highest_probability":
        assert prob.dtype == torch.float, "prob_mask should be probabilities"
    lib = _lib.load()
    p = prob.float().contiguous()
    H, W = p.shape[-2:]
    n = p.numel() // (H * W)
    out = torch.empty_like(p)
    scratch = torch.empty(lib.b200sam_ccl_scratch_bytes(n, H, W) + 256, dtype=torch.uint8, device=p.device)
    base = (scratch.data_ptr() + 255) & ~255
    # a (N, C, H, W) batch = N reference calls of C planes each (the reference's label 0 is per call)
    per_call = p.shape[-3] if p.dim() >= 3 else n
    _lib.run(p.device, lib.b200sam_ccl_select, p.data_ptr(), n, per_call, H, W, 0.5, int(selection == "largest"),
             out.data_ptr(), base, what="b200sam_ccl_select")
    return out.to(prob.dtype) if prob.dtype != torch.float else out


def remove_all_but_one_connected_component(prob_mask: torch.Tensor, selection: str, num_iter: int = 0) -> torch.Tensor:
    """prob_mask: (C, H, W) -> prob_mask * (selected component of prob_mask > 0.5), per class (reference :7-52)."""
    assert prob_mask.ndim == 3, "segmentation_mask should be 3D tensor of shape (C, H, W)"
    return _ccl(prob_mask, selection)


def remove_all_but_one_connected_component_batch(prob_masks: torch.Tensor, selection: str) -> torch.Tensor:
    """The same for a batch (N, C, H, W) in one launch sequence."""
    assert prob_masks.ndim == 4, "prob_masks should be 4D tensor of shape (N, C, H, W)"
    return _ccl(prob_masks, selection)


def structuring_element(name: str, radius: int) -> np.ndarray:
    """skimage.morphology square / disk / diamond footprints (0/1 uint8), restated (skimage is a reference-side
    dependency only)."""
    if name == "square":
        return np.ones((radius, radius), np.uint8)
    r = np.arange(-radius, radius + 1)
    yy, xx = np.meshgrid(r, r, indexing="ij")
    if name == "disk":
        return (yy ** 2 + xx ** 2 <= radius ** 2).astype(np.uint8)
    if name == "diamond":
        return (np.abs(yy) + np.abs(xx) <= radius).astype(np.uint8)
    if name == "star":
        # skimage.morphology.star(a): the (2a+1)-square overlaid with its 45-degree rotation, on a (2a+1+2*(a//2))^2 grid.
        # skimage builds the rotated square as the convex hull of its four vertex pixels; that hull is restated as the
        # diamond |dy| + |dx| <= c through the vertex centres (skimage is absent here, so this footprint is unpinned; it only
        # feeds SegEnhance.last_preprocessed_seg, which the pipeline never consumes).
        if radius == 1:
            return np.ones((3, 3), np.uint8)
        m, n = 2 * radius + 1, radius // 2
        size = m + 2 * n
        c = (size - 1) // 2
        r2 = np.arange(size) - c
        y2, x2 = np.meshgrid(r2, r2, indexing="ij")
        fp = (np.abs(y2) + np.abs(x2) <= c)
        fp[n:m + n, n:m + n] = True
        return fp.astype(np.uint8)
    raise NotImplementedError(f"structuring element {name!r} is not supported")


def morph_flat(x: torch.Tensor, kernel: np.ndarray, dilate: bool) -> torch.Tensor:
    """Flat grey-scale dilation / erosion of (..., H, W) with a 0/1 footprint anchored at its centre (kh // 2, kw // 2),
    out-of-image taps ignored (kornia.morphology.dilation / erosion, geodesic border)."""
    if x.device.type != "cuda":
        raise _lib.B200SamError("b200sam morphology has no CPU path: pass a CUDA tensor")
    lib = _lib.load()
    xin = x.float().contiguous()
    H, W = xin.shape[-2:]
    se = torch.from_numpy(np.ascontiguousarray(kernel.astype(np.uint8))).to(xin.device)
    out = torch.empty_like(xin)
    _lib.run(xin.device, lib.b200sam_morph_flat, xin.data_ptr(), xin.numel() // (H * W), H, W, se.data_ptr(), se.shape[0], se.shape[1],
                                      se.shape[0] // 2, se.shape[1] // 2, int(dilate), out.data_ptr(), what="b200sam_morph_flat")
    return out
