"""Mask statistics of the upstream API surface (reference: segment_anything/utils/amg.py): the stability score and
the mask -> box conversion, on the GPU through the C ABI (csrc/amg.cu).  Same signatures, shapes, dtypes and edge-case
behaviour (NaN for an empty low-threshold mask, [0, 0, 0, 0] for an empty mask, float zeros for an empty batch)."""
import torch

from ... import _lib


def calculate_stability_score(masks: torch.Tensor, mask_threshold: float, threshold_offset: float) -> torch.Tensor:
    """IoU between the masks thresholded at mask_threshold +/- threshold_offset (amg.py:154-176).
    masks: [..., H, W] float logits on a CUDA device -> [...] float32."""
    _lib.require_cuda(masks.device if masks.is_cuda else None)
    assert masks.is_cuda and masks.dim() >= 2, "expected [..., H, W] logits on a CUDA device"
    lib = _lib.load()
    lead, (H, W) = masks.shape[:-2], masks.shape[-2:]
    x = masks.float().contiguous().view(-1, H, W)
    n = x.shape[0]
    out = torch.empty((n,), dtype=torch.float32, device=x.device)
    if n:
        scratch = torch.empty((n, 2), dtype=torch.int32, device=x.device)
        _lib.run(x.device, lib.b200sam_stability_score, x.data_ptr(), n, H, W, float(mask_threshold + threshold_offset),
                                               float(mask_threshold - threshold_offset), out.data_ptr(),
                                               scratch.data_ptr(), what="b200sam_stability_score")
    return out.view(lead)


def batched_mask_to_box(masks: torch.Tensor) -> torch.Tensor:
    """XYXY boxes around bool masks, [0, 0, 0, 0] for an empty mask (amg.py:303-346).
    masks: C1 x C2 x ... x H x W bool on a CUDA device -> C1 x C2 x ... x 4 int64."""
    if torch.numel(masks) == 0:
        return torch.zeros(*masks.shape[:-2], 4, device=masks.device)
    _lib.require_cuda(masks.device if masks.is_cuda else None)
    assert masks.is_cuda and masks.dim() >= 2, "expected [..., H, W] masks on a CUDA device"
    lib = _lib.load()
    shape = masks.shape
    H, W = shape[-2:]
    m = (masks if masks.dtype == torch.bool else masks != 0).contiguous().view(-1, H, W)
    n = m.shape[0]
    out = torch.empty((n, 4), dtype=torch.int64, device=m.device)
    scratch = torch.empty((n, 4), dtype=torch.int32, device=m.device)
    _lib.run(m.device, lib.b200sam_mask_to_box, m.data_ptr(), n, H, W, out.data_ptr(), scratch.data_ptr(), what="b200sam_mask_to_box")
    return out.reshape(*shape[:-2], 4) if len(shape) > 2 else out[0]
