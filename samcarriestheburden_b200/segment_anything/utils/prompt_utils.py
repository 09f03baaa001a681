"""Prompt extraction from predicted masks (reference: segment_anything/utils/prompt_utils.py).

`PromptExtractor.extract()` returns the same `Prompt` objects; the per-class nonzero/mean/min/max chain of the
reference (~10 ATen calls + host syncs per class) is one pass of the CUDA kernel b200sam_prompt_extract
(csrc/prompt_extract.cu) followed by ONE device->host copy of C x (2 + 4 + 2) integers."""
from dataclasses import dataclass
from functools import cached_property
from typing import List, Optional, Tuple

import torch

from ... import _lib


@dataclass
class Prompt:
    class_idx: int
    img_size: Tuple[int, int]  # (H, W)
    pos_seeds: torch.Tensor = None
    neg_seeds: torch.Tensor = None
    box: torch.Tensor = None
    mask_logits: torch.Tensor = None


def extract_seeds_boxes(pred_masks: torch.Tensor):
    """pred_masks: bool [N,C,H,W] on a CUDA device -> (seeds int32 [N,C,2] (x,y), boxes int32 [N,C,4],
    has_seed uint8 [N,C], has_box uint8 [N,C]), all on the device (no sync)."""
    assert pred_masks.dim() == 4 and pred_masks.dtype == torch.bool
    if pred_masks.device.type != "cuda":
        raise _lib.B200SamError("b200sam prompt extraction has no CPU path: pass a CUDA tensor")
    lib = _lib.load()
    m = pred_masks.contiguous()
    N, Cn, H, W = m.shape
    dev = m.device
    seeds = torch.empty((N, Cn, 2), dtype=torch.int32, device=dev)
    boxes = torch.empty((N, Cn, 4), dtype=torch.int32, device=dev)
    has_seed = torch.empty((N, Cn), dtype=torch.uint8, device=dev)
    has_box = torch.empty((N, Cn), dtype=torch.uint8, device=dev)
    scratch = torch.empty(max(lib.b200sam_prompt_extract_scratch_bytes(N, Cn) // 8 + 1, 1), dtype=torch.int64, device=dev)
    _lib.run(dev, lib.b200sam_prompt_extract, m.data_ptr(), N, Cn, H, W, seeds.data_ptr(), boxes.data_ptr(),
                                          has_seed.data_ptr(), has_box.data_ptr(), scratch.data_ptr(), what="b200sam_prompt_extract")
    return seeds, boxes, has_seed, has_box


class PromptExtractor:
    def __init__(self, pred_mask: torch.Tensor):
        assert pred_mask.ndim == 3, "pred_mask should be 3D tensor of shape (C, H, W)"
        assert pred_mask.dtype == torch.bool, "pred_mask should be boolean tensor"
        self.pred_mask = pred_mask
        self.num_classes = pred_mask.shape[0]

    @cached_property
    def _extracted(self):
        seeds, boxes, has_seed, has_box = extract_seeds_boxes(self.pred_mask[None])
        packed = torch.cat([seeds[0], boxes[0], has_seed[0][:, None].int(), has_box[0][:, None].int()], dim=1).cpu()
        return packed  # [C, 8] int32: sx, sy, xmin, ymin, xmax, ymax, has_seed, has_box

    @cached_property
    def seeds(self) -> List[Optional[torch.Tensor]]:
        e = self._extracted
        dev = self.pred_mask.device
        return [e[c, 0:2].reshape(1, 2).to(dev) if int(e[c, 6]) else None for c in range(self.num_classes)]

    def _extract_seeds(self, class_idx: int):
        assert class_idx < self.num_classes, "class_idx exceeds number of classes"
        return self.seeds[class_idx]

    def _extract_box(self, class_idx: int):
        assert class_idx < self.num_classes, "class_idx exceeds number of classes"
        e = self._extracted
        return e[class_idx, 2:6].clone().to(self.pred_mask.device) if int(e[class_idx, 7]) else None

    def extract(self, seeds: bool = True, boxes: bool = True, mask: bool = False) -> List[Prompt]:
        """Same contract as the reference (prompt_utils.py:112-143); `mask=True` is unsupported there too
        ("not working yet") and is rejected here."""
        if mask:
            raise NotImplementedError("mask-logit prompts are marked 'not working yet' in the reference and are "
                                      "not part of the refinement hot path")
        prompts = []
        for class_idx in range(self.num_classes):
            if self.seeds[class_idx] is None:
                continue
            p = Prompt(class_idx, tuple(self.pred_mask.shape[-2:]))
            if seeds:
                p.pos_seeds = self.seeds[class_idx]
                others = [self.seeds[i] for i in range(self.num_classes) if i != class_idx and self.seeds[i] is not None]
                p.neg_seeds = torch.cat(others)  # raises on an empty list exactly like the reference (:122-123)
            if boxes:
                p.box = self._extract_box(class_idx)
            prompts.append(p)
        return prompts


def scale_coords(coords: torch.Tensor, original_size: Tuple[int, ...], target_size: Tuple[int, ...]) -> torch.Tensor:
    """(N,2) (x,y) coords from `original_size` (H,W) to `target_size` (H,W) (reference prompt_utils.py:146-166)."""
    assert coords.ndim == 2, "coords should be 2D tensor of shape (N, 2)"
    assert coords.shape[1] == len(original_size) == len(target_size), \
        "coords should have same number of dimensions as original_size and target_size"
    original_size = torch.tensor(original_size, dtype=torch.float, device=coords.device)
    target_size = torch.tensor(target_size, dtype=torch.float, device=coords.device)
    return coords.float() * (target_size / original_size).flip(-1)


def scale_box(box: torch.Tensor, original_size: Tuple[int, ...], target_size: Tuple[int, ...]) -> torch.Tensor:
    assert box.ndim == 2, "box should be 2D tensor of shape (N, 4)"
    assert box.shape[1] == 4, "box should have length 4"
    return scale_coords(box.reshape(-1, 2), original_size, target_size).reshape(-1, 4)
