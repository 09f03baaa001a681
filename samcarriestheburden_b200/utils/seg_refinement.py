"""SAM-based segmentation refinement (reference: utils/seg_refinement.py:75-116, `SAMSegRefiner`).

The reference loops over classes with two B=1 decoder calls each; here every pass is ONE batched decode over all
prompts of the image, the pass-1 full-resolution mask (dead work in the reference when self-refinement is on,
SURVEY.md call stack B) is skipped, and upscale + threshold + nearest-exact resample run as one fused kernel."""
from __future__ import annotations

from abc import ABC, abstractmethod
from typing import List, Union

import torch

from ..segment_anything.sam_mask_decoder_head import SAMMaskDecoderHead
from ..segment_anything.utils.prompt_utils import PromptExtractor


class SegRefiner(ABC):
    @abstractmethod
    def refine(self, seg: torch.Tensor, file_name: str = None) -> torch.Tensor:
        pass


class SAMSegRefiner(SegRefiner):
    def __init__(self, sam_type: str, device: str, prompts2use: Union[List[List[str]], List[str]],
                 sam_predictor: SAMMaskDecoderHead = None):
        """Same arguments as the reference (:76-97); `sam_predictor` optionally injects a ready decoder head
        (e.g. over an in-HBM EmbeddingStore) instead of the hard-coded checkpoint / h5 paths."""
        if sam_predictor is None:
            if sam_type == "SAM":
                cfg = ("data/sam_vit_h_4b8939.pth", "vit_h", "data/graz_sam_img_embedding.h5")
            elif sam_type == "MedSAM":
                cfg = ("data/medsam_vit_b.pth", "vit_b", "data/graz_medsam_img_embedding.h5")
            else:
                raise NotImplementedError(f"Unknown SAM type: {sam_type}")
            sam_predictor = SAMMaskDecoderHead(cfg[0], cfg[1], device, cfg[2])
        self.sam_predictor = sam_predictor
        if isinstance(prompts2use[0], list):
            self.prompts2use1st = prompts2use[0]
            assert len(prompts2use[1]) > 0, "2nd prompt list should not be empty"
            self.prompts2use2nd = prompts2use[1]
            self.self_refine = True
        else:
            self.prompts2use1st = prompts2use
            self.prompts2use2nd = None
            self.self_refine = False

    @torch.inference_mode()
    def refine(self, seg: torch.Tensor, file_name: str):
        """seg: [C,H,W] (bool or probabilities>0 -> bool like the reference's `.bool()`), returns
        (seg bool [C,H,W], est_dice float [C] with NaN for classes without prompts)."""
        seg = seg.bool().to(self.sam_predictor.device)
        prompts = PromptExtractor(seg).extract()
        est_dice = torch.full((seg.shape[0],), float("nan"))
        if not prompts:
            return seg, est_dice
        small_size = tuple(seg.shape[-2:])
        if self.prompts2use2nd is None:
            _, score, _, small = self.sam_predictor.predict_masks_batched(file_name, prompts, self.prompts2use1st,
                                                                        small_size=small_size)
        else:
            _, _, low1, _ = self.sam_predictor.predict_masks_batched(file_name, prompts, self.prompts2use1st,
                                                                     upscale=False)
            _, score, _, small = self.sam_predictor.predict_masks_batched(file_name, prompts, self.prompts2use2nd,
                                                                        mask_prev_iter=low1, small_size=small_size)
        idx = torch.tensor([p.class_idx for p in prompts], device=seg.device)
        seg[idx] = small[:, 0]
        s = score[:, 0].float().cpu()
        est_dice[idx.cpu()] = 2 * s / (1 + s)
        return seg, est_dice
