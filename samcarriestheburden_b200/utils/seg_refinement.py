"""SAM-based segmentation refinement (reference: utils/seg_refinement.py:75-116, `SAMSegRefiner`).

The reference loops over classes with two B=1 decoder calls each; here every pass is ONE batched decode over all
prompts of the image, the pass-1 full-resolution mask (dead work in the reference when self-refinement is on,
SURVEY.md call stack B) is skipped, and upscale + threshold + nearest-exact resample run as one fused kernel."""
from __future__ import annotations

from abc import ABC, abstractmethod
from typing import List, Union

import torch

from ..segment_anything.modeling.sam import upscale_masks
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
        ex = PromptExtractor(seg)
        packed = ex._extracted  # CPU int32 [C, 8]: sx, sy, xmin, ymin, xmax, ymax, has_seed, has_box (one D2H copy)
        idx = torch.nonzero(packed[:, 6]).flatten()
        est_dice = torch.full((seg.shape[0],), float("nan"))
        K = int(idx.numel())
        if K == 0:
            return seg, est_dice
        if K == 1:
            ex.extract()  # raises like the reference: torch.cat of an empty neg-seed list (prompt_utils.py:122-123)
        head = self.sam_predictor
        feats, input_size, original_size = head._entry(file_name)
        # vectorised restatement of scale_coords / scale_box (prompt_utils.py:146-184) for all K prompts at once:
        # fp32 (target / original) flipped to (x, y), one multiply per coordinate - same arithmetic, on the host
        scale = (torch.tensor(input_size, dtype=torch.float) / torch.tensor(tuple(seg.shape[-2:]), dtype=torch.float)).flip(-1)
        pos = packed[idx, 0:2]                                   # [K, 2] (x, y)
        boxes = (packed[idx, 2:6].reshape(K, 2, 2).float() * scale).reshape(K, 4)
        sel = (~torch.eye(K, dtype=torch.bool)).nonzero()[:, 1].reshape(K, K - 1)  # other classes, in class order
        pts = torch.cat([pos[:, None, :], pos[sel]], dim=1).float() * scale       # [K, K, 2]: pos seed, then negatives
        labs = torch.zeros((K, K), dtype=torch.int32)
        labs[:, 0] = 1

        def run(kinds, mask_prev, upscale, small_size):
            use_pts = [k for k in kinds if k in ("pos_points", "neg_points")]
            p = l = None
            if use_pts:
                cols = ([0] if "pos_points" in kinds else []) + (list(range(1, K)) if "neg_points" in kinds else [])
                p, l = pts[:, cols], labs[:, cols]
            bx = boxes if "box" in kinds else None
            low, iou = head.sam.decode_prompts(feats, p, l, bx, mask_prev, multimask_output=False)
            small = None
            if upscale:
                _, small = upscale_masks(low, input_size, original_size, head.img_enc_img_size, head.mask_threshold,
                                         small_size=small_size)
            return low, iou, small

        small_size = tuple(seg.shape[-2:])
        if self.prompts2use2nd is None:
            _, score, small = run(self.prompts2use1st, None, True, small_size)
        else:
            low1, _, _ = run(self.prompts2use1st, None, False, None)  # pass-1 native mask is dead work
            _, score, small = run(self.prompts2use2nd, low1, True, small_size)
        seg[idx.to(seg.device)] = small[:, 0]
        s = score[:, 0].float().cpu()
        est_dice[idx] = 2 * s / (1 + s)
        return seg, est_dice
