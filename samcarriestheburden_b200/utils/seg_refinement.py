"""SAM-based segmentation refinement (reference: utils/seg_refinement.py:75-116, `SAMSegRefiner`).

The reference loops over images and classes with two B=1 decoder calls each; here every pass is ONE batched decode
over all prompts of one image (`refine`) or of a whole batch of images (`refine_batch`), the pass-1 full-resolution
mask (dead work in the reference when self-refinement is on, SURVEY.md call stack B) is skipped, and upscale +
threshold + nearest-exact resample run as one fused kernel."""
from __future__ import annotations

from abc import ABC, abstractmethod
from typing import List, Sequence, Union

import torch

from ..segment_anything.modeling.sam import upscale_masks
from ..segment_anything.sam_mask_decoder_head import SAMMaskDecoderHead
from ..segment_anything.utils.prompt_utils import PromptExtractor, extract_seeds_boxes
from . import segmentation_preprocessing


class SegRefiner(ABC):
    @abstractmethod
    def refine(self, seg: torch.Tensor, file_name: str = None) -> torch.Tensor:
        pass


class SegEnhance:
    """Connected-component selection (+ optional flat morphology) in front of a refiner
    (reference: utils/seg_refinement.py:20-72, same constructor arguments)."""

    def __init__(self, refiner: SegRefiner, ccl_selection, morph_op: str, struct_element: str, radius: int, device: str):
        self.last_preprocessed_seg = None
        self.refiner = refiner
        self.ccl_selection = ccl_selection
        if morph_op not in ("erosion", "dilation"):
            raise KeyError(morph_op)
        if struct_element not in ("square", "disk", "diamond", "star"):
            raise KeyError(struct_element)
        self._dilate = morph_op == "dilation"
        if struct_element == "square" and radius == 0:
            radius = 1  # identity for the square element, like the reference
        self._identity = radius == 0 or (struct_element == "square" and radius == 1)
        self._kernel = None if self._identity else segmentation_preprocessing.structuring_element(struct_element, radius)

    def _pre(self, seg: torch.Tensor) -> torch.Tensor:
        if self.ccl_selection is None:
            return seg
        return segmentation_preprocessing._ccl(seg, self.ccl_selection)

    @torch.inference_mode()
    def enhance(self, seg: torch.Tensor, file_name: str = None):
        assert seg.ndim == 3, "seg should be 3D tensor of shape (C, H, W)"
        seg = self._pre(seg.to(self.refiner.sam_predictor.device) if hasattr(self.refiner, "sam_predictor") else seg)
        self.last_preprocessed_seg = seg.float() if self._identity else \
            segmentation_preprocessing.morph_flat(seg, self._kernel, self._dilate)
        return self.refiner.refine(seg, file_name)

    @torch.inference_mode()
    def enhance_batch(self, segs: torch.Tensor, file_names: Sequence[str]):
        """(N, C, H, W) probability maps of N images: one CCL launch sequence + one batched refinement."""
        assert segs.ndim == 4, "segs should be 4D tensor of shape (N, C, H, W)"
        segs = self._pre(segs.to(self.refiner.sam_predictor.device))
        self.last_preprocessed_seg = segs.float() if self._identity else \
            segmentation_preprocessing.morph_flat(segs, self._kernel, self._dilate)
        return self.refiner.refine_batch(segs, file_names)


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
        segs, est = self.refine_batch(seg[None], [file_name])
        return segs[0], est[0]

    @torch.inference_mode()
    def refine_batch(self, segs: torch.Tensor, file_names: Sequence[str]):
        """`refine` for N images at once: segs [N,C,H,W], file_names N embedding keys -> (segs bool [N,C,H,W],
        est_dice [N,C]).  One prompt-extraction launch and ONE device->host copy for the whole batch, each decoder
        pass one launch sequence over the prompts of all images (ragged: images with fewer classes carry fewer
        negative points; absent slots are masked inside the kernels), one fused upscale launch per distinct
        (input_size, original_size).  Every prompt computes exactly what the per-image call computes."""
        head = self.sam_predictor
        dev = head.device
        seg = segs.bool().to(dev)
        N, Cn, H, W = seg.shape
        assert len(file_names) == N, "one file name per image"
        seeds, boxes, has_seed, has_box = extract_seeds_boxes(seg)
        packed = torch.cat([seeds, boxes, has_seed[..., None].int()], dim=2).cpu()  # [N, C, 7]: sx sy x0 y0 x1 y1 has_seed
        est_dice = torch.full((N, Cn), float("nan"))
        present = packed[:, :, 6] != 0
        Ks = present.sum(1)                                       # prompts per image
        NB = int(Ks.sum())
        if NB == 0:
            return seg, est_dice
        if bool((Ks == 1).any()):
            # the reference raises here: torch.cat of an empty neg-seed list (prompt_utils.py:122-123)
            PromptExtractor(seg[int((Ks == 1).nonzero()[0])]).extract()
        img_of, cls_of = present.nonzero(as_tuple=True)           # [NB] each, image-major / class order
        entries = [head._entry(fn) for fn in file_names]
        feats = torch.cat([e[0].reshape(1, 256, 64, 64) for e in entries])
        in_sizes = torch.tensor([e[1] for e in entries], dtype=torch.float)      # [N, 2] (h, w)
        # vectorised restatement of scale_coords / scale_box (prompt_utils.py:146-184): fp32 (target / original)
        # flipped to (x, y), one multiply per coordinate - same arithmetic, on the host
        scale = (in_sizes / torch.tensor((H, W), dtype=torch.float)).flip(-1)[img_of]   # [NB, 2]
        pos = packed[img_of, cls_of, 0:2].float()                 # [NB, 2] (x, y), unscaled
        box = (packed[img_of, cls_of, 2:6].reshape(NB, 2, 2).float() * scale[:, None, :])  # [NB, 2, 2]
        Kmax = int(Ks.max())
        # negatives of prompt p = seeds of the other present classes of its image, in class order
        first = torch.cumsum(Ks, 0) - Ks                          # first prompt of each image
        rank = torch.arange(NB) - first[img_of]                   # prompt's position among its image's prompts
        Kp = Ks[img_of]                                           # [NB] prompts of the prompt's image
        j = torch.arange(Kmax - 1)[None, :]                       # candidate negative slot
        src = j + (j >= rank[:, None]).long()                     # skip the prompt itself
        neg_valid = src < Kp[:, None]
        neg = pos[(first[img_of][:, None] + src.clamp(max=Kmax - 1)).clamp(max=NB - 1)]   # [NB, Kmax-1, 2]

        def assemble(kinds):
            """[points | pad point if no box | box corners | absent slots] per prompt (prompt_encoder.py:73-100)"""
            cs, ls = [], []
            if "pos_points" in kinds:
                cs.append((pos * scale)[:, None, :])
                ls.append(torch.ones((NB, 1), dtype=torch.int32))
            if "neg_points" in kinds:
                cs.append(neg * scale[:, None, :])
                ls.append(torch.where(neg_valid, 0, -2).to(torch.int32))
            if cs and "box" not in kinds:
                cs.append(torch.zeros((NB, 1, 2)))
                ls.append(torch.full((NB, 1), -1, dtype=torch.int32))
            if "box" in kinds:
                cs.append(box)
                ls.append(torch.tensor([[2, 3]], dtype=torch.int32).expand(NB, 2))
            c, l = torch.cat(cs, dim=1), torch.cat(ls, dim=1)
            if bool((l == -2).any()):  # ragged: present slots first (stable), absent slots trailing
                order = torch.argsort((l == -2).to(torch.int8), dim=1, stable=True)
                c = torch.gather(c, 1, order[..., None].expand(-1, -1, 2))
                l = torch.gather(l, 1, order)
            return c.to(dev, non_blocking=True), l.to(dev, non_blocking=True)

        image_of_dev = img_of.to(torch.int32).to(dev, non_blocking=True)
        engine = head.sam.decoder_engine()

        def run(kinds, mask_prev):
            c, l = assemble(kinds)
            return engine.decode(feats, c, l, mask_prev, False, image_of=image_of_dev)

        if self.prompts2use2nd is None:
            low, score = run(self.prompts2use1st, None)
        else:
            low1, _ = run(self.prompts2use1st, None)  # the pass-1 native-resolution mask is dead work
            low, score = run(self.prompts2use2nd, low1)
        # fused upscale + threshold + nearest-exact resample, one launch per distinct (input, original) size
        # (the native-resolution masks are materialised like in the reference and kept in `last_native_masks`)
        small = torch.empty((NB, 1, H, W), dtype=torch.bool, device=dev)
        self.last_native_masks = []  # [(prompt indices | None = all, bool [k,1,H0,W0])]
        groups = {}
        for n, e in enumerate(entries):
            if int(Ks[n]) > 0:
                groups.setdefault((tuple(e[1]), tuple(e[2])), []).append(n)
        for (in_size, orig_size), imgs in groups.items():
            if len(groups) == 1:
                sel = None
                lo = low
            else:
                sel = torch.cat([torch.arange(int(first[n]), int(first[n] + Ks[n])) for n in imgs]).to(dev)
                lo = low[sel]
            native, sm = upscale_masks(lo, in_size, orig_size, head.img_enc_img_size, head.mask_threshold,
                                       small_size=(H, W))
            self.last_native_masks.append((sel, native))
            if sel is None:
                small = sm
            else:
                small[sel] = sm
        seg[img_of.to(dev), cls_of.to(dev)] = small[:, 0]
        s_ = score[:, 0].float().cpu()
        est_dice[img_of, cls_of] = 2 * s_ / (1 + s_)
        return seg, est_dice
