"""Decoder-only SAM predictor over precomputed embeddings (reference: segment_anything/sam_mask_decoder_head.py).

`predict_mask` keeps the reference's one-prompt contract; `predict_masks_batched` is the B200 path that
decodes ALL prompts of an image in one batched launch sequence (the reference re-reads and
gunzips the 4 MiB embedding and launches ~150 B=1 kernels per call)."""
from __future__ import annotations

from copy import deepcopy
from pathlib import Path
from typing import Dict, List, Sequence, Tuple, Union

import numpy as np
import torch

from .build_sam import sam_model_registry
from .modeling.sam import upscale_masks
from .utils.prompt_utils import Prompt, scale_box, scale_coords

KNOWN_PROMPTS = ["pos_points", "neg_points", "box"]


class _Attrs(dict):
    pass


class _Entry(dict):
    def __init__(self, features: torch.Tensor, original_size, input_size):
        super().__init__(features=features)
        self.attrs = {"original_size": np.asarray(original_size), "input_size": np.asarray(input_size)}


class EmbeddingStore:
    """In-HBM stand-in for the reference's `img_embedding` h5 group: embeddings stay resident on the device
    (what `generate_img_embeddings.py:67-70` writes per image: features + original_size + input_size attrs)."""

    def __init__(self, checkpoint_name: str = "", img_encoder_img_size: int = 1024):
        self.attrs = {"checkpoint": checkpoint_name, "img_encoder_img_size": img_encoder_img_size}
        self._entries: Dict[str, _Entry] = {}

    def add(self, name: str, features: torch.Tensor, original_size, input_size) -> None:
        self._entries[name] = _Entry(features, original_size, input_size)

    def __getitem__(self, name: str) -> _Entry:
        return self._entries[name]

    def __contains__(self, name: str) -> bool:
        return name in self._entries


class SAMMaskDecoderHead:
    def __init__(self, sam_checkpoint, model_type: str, device: str, img_embedding_h5, sam_model=None):
        """Same arguments as the reference (:13-35).  `img_embedding_h5` may be an h5 path (needs h5py) or an
        `EmbeddingStore`; `sam_model` optionally supplies an already-built `Sam` instead of loading a checkpoint."""
        self.device = torch.device(device)
        if isinstance(img_embedding_h5, EmbeddingStore):
            self.img_embedding = img_embedding_h5
            attrs = img_embedding_h5.attrs
        else:
            import h5py  # storage dependency of the reference; only needed for the on-disk path
            h5_file = h5py.File(Path(img_embedding_h5), "r")
            self.img_embedding = h5_file["img_embedding"]
            attrs = h5_file.attrs
        self.img_enc_img_size = int(attrs["img_encoder_img_size"])
        if sam_model is None:
            sam_checkpoint = Path(sam_checkpoint)
            assert attrs["checkpoint"] == sam_checkpoint.name, "SAM checkpoint mismatch"
            sam_model = sam_model_registry[model_type](checkpoint=sam_checkpoint)
        # `.to()` re-packs the weights of the CUDA engines: only move the model if it is not on the device already
        cur = sam_model.device
        same = cur.type == self.device.type and (
            self.device.index is None or cur.index == self.device.index or
            (cur.index is None and self.device.index == torch.cuda.current_device()))
        self.sam = sam_model if same else sam_model.to(device=self.device)
        self.prompt_encoder = self.sam.prompt_encoder
        self.mask_decoder = self.sam.mask_decoder
        self.mask_threshold = self.sam.mask_threshold
        self._feature_cache: Dict[str, torch.Tensor] = {}

    # ------------------------------------------------------------------------------------------------
    def _entry(self, img_name: str):
        ds = self.img_embedding[img_name]
        input_size = [int(v) for v in np.asarray(ds.attrs["input_size"]).tolist()]
        original_size = [int(v) for v in np.asarray(ds.attrs["original_size"]).tolist()]
        feats = self._feature_cache.get(img_name)
        if feats is None:
            f = ds["features"]
            feats = f if isinstance(f, torch.Tensor) else torch.from_numpy(np.asarray(f[:]))
            feats = feats.to(self.device, non_blocking=True)
            self._feature_cache = {img_name: feats}  # keep the current image's embedding resident in HBM
        return feats, input_size, original_size

    def _gather(self, prompts: Sequence[Prompt], prompt2use: Sequence[str], input_size):
        pts, labs, boxes = [], [], None
        if "pos_points" in prompt2use:
            assert all(p.pos_seeds is not None for p in prompts), "pos_seeds are not available"
            pos = torch.stack([scale_coords(p.pos_seeds, p.img_size, input_size) for p in prompts])
            pts.append(pos)
            labs.append(torch.ones(pos.shape[:2], dtype=torch.int32, device=pos.device))
        if "neg_points" in prompt2use:
            assert all(p.neg_seeds is not None for p in prompts), "neg_seeds are not available"
            neg = torch.stack([scale_coords(p.neg_seeds, p.img_size, input_size) for p in prompts])
            pts.append(neg)
            labs.append(torch.zeros(neg.shape[:2], dtype=torch.int32, device=neg.device))
        if "box" in prompt2use:
            assert all(p.box is not None for p in prompts), "box is not available"
            boxes = torch.cat([scale_box(p.box.unsqueeze(0), p.img_size, input_size) for p in prompts]).float()
        points = torch.cat(pts, dim=1).float() if pts else None
        labels = torch.cat(labs, dim=1) if labs else None
        return points, labels, boxes

    @torch.inference_mode()
    def predict_masks_batched(self, img_name: str, prompts: Sequence[Prompt], prompt2use: Union[str, List[str]],
                              mask_prev_iter: torch.Tensor = None, upscale: bool = True, small_size=None):
        """All prompts of one image in one batched decode.  Returns (masks [K,1,H0,W0] bool | None, iou [K,1],
        low_res [K,1,256,256], small [K,1,h,w] bool | None)."""
        if isinstance(prompt2use, str):
            prompt2use = [prompt2use]
        assert all(p in KNOWN_PROMPTS for p in prompt2use), f"Prompt must be one of {KNOWN_PROMPTS}"
        feats, input_size, original_size = self._entry(img_name)
        points, labels, boxes = self._gather(prompts, prompt2use, input_size)
        low, iou = self.sam.decode_prompts(feats, points, labels, boxes, mask_prev_iter, multimask_output=False)
        masks = small = None
        if upscale:
            res = upscale_masks(low, input_size, original_size, self.img_enc_img_size, self.mask_threshold,
                                small_size=small_size)
            masks, small = res if small_size is not None else (res, None)
        return masks, iou, low, small

    @torch.inference_mode()
    def predict_mask(self, img_name: str, given_prompt: Prompt, prompt2use: Union[str, List[str]],
                     mask_prev_iter: torch.Tensor = None) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        """One prompt -> (bool mask 1x1xH0xW0, iou 1x1, low-res logits 1x1x256x256) (reference :37-104)."""
        masks, iou, low, _ = self.predict_masks_batched(img_name, [deepcopy(given_prompt)], prompt2use, mask_prev_iter)
        return masks, iou, low

    def postprocess_masks(self, masks: torch.Tensor, input_size: Tuple[int, ...],
                          original_size: Tuple[int, ...]) -> torch.Tensor:
        """Remove padding and upscale to the original size (reference :106-135) — one fused kernel."""
        return upscale_masks(masks, input_size, original_size, self.img_enc_img_size, return_logits=True)
