"""Prompt encoder parameters + dense PE (reference: segment_anything/modeling/prompt_encoder.py).
Sparse/dense prompt embedding is fused into b200sam_decode (csrc/decoder.cu, decoder_ops.cu)."""
from __future__ import annotations

from typing import Optional, Tuple, Type

import torch
import torch.nn as nn

from .common import FusedAway, LayerNorm2d


class PositionEmbeddingRandom(FusedAway):
    def __init__(self, num_pos_feats: int = 64, scale: Optional[float] = None) -> None:
        super().__init__()
        if scale is None or scale <= 0.0:
            scale = 1.0
        self.register_buffer("positional_encoding_gaussian_matrix", scale * torch.randn((2, num_pos_feats)))


class PromptEncoder(nn.Module):
    def __init__(self, embed_dim: int, image_embedding_size: Tuple[int, int], input_image_size: Tuple[int, int],
                 mask_in_chans: int, activation: Type[nn.Module] = nn.GELU) -> None:
        super().__init__()
        if embed_dim != 256 or tuple(image_embedding_size) != (64, 64) or tuple(input_image_size) != (1024, 1024) \
                or mask_in_chans != 16 or activation is not nn.GELU:
            raise NotImplementedError("b200sam implements SAM's prompt encoder only: dim 256, 64x64 embedding, "
                                      "1024x1024 input, 16 mask channels (build_sam.py:81-85)")
        self.embed_dim = embed_dim
        self.input_image_size = input_image_size
        self.image_embedding_size = image_embedding_size
        self.pe_layer = PositionEmbeddingRandom(embed_dim // 2)
        self.num_point_embeddings: int = 4
        self.point_embeddings = nn.ModuleList([nn.Embedding(1, embed_dim) for _ in range(self.num_point_embeddings)])
        self.not_a_point_embed = nn.Embedding(1, embed_dim)
        self.mask_input_size = (4 * image_embedding_size[0], 4 * image_embedding_size[1])
        self.mask_downscaling = nn.Sequential(
            nn.Conv2d(1, mask_in_chans // 4, kernel_size=2, stride=2), LayerNorm2d(mask_in_chans // 4), activation(),
            nn.Conv2d(mask_in_chans // 4, mask_in_chans, kernel_size=2, stride=2), LayerNorm2d(mask_in_chans),
            activation(), nn.Conv2d(mask_in_chans, embed_dim, kernel_size=1))
        self.no_mask_embed = nn.Embedding(1, embed_dim)
        self._owner = None  # set by Sam: the fused decode path lives there

    def get_dense_pe(self) -> torch.Tensor:
        """1 x 256 x 64 x 64 dense positional encoding (reference prompt_encoder.py:62-71), computed by the
        CUDA dense-PE kernel when the decoder engine is created and returned here as an NCHW view."""
        if self._owner is None or self._owner() is None:
            raise RuntimeError("PromptEncoder.get_dense_pe needs the owning Sam model on a CUDA device")
        return self._owner().decoder_engine().dense_pe()

    def _engine(self):
        if self._owner is None or self._owner() is None:
            raise RuntimeError("PromptEncoder needs the owning Sam model on a CUDA device")
        return self._owner().decoder_engine()

    @torch.no_grad()
    def forward(self, points: Optional[Tuple[torch.Tensor, torch.Tensor]], boxes: Optional[torch.Tensor],
                masks: Optional[torch.Tensor]) -> Tuple[torch.Tensor, torch.Tensor]:
        """Standalone prompt embedding (reference prompt_encoder.py:128-168): points = (coords B x N x 2, labels B x N),
        boxes B x 4, masks B x 1 x 256 x 256 -> (sparse B x n x 256, dense B x 256 x 64 x 64).  The pipeline itself uses
        the fused Sam.decode_prompts; this entry point runs the same CUDA kernels for callers of the module API."""
        from .sam import assemble_prompt_points
        eng = self._engine()
        if points is not None:
            bs = points[0].shape[0]
        elif boxes is not None:
            bs = boxes.shape[0]
        elif masks is not None:
            bs = masks.shape[0]
        else:
            bs = 1
        c, l = assemble_prompt_points(points[0] if points is not None else None,
                                      points[1] if points is not None else None, boxes)
        c = c.to(eng.device) if c is not None else None
        l = l.to(eng.device) if l is not None else None
        m = masks.to(eng.device) if masks is not None else None
        return eng.prompt_encode(c, l, m, bs)
