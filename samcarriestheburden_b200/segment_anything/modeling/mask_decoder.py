"""Mask decoder parameter container (reference: segment_anything/modeling/mask_decoder.py)."""
from __future__ import annotations

from typing import Tuple, Type

import torch
import torch.nn as nn

from .common import FusedAway, LayerNorm2d


class MLP(FusedAway):
    def __init__(self, input_dim: int, hidden_dim: int, output_dim: int, num_layers: int,
                 sigmoid_output: bool = False) -> None:
        super().__init__()
        self.num_layers = num_layers
        h = [hidden_dim] * (num_layers - 1)
        self.layers = nn.ModuleList(nn.Linear(n, k) for n, k in zip([input_dim] + h, h + [output_dim]))
        self.sigmoid_output = sigmoid_output


class MaskDecoder(nn.Module):
    def __init__(self, *, transformer_dim: int, transformer: nn.Module, num_multimask_outputs: int = 3,
                 activation: Type[nn.Module] = nn.GELU, iou_head_depth: int = 3, iou_head_hidden_dim: int = 256) -> None:
        super().__init__()
        if (transformer_dim, num_multimask_outputs, iou_head_depth, iou_head_hidden_dim) != (256, 3, 3, 256) \
                or activation is not nn.GELU:
            raise NotImplementedError("b200sam implements SAM's mask decoder only: dim 256, 3 multimask outputs, "
                                      "3-layer IoU head of width 256 (build_sam.py:86-97)")
        self.transformer_dim = transformer_dim
        self.transformer = transformer
        self.num_multimask_outputs = num_multimask_outputs
        self.iou_token = nn.Embedding(1, transformer_dim)
        self.num_mask_tokens = num_multimask_outputs + 1
        self.mask_tokens = nn.Embedding(self.num_mask_tokens, transformer_dim)
        self.output_upscaling = nn.Sequential(
            nn.ConvTranspose2d(transformer_dim, transformer_dim // 4, kernel_size=2, stride=2),
            LayerNorm2d(transformer_dim // 4), activation(),
            nn.ConvTranspose2d(transformer_dim // 4, transformer_dim // 8, kernel_size=2, stride=2), activation())
        self.output_hypernetworks_mlps = nn.ModuleList(
            [MLP(transformer_dim, transformer_dim, transformer_dim // 8, 3) for _ in range(self.num_mask_tokens)])
        self.iou_prediction_head = MLP(transformer_dim, iou_head_hidden_dim, self.num_mask_tokens, iou_head_depth)

        self._owner = None  # set by Sam: the CUDA decoder engine lives there

    @torch.no_grad()
    def forward(self, image_embeddings: torch.Tensor, image_pe: torch.Tensor, sparse_prompt_embeddings: torch.Tensor,
                dense_prompt_embeddings: torch.Tensor, multimask_output: bool) -> Tuple[torch.Tensor, torch.Tensor]:
        """Standalone mask prediction from prompt EMBEDDINGS (reference mask_decoder.py:71-110) -> (masks B x (1|3) x 256
        x 256, iou B x (1|3)).  `image_pe` must be the model's own dense positional encoding
        (`prompt_encoder.get_dense_pe()`, what every caller in the reference passes): its projections are folded into
        per-token tables when the engine is built."""
        if self._owner is None or self._owner() is None:
            raise RuntimeError("MaskDecoder needs the owning Sam model on a CUDA device")
        eng = self._owner().decoder_engine()
        own_pe = eng.dense_pe()
        if image_pe.data_ptr() != own_pe.data_ptr() and not torch.equal(image_pe.to(own_pe.device).expand_as(own_pe), own_pe):
            raise NotImplementedError("b200sam's MaskDecoder only supports image_pe == prompt_encoder.get_dense_pe()")
        return eng.decode_embedded(image_embeddings.to(eng.device), sparse_prompt_embeddings.to(eng.device),
                                   dense_prompt_embeddings.to(eng.device), bool(multimask_output))
