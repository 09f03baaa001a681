"""`SamPredictor` (reference: segment_anything/predictor.py) on top of the fused CUDA paths of `Sam`."""
from typing import Optional, Tuple

import numpy as np
import torch

from .modeling import Sam
from .modeling.sam import upscale_masks
from .utils.transforms import ResizeLongestSide


class SamPredictor:
    def __init__(self, sam_model: Sam) -> None:
        super().__init__()
        self.model = sam_model
        self.transform = ResizeLongestSide(sam_model.image_encoder.img_size)
        self.reset_image()

    def set_image(self, image: np.ndarray, image_format: str = "RGB") -> None:
        """HWC uint8 image -> embedding (reference predictor.py:34-60)."""
        assert image_format in ["RGB", "BGR"], f"image_format must be in ['RGB', 'BGR'], is {image_format}."
        if image_format != self.model.image_format:
            image = image[..., ::-1]
        # ResizeLongestSide.apply_image on the GPU (Pillow-exact fixed-point resample, csrc/resize.cu): the native
        # uint8 pixels are uploaded once and resized / transposed to CHW there (bit-identical to the host path)
        input_image_torch = self.transform.apply_image_cuda(np.ascontiguousarray(image), device=self.device)[None]
        self.set_torch_image(input_image_torch, image.shape[:2])

    @torch.no_grad()
    def set_torch_image(self, transformed_image: torch.Tensor, original_image_size: Tuple[int, ...]) -> None:
        """1x3xHxW transformed image -> embedding (reference predictor.py:62-90); preprocess is fused in."""
        assert (len(transformed_image.shape) == 4 and transformed_image.shape[1] == 3
                and max(*transformed_image.shape[2:]) == self.model.image_encoder.img_size), \
            f"set_torch_image input must be BCHW with long side {self.model.image_encoder.img_size}."
        self.reset_image()
        self.original_size = original_image_size
        self.input_size = tuple(transformed_image.shape[-2:])
        self.features = self.model.encode_image(transformed_image)
        self.is_image_set = True

    def predict(self, point_coords: Optional[np.ndarray] = None, point_labels: Optional[np.ndarray] = None,
                box: Optional[np.ndarray] = None, mask_input: Optional[np.ndarray] = None,
                multimask_output: bool = True, return_logits: bool = False
                ) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
        """numpy prompts -> (masks CxHxW, iou C, low_res Cx256x256) (reference predictor.py:92-166)."""
        if not self.is_image_set:
            raise RuntimeError("An image must be set with .set_image(...) before mask prediction.")
        coords_torch, labels_torch, box_torch, mask_input_torch = None, None, None, None
        if point_coords is not None:
            assert point_labels is not None, "point_labels must be supplied if point_coords is supplied."
            point_coords = self.transform.apply_coords(point_coords, self.original_size)
            coords_torch = torch.as_tensor(point_coords, dtype=torch.float, device=self.device)
            labels_torch = torch.as_tensor(point_labels, dtype=torch.int, device=self.device)
            coords_torch, labels_torch = coords_torch[None, :, :], labels_torch[None, :]
        if box is not None:
            box = self.transform.apply_boxes(box, self.original_size)
            box_torch = torch.as_tensor(box, dtype=torch.float, device=self.device)
            box_torch = box_torch[None, :]
        if mask_input is not None:
            mask_input_torch = torch.as_tensor(mask_input, dtype=torch.float, device=self.device)
            mask_input_torch = mask_input_torch[None, :, :, :]
        masks, iou_predictions, low_res_masks = self.predict_torch(
            coords_torch, labels_torch, box_torch, mask_input_torch, multimask_output, return_logits=return_logits)
        return (masks[0].detach().cpu().numpy(), iou_predictions[0].detach().cpu().numpy(),
                low_res_masks[0].detach().cpu().numpy())

    @torch.no_grad()
    def predict_torch(self, point_coords: Optional[torch.Tensor], point_labels: Optional[torch.Tensor],
                      boxes: Optional[torch.Tensor] = None, mask_input: Optional[torch.Tensor] = None,
                      multimask_output: bool = True, return_logits: bool = False
                      ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        """Batched torch prompts in the input frame (reference predictor.py:168-243)."""
        if not self.is_image_set:
            raise RuntimeError("An image must be set with .set_image(...) before mask prediction.")
        if boxes is not None and boxes.dim() == 3:
            boxes = boxes.reshape(boxes.shape[0], 4)
        low_res_masks, iou_predictions = self.model.decode_prompts(
            self.features, point_coords, point_labels, boxes, mask_input, multimask_output)
        masks = upscale_masks(low_res_masks, self.input_size, self.original_size, self.model.image_encoder.img_size,
                              self.model.mask_threshold, return_logits=return_logits)
        return masks, iou_predictions, low_res_masks

    def get_image_embedding(self) -> torch.Tensor:
        if not self.is_image_set:
            raise RuntimeError("An image must be set with .set_image(...) to generate an embedding.")
        assert self.features is not None, "Features must exist if an image has been set."
        return self.features

    @property
    def device(self) -> torch.device:
        return self.model.device

    def reset_image(self) -> None:
        self.is_image_set = False
        self.features = None
        self.orig_h = None
        self.orig_w = None
        self.input_h = None
        self.input_w = None
