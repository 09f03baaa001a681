"""Drop-in for the reference's `segment_anything` package surface used by the refinement hot path
(reference segment_anything/__init__.py:7-15).  `SamAutomaticMaskGenerator` is out of scope."""
from .build_sam import build_sam, build_sam_vit_b, build_sam_vit_h, build_sam_vit_l, sam_model_registry
from .predictor import SamPredictor
