"""Live check of the oracle against the reference code itself (only where /root/reference exists, i.e. the
build container).  Full-size ViT-B + decoder + the complete two-pass refinement loop."""
import sys
import types
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import sam_oracle as O

REF = Path("/root/reference")
pytestmark = pytest.mark.skipif(not REF.exists(), reason="reference checkout not present (GPU box)")


def _import_reference():
    if str(REF) not in sys.path:
        sys.path.insert(0, str(REF))
    for name in ("h5py", "kornia", "kornia.contrib", "kornia.morphology", "skimage", "skimage.morphology", "pyamg"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["kornia.contrib"].connected_components = None
    for fn in ("dilation", "erosion"):
        setattr(sys.modules["kornia.morphology"], fn, None)
    for fn in ("square", "disk", "diamond", "star"):
        setattr(sys.modules["skimage.morphology"], fn, None)
    import segment_anything  # noqa: F401
    return segment_anything


class _Attrs(dict):
    pass


class _FakeH5:
    """In-memory stand-in for the embeddings h5 file (storage only, no arithmetic)."""

    def __init__(self, feats, original_size, input_size):
        grp = {"features": feats}

        class G(dict):
            attrs = {"original_size": np.array(original_size), "input_size": np.array(input_size)}

        g = G(grp)
        self._d = {"img_embedding": {"img": g}}
        self.attrs = {"img_encoder_img_size": 1024, "checkpoint": "ckpt.pth"}

    def __getitem__(self, k):
        return self._d[k]


@pytest.mark.slow
def test_vit_b_end_to_end_against_reference():
    sa = _import_reference()
    from segment_anything import SamPredictor, sam_model_registry
    torch.manual_seed(0)
    sd = O.random_state_dict("vit_b", seed=0)
    sam = sam_model_registry["vit_b"]()
    sam.load_state_dict(sd, strict=True)  # names/shapes identical to the reference (build_sam.py:102-107)
    pred = SamPredictor(sam)
    img = O.synthetic_radiograph(3, 754, 589)
    pred.set_image(img)
    ref_feats = pred.features
    # oracle encoder on the same resized image
    from segment_anything.utils.transforms import ResizeLongestSide
    resized = ResizeLongestSide(1024).apply_image(img)
    assert resized.shape[:2] == O.get_preprocess_shape(754, 589)
    x = O.preprocess(torch.from_numpy(resized).permute(2, 0, 1).float())[None]
    feats = O.image_encoder(sd, x, **O.VIT_CONFIGS["vit_b"])
    rel = float((feats - ref_feats).norm() / ref_feats.norm())
    assert rel < 1e-5, rel

    # SamPredictor.predict with a box (config 1)
    box = np.array([100.0, 150.0, 400.0, 600.0])
    m_ref, iou_ref, low_ref = pred.predict(box=box, multimask_output=False)
    bt = torch.from_numpy(box[None] * (np.array(pred.input_size[::-1] * 2) / np.array([589, 754] * 2))).float()
    sp, de = O.prompt_encoder(sd, None, bt, None)
    low, iou = O.mask_decoder(sd, feats, O.dense_pe(sd), sp, de, False)
    assert np.allclose(low[0].numpy(), low_ref, atol=2e-4), np.abs(low[0].numpy() - low_ref).max()
    m = (O.postprocess_masks(low, pred.input_size, (754, 589)) > 0)[0].numpy()
    assert (m != m_ref).sum() <= 5

    # full two-pass refinement through the reference's SAMSegRefiner with an in-memory h5 stand-in
    import segment_anything.sam_mask_decoder_head as head
    from utils.seg_refinement import SAMSegRefiner
    fake = _FakeH5(ref_feats.numpy(), (754, 589), pred.input_size)
    sys.modules["h5py"].File = lambda *a, **k: fake
    head.sam_model_registry = {"vit_h": lambda checkpoint=None: sam}
    orig_init = SAMSegRefiner.__init__
    refiner = SAMSegRefiner.__new__(SAMSegRefiner)
    refiner.sam_predictor = head.SAMMaskDecoderHead("ckpt.pth", "vit_h", "cpu", Path("x.h5"))
    refiner.prompts2use1st, refiner.prompts2use2nd, refiner.self_refine = ["box"], ["pos_points", "neg_points"], True
    seg = O.synthetic_unet_masks(7)
    ref_seg, ref_dice = refiner.refine(torch.from_numpy(seg.copy()), "img")
    got_seg, got_dice, _, _ = O.refine(sd, ref_feats, seg, pred.input_size, (754, 589))
    mism = int((got_seg != ref_seg.numpy()).sum())
    assert mism <= 10, mism
    assert np.allclose(got_dice, ref_dice.numpy(), atol=1e-4, equal_nan=True)
