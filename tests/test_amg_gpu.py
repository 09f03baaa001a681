"""GPU mask statistics through the Python mirror of segment_anything/utils/amg.py against the oracle and the
reference goldens.  Integer counts / coordinates: bit-exact."""
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import sam_oracle as O

pytestmark = pytest.mark.gpu
GOLD = np.load(Path(__file__).parent / "golden" / "amg_golden.npz")


def test_stability_score_golden_and_full_size():
    from samcarriestheburden_b200.segment_anything.utils.amg import calculate_stability_score
    x = torch.from_numpy(GOLD["logits"]).cuda()
    for k in range(3):
        thr, off = GOLD[f"args{k}"]
        got = calculate_stability_score(x, float(thr), float(off)).cpu().numpy()
        assert got.shape == (3, 4) and np.array_equal(got, GOLD[f"score{k}"], equal_nan=True)
    g = torch.Generator().manual_seed(0)
    big = torch.randn((5, 1024, 1024), generator=g)
    got = calculate_stability_score(big.cuda(), 0.0, 1.0).cpu().numpy()
    assert np.array_equal(got, O.stability_score(big.numpy(), 0.0, 1.0))
    assert calculate_stability_score(big[:0].cuda(), 0.0, 1.0).shape == (0,)


def test_mask_to_box_golden_shapes_and_empty():
    from samcarriestheburden_b200.segment_anything.utils.amg import batched_mask_to_box
    m = torch.from_numpy(GOLD["masks"]).cuda()
    got = batched_mask_to_box(m)
    assert got.dtype == torch.int64 and np.array_equal(got.cpu().numpy(), GOLD["boxes"])
    assert np.array_equal(batched_mask_to_box(m[0, 1]).cpu().numpy(), GOLD["boxes_2d"])  # 2-D input -> [4]
    g = torch.Generator().manual_seed(1)
    big = torch.rand((3, 1182, 754), generator=g) > 0.9999
    assert np.array_equal(batched_mask_to_box(big.cuda()).cpu().numpy(), O.mask_to_box(big.numpy()))
    empty = batched_mask_to_box(torch.zeros((0, 4, 8, 8), dtype=torch.bool, device="cuda"))
    assert empty.shape == (0, 4, 4) and empty.dtype == torch.float32  # the reference returns float zeros here


def test_predict_masks_feed_the_statistics():
    """The upstream AMG recipe on SamPredictor outputs: logits -> stability score, masks -> boxes."""
    from samcarriestheburden_b200.segment_anything import SamPredictor, sam_model_registry
    from samcarriestheburden_b200.segment_anything.utils.amg import batched_mask_to_box, calculate_stability_score
    sam = sam_model_registry["vit_b"]()
    sam.load_state_dict(O.random_state_dict("vit_b", seed=0), strict=True)
    sam = sam.to("cuda")
    pred = SamPredictor(sam)
    pred.set_image(O.synthetic_radiograph(2, 512, 384))
    pts = torch.tensor([[[200.0, 300.0]], [[100.0, 100.0]]], device="cuda")
    lab = torch.ones((2, 1), dtype=torch.int, device="cuda")
    logits, _, _ = pred.predict_torch(pts, lab, multimask_output=True, return_logits=True)
    score = calculate_stability_score(logits, sam.mask_threshold, 1.0)
    boxes = batched_mask_to_box(logits > sam.mask_threshold)
    assert score.shape == (2, 3) and boxes.shape == (2, 3, 4)
    assert np.array_equal(score.cpu().numpy(), O.stability_score(logits.cpu().numpy(), sam.mask_threshold, 1.0),
                          equal_nan=True)
    assert np.array_equal(boxes.cpu().numpy(), O.mask_to_box((logits > sam.mask_threshold).cpu().numpy()))
