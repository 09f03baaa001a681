"""Mask statistics: the oracle restatement against golden vectors produced by the reference's own
segment_anything/utils/amg.py (tests/golden/make_golden_amg.py)."""
from pathlib import Path

import numpy as np

from oracle import sam_oracle as O

GOLD = np.load(Path(__file__).parent / "golden" / "amg_golden.npz")


def test_stability_score_matches_reference():
    for k in range(3):
        thr, off = GOLD[f"args{k}"]
        got = O.stability_score(GOLD["logits"], float(thr), float(off))
        want = GOLD[f"score{k}"]
        assert got.dtype == np.float32 and got.shape == want.shape
        assert np.array_equal(got, want, equal_nan=True)  # integer counts, one fp32 division: bit-exact
    assert np.isnan(GOLD["score0"][0, 0]) and GOLD["score0"][0, 1] == 0.0  # the edge cases are in the fixture


def test_mask_to_box_matches_reference():
    got = O.mask_to_box(GOLD["masks"])
    assert got.dtype == np.int64 and np.array_equal(got, GOLD["boxes"])
    assert np.array_equal(O.mask_to_box(GOLD["masks"][0, 1]), GOLD["boxes_2d"])
    assert np.array_equal(GOLD["boxes"][1, 2], [0, 0, 0, 0]) and np.array_equal(GOLD["boxes"][2, 3], [52, 36, 52, 36])
