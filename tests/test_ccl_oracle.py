"""CPU: the CCL / component-selection oracle against the golden outputs of the reference's own function
(tests/golden/make_golden_ccl.py) and against hand-made edge cases."""
from pathlib import Path

import numpy as np
import pytest

from oracle import sam_oracle as O

GOLD = Path(__file__).parent / "golden" / "ccl_golden.npz"


def golden_case(g, seed, sel):
    C, H, W = (int(v) for v in g[f"s{seed}_shape"])
    ref = np.zeros(C * H * W, np.float32)
    ref[g[f"s{seed}_{sel}_idx"]] = g[f"s{seed}_{sel}_val"]
    return O.synthetic_unet_probs(seed, C, H, W), ref.reshape(C, H, W)


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
@pytest.mark.parametrize("sel", ["highest_probability", "largest"])
def test_oracle_matches_reference_function(seed, sel):
    g = np.load(GOLD)
    prob, ref = golden_case(g, seed, sel)
    got = O.remove_all_but_one_connected_component(prob, sel)
    assert np.array_equal(got, ref)  # bit-exact: values are copies of the input probabilities


def test_ccl_labels_edge_cases():
    m = np.zeros((2, 4, 5), bool)
    m[0, 0, 0] = True            # lone pixel at global index 0: label 0 == background in the reference
    m[0, 1, 2] = m[0, 2, 3] = True  # diagonal neighbours are one component (8-connectivity)
    m[1, 3, :] = True
    lbl = O.ccl_labels(m)
    assert lbl[0, 0, 0] == 0
    assert lbl[0, 1, 2] == lbl[0, 2, 3] == 2 * 5 + 3
    assert (lbl[1, 3] == 20 + 19).all()
    # empty class and single-component class
    prob = np.zeros((2, 4, 5), np.float32)
    prob[1, 1:3, 1:3] = 0.9
    out = O.remove_all_but_one_connected_component(prob, "highest_probability")
    assert np.array_equal(out, prob)
    # exactly 0.5 is not foreground (prob > 0.5)
    prob[0, 0, 1] = 0.5
    assert O.remove_all_but_one_connected_component(prob, "largest")[0].sum() == 0
