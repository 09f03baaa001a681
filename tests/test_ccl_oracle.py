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


def _kornia_published_ccl(mask: np.ndarray, num_iterations: int) -> np.ndarray:
    """kornia 0.7.0 `kornia.contrib.connected_components` as published (environment.yml:85; call site
    utils/segmentation_preprocessing.py:23): `num_iterations` rounds of 3x3 max-pooling of the batch-global pixel indices
    inside the mask.  Restated here (kornia is not installable offline) to pin the LABELLING the oracle derives from
    scipy.ndimage.label: after convergence both must give the same labels, not only the same partition."""
    import torch
    import torch.nn.functional as F
    C, H, W = mask.shape
    m = torch.from_numpy(mask).view(C, 1, H, W)
    out = torch.arange(C * H * W, dtype=torch.float32).view(C, 1, H, W)
    out[~m] = 0
    for _ in range(num_iterations):
        out[m] = F.max_pool2d(out, kernel_size=3, stride=1, padding=1)[m]
    return out.view(C, H, W).numpy().astype(np.int64)


@pytest.mark.parametrize("seed", range(6))
def test_ccl_labelling_pinned_to_scipy_and_published_kornia_algorithm(seed):
    """Property test: on random masks - blobs, salt noise, thin diagonal chains, touching corners - the
    oracle's labels (scipy.ndimage.label with the 8-connected 3x3 structure + per-component maximum of the global index)
    are (a) a relabelling of scipy's own partition and (b) IDENTICAL to the published kornia iteration run to convergence."""
    from scipy import ndimage
    rng = np.random.default_rng(seed)
    C, H, W = 3, 24 + seed, 31 - seed
    kind = seed % 3
    if kind == 0:
        mask = rng.random((C, H, W)) < 0.45                      # salt noise: many small 8-connected components
    elif kind == 1:
        mask = ndimage.binary_dilation(rng.random((C, H, W)) < 0.03, iterations=2, structure=np.ones((1, 3, 3), bool))
    else:
        mask = np.zeros((C, H, W), bool)
        for c in range(C):                                       # diagonal chains: only 8-connectivity joins them
            for k in range(min(H, W) - 1):
                mask[c, k, (k + c) % W] = True
            mask[c, H - 1, ::2] = True
    lbl = O.ccl_labels(mask)
    assert ((lbl > 0) | ~mask | (np.arange(mask.size).reshape(mask.shape) == 0)).all()
    for c in range(C):
        sc, n = ndimage.label(mask[c], structure=np.ones((3, 3), bool))
        pairs = {(int(a), int(b)) for a, b in zip(sc[mask[c]].ravel(), lbl[c][mask[c]].ravel())}
        fwd, bwd = {}, {}
        for a, b in pairs:                                       # bijection between scipy's ids and the oracle's labels
            if (c, a) == (0, int(sc[0, 0])) and b == 0:
                continue                                          # the reference's label-0 artefact (lone pixel 0)
            assert fwd.setdefault(a, b) == b and bwd.setdefault(b, a) == a
    conv = _kornia_published_ccl(mask, num_iterations=H * W)     # geodesic diameter <= H * W: converged
    assert np.array_equal(conv, lbl)
    # not yet converged after a few iterations on the chains (why the reference's num_iter matters for spirals)
    if kind == 2:
        assert not np.array_equal(_kornia_published_ccl(mask, num_iterations=3), lbl)
