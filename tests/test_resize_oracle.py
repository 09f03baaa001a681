"""Image-ingest resize: the oracle restatement of Pillow's antialiased bilinear resample against the
golden vectors produced by Pillow / the reference's ResizeLongestSide.apply_image (tests/golden/make_golden_resize.py),
and the library's HOST coefficient builder against the oracle's (no GPU needed: b200sam_resize_coeffs_host is host code)."""
import ctypes as C
from pathlib import Path

import numpy as np
import pytest

from oracle import sam_oracle as O

GOLD = np.load(Path(__file__).parent / "golden" / "resize_golden.npz")
N_CASES = sum(1 for k in GOLD.files if k.startswith("in"))


@pytest.mark.parametrize("i", range(N_CASES))
def test_oracle_resize_matches_pillow_golden(i):
    img, want = GOLD[f"in{i}"], GOLD[f"out{i}"]
    got = O.resize_bilinear_u8(img, want.shape[0], want.shape[1])
    assert got.dtype == np.uint8 and got.shape == want.shape
    assert np.array_equal(got, want)  # bit-exact


def test_oracle_apply_image_matches_reference_entry_point():
    if "ref_in" not in GOLD.files:
        pytest.skip("golden file was generated without /root/reference")
    got = O.apply_image(GOLD["ref_in"], 128)
    assert np.array_equal(got, GOLD["ref_out"])


def test_oracle_resize_matches_live_pillow():
    Image = pytest.importorskip("PIL.Image")
    rng = np.random.default_rng(5)
    for (H, W, oh, ow) in [(45, 70, 31, 33), (20, 20, 47, 59), (96, 64, 32, 21)]:
        img = rng.integers(0, 256, size=(H, W, 3), dtype=np.uint8)
        want = np.array(Image.fromarray(img).resize((ow, oh), resample=Image.BILINEAR))
        assert np.array_equal(O.resize_bilinear_u8(img, oh, ow), want)


@pytest.mark.parametrize("in_size,out_size", [(97, 64), (37, 51), (1182, 1024), (754, 653), (2570, 1024), (2040, 813),
                                              (578, 672), (1024, 1024), (5, 1), (1, 7)])
def test_host_coefficients_match_oracle(in_size, out_size):
    from samcarriestheburden_b200 import _lib
    lib = _lib.load()
    ksize = lib.b200sam_resize_ksize(in_size, out_size)
    b_want, k_want = O.pil_resize_coeffs(in_size, out_size)
    assert ksize == k_want.shape[1]
    bounds = np.zeros((out_size, 2), np.int32)
    kk = np.zeros((out_size, ksize), np.int32)
    rc = lib.b200sam_resize_coeffs_host(in_size, out_size, bounds.ctypes.data_as(C.c_void_p), kk.ctypes.data_as(C.c_void_p))
    assert rc == 0
    assert np.array_equal(bounds, b_want) and np.array_equal(kk, k_want)
    # every row of weights sums to 2^22 within the per-tap rounding
    assert np.all(np.abs(kk.sum(1) - (1 << 22)) <= ksize)
