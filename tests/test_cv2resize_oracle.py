"""CPU: the oracle's restatement of OpenCV's uint8 INTER_LINEAR resize (the U-Net ingest of
scripts/save_refined_segmentations.py:63) against golden outputs produced by cv2 itself
(tests/golden/make_golden_cv2resize.py), bit-exact."""
import importlib.util
from pathlib import Path

import numpy as np
import pytest

from oracle import sam_oracle as O

GOLD = Path(__file__).parent / "golden" / "cv2resize_golden.npz"
_spec = importlib.util.spec_from_file_location("make_golden_cv2resize", Path(__file__).parent / "golden" / "make_golden_cv2resize.py")


def golden_cases():
    g = np.load(GOLD)
    seeds = sorted(int(k.split("_")[1]) for k in g.files if k.startswith("shape_"))
    return g, seeds


def make_image(seed, H, W):
    """Same generator as the golden script (kept in sync by test_generator_matches_script when cv2 is importable)."""
    rng = np.random.default_rng(1000 + seed)
    yy, xx = np.mgrid[0:H, 0:W].astype(np.float64)
    smooth = 127.5 + 127.5 * np.sin(yy / 37.0 + seed) * np.cos(xx / 23.0 - seed)
    return np.clip(0.5 * rng.integers(0, 256, (H, W)) + 0.5 * smooth, 0, 255).astype(np.uint8)


@pytest.mark.parametrize("seed", golden_cases()[1])
def test_oracle_resize_matches_cv2_golden(seed):
    g, _ = golden_cases()
    H, W = (int(v) for v in g[f"shape_{seed}"])
    got = O.cv2_resize_linear_u8(make_image(seed, H, W), 384, 224)
    assert np.array_equal(got, g[f"out_{seed}"]), (seed, H, W, int((got != g[f"out_{seed}"]).sum()))


def test_generator_matches_script_and_live_cv2():
    cv2 = pytest.importorskip("cv2")
    mod = importlib.util.module_from_spec(_spec)
    _spec.loader.exec_module(mod)
    assert np.array_equal(mod.image(4, 100, 100), make_image(4, 100, 100))
    rng = np.random.default_rng(7)
    for (H, W, dh, dw) in [(333, 517, 384, 224), (64, 64, 97, 131), (700, 100, 384, 224), (1, 9, 5, 3)]:
        img = rng.integers(0, 256, (H, W), dtype=np.uint8)
        assert np.array_equal(O.cv2_resize_linear_u8(img, dh, dw), cv2.resize(img, (dw, dh), interpolation=cv2.INTER_LINEAR))


def test_library_host_tables_match_oracle():
    """The C library's host coefficient builder (b200sam_cvresize_coeffs_host) == the oracle's tables, both axes."""
    import ctypes as C
    from samcarriestheburden_b200 import _lib
    lib = _lib.load()
    for (ssize, dsize) in [(754, 224), (1182, 384), (2040, 224), (2570, 384), (100, 224), (37, 384), (224, 224), (2, 384), (1, 5)]:
        for clamp in (0, 1):
            idx = np.zeros((dsize, 2), np.int32)
            w = np.zeros((dsize, 2), np.int32)
            assert lib.b200sam_cvresize_coeffs_host(ssize, dsize, clamp, idx.ctypes.data_as(C.c_void_p),
                                                    w.ctypes.data_as(C.c_void_p)) == 0
            i0, i1, ww = O.cv2_linear_coeffs(ssize, dsize, bool(clamp))
            assert np.array_equal(idx[:, 0], i0) and np.array_equal(idx[:, 1], i1), (ssize, dsize, clamp)
            assert np.array_equal(w, ww), (ssize, dsize, clamp)


@pytest.mark.parametrize("seed", [20, 21, 22])
def test_oracle_cubic_matches_cv2_golden(seed):
    """MedSAM ingest (scripts/generate_img_embeddings.py:49-53): the oracle's INTER_CUBIC restatement against goldens written
    by cv2 with Intel IPP switched off (OpenCV's own code path), bit-exact."""
    g, _ = golden_cases()
    H, W = (int(v) for v in g[f"cubic_shape_{seed}"])
    got = O.cv2_resize_cubic_u8(make_image(seed, H, W), 1024, 1024)
    assert np.array_equal(got[::8], g[f"cubic_rows_{seed}"])
    assert [int(got.min()), int(got.max())] == [int(v) for v in g[f"cubic_minmax_{seed}"]]


def test_library_cubic_tables_match_oracle():
    import ctypes as C
    from samcarriestheburden_b200 import _lib
    lib = _lib.load()
    for (ssize, dsize) in [(754, 1024), (1182, 1024), (2040, 1024), (200, 1024), (1024, 1024), (3, 1024), (2920, 1024)]:
        idx = np.zeros((dsize, 4), np.int32)
        w = np.zeros((dsize, 4), np.int32)
        assert lib.b200sam_cvresize_cubic_coeffs_host(ssize, dsize, idx.ctypes.data_as(C.c_void_p), w.ctypes.data_as(C.c_void_p)) == 0
        oi, ow = O.cv2_cubic_coeffs(ssize, dsize)
        assert np.array_equal(idx, oi) and np.array_equal(w, ow), (ssize, dsize)


def test_medsam_preprocess_oracle_against_live_cv2():
    cv2 = pytest.importorskip("cv2")
    if hasattr(cv2, "ipp"):
        cv2.ipp.setUseIPP(False)
    gray = make_image(31, 411, 263)
    rgb = cv2.cvtColor(gray, cv2.COLOR_GRAY2RGB)
    r = cv2.resize(rgb, (1024, 1024), interpolation=cv2.INTER_CUBIC)
    ref = (r - r.min()) / np.clip(r.max() - r.min(), a_min=1e-8, a_max=None)      # generate_img_embeddings.py:55-56
    import torch
    ref_t = torch.tensor(ref).float().permute(2, 0, 1).unsqueeze(0)               # :60
    assert torch.equal(O.medsam_preprocess(gray), ref_t)
