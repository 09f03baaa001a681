"""Golden vectors for the U-Net ingest resize (scripts/save_refined_segmentations.py:63), generated in the build container
with OpenCV itself (cv2 is installed here; it is not on the GPU box's critical path and is never imported by the product):

    python tests/golden/make_golden_cv2resize.py        (writes tests/golden/cv2resize_golden.npz)

Seeded uint8 grey images at the CVAT native sizes (data/cvat_annotation_xml: 578x881 ... 2320x2920, H x W below) and a few
degenerate shapes -> cv2.resize(img, (224, 384), interpolation=cv2.INTER_LINEAR).  Inputs are regenerated from the seed by
the tests (numpy only); only the outputs are stored."""
from pathlib import Path

import cv2
import numpy as np

ROOT = Path(__file__).resolve().parents[2]
CASES = [(0, 1182, 754), (1, 881, 578), (2, 2570, 2040), (3, 384, 224), (4, 100, 100), (5, 37, 53), (6, 2920, 2320),
         (7, 385, 225), (8, 50, 400), (9, 2, 2), (10, 1098, 646)]


CUBIC_CASES = [(20, 1182, 754), (21, 881, 578), (22, 300, 200)]


def image(seed: int, H: int, W: int) -> np.ndarray:
    """Seeded uint8 test image from numpy alone (the tests regenerate it): uniform noise + a smooth low-frequency pattern."""
    rng = np.random.default_rng(1000 + seed)
    yy, xx = np.mgrid[0:H, 0:W].astype(np.float64)
    smooth = 127.5 + 127.5 * np.sin(yy / 37.0 + seed) * np.cos(xx / 23.0 - seed)
    return np.clip(0.5 * rng.integers(0, 256, (H, W)) + 0.5 * smooth, 0, 255).astype(np.uint8)


def main():
    out = {"cv2_version": np.array(cv2.__version__)}
    for seed, H, W in CASES:
        out[f"out_{seed}"] = cv2.resize(image(seed, H, W), (224, 384), interpolation=cv2.INTER_LINEAR)
        out[f"shape_{seed}"] = np.array([H, W])
    # MedSAM ingest (scripts/generate_img_embeddings.py:49-53): INTER_CUBIC to 1024 x 1024 with Intel IPP switched OFF, i.e.
    # OpenCV's own published code path (conda's opencv has no IPP; with IPP ~4 % of the pixels differ by 1 LSB).  Stored as
    # the uint8 result of three cases (rows subsampled 1:8 to keep the fixture small) + min / max of the full result.
    if hasattr(cv2, "ipp"):
        cv2.ipp.setUseIPP(False)
    for seed, H, W in CUBIC_CASES:
        full = cv2.resize(image(seed, H, W), (1024, 1024), interpolation=cv2.INTER_CUBIC)
        out[f"cubic_rows_{seed}"] = full[::8]
        out[f"cubic_minmax_{seed}"] = np.array([full.min(), full.max()])
        out[f"cubic_shape_{seed}"] = np.array([H, W])
    np.savez_compressed(ROOT / "tests" / "golden" / "cv2resize_golden.npz", **out)
    print("wrote cv2resize_golden.npz with", len(CASES), "cases")


if __name__ == "__main__":
    main()
