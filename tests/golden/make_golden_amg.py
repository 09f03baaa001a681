"""Golden vectors for the mask statistics (run in the build container): the reference's own
segment_anything/utils/amg.py `calculate_stability_score` and `batched_mask_to_box` on seeded inputs.

    python tests/golden/make_golden_amg.py   ->  tests/golden/amg_golden.npz"""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, "/root/reference")
from segment_anything.utils.amg import batched_mask_to_box, calculate_stability_score  # noqa: E402

HERE = Path(__file__).resolve().parent


def main():
    rng = np.random.default_rng(42)
    logits = rng.normal(0.0, 2.0, size=(3, 4, 37, 53)).astype(np.float32)
    logits[0, 0] = -10.0          # empty at both thresholds -> 0 / 0 = NaN
    logits[0, 1, 5, 7] = 0.5      # inside the low mask only -> 0
    logits[0, 1][logits[0, 1] != 0.5] = -10.0
    out = {"logits": logits}
    for k, (thr, off) in enumerate([(0.0, 1.0), (0.3, 0.7), (-0.2, 0.05)]):
        out[f"score{k}"] = calculate_stability_score(torch.from_numpy(logits), thr, off).numpy()
        out[f"args{k}"] = np.array([thr, off], np.float64)
    masks = logits > 1.5
    masks[1, 2] = False           # empty mask -> zeros
    masks[2, 3] = False
    masks[2, 3, 36, 52] = True    # single pixel in the corner
    out["masks"] = masks
    out["boxes"] = batched_mask_to_box(torch.from_numpy(masks)).numpy()
    out["boxes_2d"] = batched_mask_to_box(torch.from_numpy(masks[0, 1])).numpy()
    np.savez_compressed(HERE / "amg_golden.npz", **out)
    print({k: (v.shape, v.dtype) for k, v in out.items()})


if __name__ == "__main__":
    main()
