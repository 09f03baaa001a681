"""Golden vectors for the image-ingest resize (run in the build container; Pillow is the library the reference's
ResizeLongestSide.apply_image resolves to through torchvision, utils/transforms.py:26-31).

    python tests/golden/make_golden_resize.py   ->  tests/golden/resize_golden.npz

Small seeded uint8 images are resized with `PIL.Image.resize(..., BILINEAR)` exactly as the reference does (and, where
/root/reference is present, through the reference's own ResizeLongestSide.apply_image as a cross-check)."""
import sys
from pathlib import Path

import numpy as np
from PIL import Image

HERE = Path(__file__).resolve().parent
CASES = [  # (H, W, C, out_h, out_w)
    (61, 97, 3, 40, 64),      # down-scale both axes
    (50, 37, 3, 69, 51),      # up-scale both axes
    (118, 75, 1, 64, 41),     # gray, CVAT-like aspect (1182 x 754 / 10)
    (64, 90, 3, 64, 45),      # width only
    (33, 20, 3, 100, 20),     # height only
    (257, 204, 3, 128, 102),  # 2570 x 2040 / 10 -> long side 128
]


def main():
    out = {}
    for i, (H, W, C, oh, ow) in enumerate(CASES):
        rng = np.random.default_rng(1000 + i)
        img = rng.integers(0, 256, size=(H, W, C), dtype=np.uint8)
        if i == 0:
            img[:8] = 255; img[8:16] = 0  # saturated bands
        pil = Image.fromarray(img[..., 0] if C == 1 else img)
        res = np.array(pil.resize((ow, oh), resample=Image.BILINEAR))
        out[f"in{i}"] = img
        out[f"out{i}"] = res.reshape(oh, ow, C)
    # the reference's own entry point on a long-side case (needs torchvision; skipped where the reference is absent)
    ref_root = Path("/root/reference")
    if ref_root.exists():
        sys.path.insert(0, str(ref_root))
        from segment_anything.utils.transforms import ResizeLongestSide
        rng = np.random.default_rng(77)
        img = rng.integers(0, 256, size=(237, 151, 3), dtype=np.uint8)
        out["ref_in"] = img
        out["ref_out"] = ResizeLongestSide(128).apply_image(img)
    np.savez_compressed(HERE / "resize_golden.npz", **out)
    print("wrote", HERE / "resize_golden.npz", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
