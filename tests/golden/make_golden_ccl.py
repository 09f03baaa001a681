"""Golden vectors for the connected-component pre-processing step, generated in the build container:

    python tests/golden/make_golden_ccl.py      (needs /root/reference; writes tests/golden/ccl_golden.npz)

The reference's own `remove_all_but_one_connected_component` (utils/segmentation_preprocessing.py:7-52) is imported
and run unmodified.  Its only missing dependencies are stubbed: `kornia.contrib.connected_components` by a torch
restatement of kornia 0.7.0's published algorithm (num_iterations rounds of 3x3 max-pooling of the batch-global pixel
indices inside the mask), `kornia.morphology` / `skimage.morphology` by empty modules (not reached)."""
import sys
import types
from pathlib import Path

import numpy as np
import torch
import torch.nn.functional as F

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import sam_oracle as O  # noqa: E402  (synthetic inputs)


def connected_components(image: torch.Tensor, num_iterations: int = 100) -> torch.Tensor:
    H, W = image.shape[-2:]
    image_view = image.view(-1, 1, H, W)
    mask = image_view == 1
    B = image_view.shape[0]
    out = torch.arange(B * H * W, device=image.device, dtype=image.dtype).view((-1, 1, H, W))
    out[~mask] = 0
    for _ in range(num_iterations):
        out[mask] = F.max_pool2d(out, kernel_size=3, stride=1, padding=1)[mask]
    return out.view_as(image)


def main():
    kornia = types.ModuleType("kornia")
    contrib = types.ModuleType("kornia.contrib")
    contrib.connected_components = connected_components
    morph = types.ModuleType("kornia.morphology")
    morph.erosion = morph.dilation = None
    sk = types.ModuleType("skimage")
    skm = types.ModuleType("skimage.morphology")
    skm.disk = skm.square = skm.diamond = skm.star = None
    sys.modules.update({"kornia": kornia, "kornia.contrib": contrib, "kornia.morphology": morph, "skimage": sk,
                        "skimage.morphology": skm})
    sys.path.insert(0, "/root/reference")
    from utils.segmentation_preprocessing import remove_all_but_one_connected_component as ref_fn

    out = {}
    # (seed, C, H, W): two full-size maps (17 x 384 x 224, num_iter = 384 like SegEnhance) and two small ones
    for seed, C, H, W in [(0, 17, 384, 224), (1, 17, 384, 224), (2, 5, 96, 64), (3, 3, 40, 72)]:
        prob = O.synthetic_unet_probs(seed, C, H, W)
        for sel in ("highest_probability", "largest"):
            with torch.inference_mode():
                ref = ref_fn(torch.from_numpy(prob), sel, num_iter=max(H, W)).numpy()
            nz = np.flatnonzero(ref)
            out[f"s{seed}_{sel}_idx"] = nz.astype(np.int32)
            out[f"s{seed}_{sel}_val"] = ref.ravel()[nz]
        out[f"s{seed}_shape"] = np.array([C, H, W])
    np.savez_compressed(ROOT / "tests" / "golden" / "ccl_golden.npz", **out)
    print("wrote ccl_golden.npz", {k: v.shape for k, v in out.items() if k.endswith("idx")})


if __name__ == "__main__":
    main()
