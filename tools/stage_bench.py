"""HBM-bound stages, batched launches: prompt extraction and upscale+threshold (CUDA events, L2 flushed)."""
import json
import statistics
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from samcarriestheburden_b200 import synthetic as O  # noqa: E402  (synthetic inputs)
from samcarriestheburden_b200.segment_anything.modeling.sam import upscale_masks  # noqa: E402
from samcarriestheburden_b200.segment_anything.utils.prompt_utils import extract_seeds_boxes  # noqa: E402
from samcarriestheburden_b200.segment_anything.utils.transforms import ResizeLongestSide  # noqa: E402


def timed(fn, reps=8, warm=3):
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    evs = []
    for i in range(reps + warm):  # queued back to back, read after one synchronise (no idle gaps between samples)
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        evs.append((e0, e1))
    torch.cuda.synchronize()
    ms = [a.elapsed_time(b) for a, b in evs[warm:]]
    return statistics.mean(ms)


def run(peak_gbs=6545.9):
    out = {}
    n_img = 256
    masks = torch.from_numpy(np.stack([O.synthetic_unet_masks(i % 16) for i in range(n_img)])).cuda()
    t = timed(lambda: extract_seeds_boxes(masks))
    by = masks.numel()
    out["prompt_extraction"] = {"images": n_img, "ms": t, "bytes": by, "gbs": by / t / 1e6, "frac_hbm": by / t / 1e6 / peak_gbs}
    for (oh, ow) in [(1024, 1024), (1182, 754)]:
        n = 256
        low = torch.randn((n, 1, 256, 256), device="cuda")
        inp = ResizeLongestSide.get_preprocess_shape(oh, ow, 1024)
        t = timed(lambda: upscale_masks(low, inp, (oh, ow), small_size=(384, 224)))
        by = n * (256 * 256 * 4 + oh * ow + 384 * 224)
        out[f"upscale_threshold_{oh}x{ow}"] = {"masks": n, "ms": t, "bytes": by, "gbs": by / t / 1e6,
                                               "frac_hbm": by / t / 1e6 / peak_gbs}
    # image ingest: native-resolution uint8 radiographs (largest CVAT size) -> encoder input size, Pillow-exact resize
    tr = ResizeLongestSide(1024)
    H, W = 2570, 2040
    n = 16
    imgs = [torch.randint(0, 256, (H, W, 3), dtype=torch.uint8, device="cuda") for _ in range(n)]
    oh, ow = tr.get_preprocess_shape(H, W, 1024)
    t = timed(lambda: [tr.apply_image_cuda(im) for im in imgs])
    by = n * (H * W * 3 + oh * ow * 3)
    out[f"ingest_resize_{H}x{W}"] = {"images": n, "ms": t, "bytes": by, "gbs": by / t / 1e6,
                                      "frac_hbm": by / t / 1e6 / peak_gbs}
    return out


if __name__ == "__main__":
    print(json.dumps(run(), indent=1))
