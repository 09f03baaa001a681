"""Profiling driver for the decode stage: one batched refinement (8 images) + the batched HBM-bound launches.
Run under `ncu --profile-from-start off` (the region of interest is bracketed by cudaProfilerStart/Stop)."""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from bench import build_model  # noqa: E402
from samcarriestheburden_b200 import synthetic as O  # noqa: E402  (synthetic inputs)
from samcarriestheburden_b200.segment_anything.modeling.sam import upscale_masks  # noqa: E402
from samcarriestheburden_b200.segment_anything.sam_mask_decoder_head import EmbeddingStore, SAMMaskDecoderHead  # noqa: E402
from samcarriestheburden_b200.segment_anything.utils.prompt_utils import extract_seeds_boxes  # noqa: E402
from samcarriestheburden_b200.utils.seg_refinement import SAMSegRefiner  # noqa: E402

n_images = int(sys.argv[1]) if len(sys.argv) > 1 else 8
what = sys.argv[2] if len(sys.argv) > 2 else "all"
dev = torch.device("cuda", 0)
sam = build_model("vit_b", dev)  # the decoder is identical across model sizes
store = EmbeddingStore()
g = torch.Generator().manual_seed(0)
segs = []
for i in range(n_images):
    store.add(f"img{i}", torch.randn((1, 256, 64, 64), generator=g).to(dev), (1024, 1024), (1024, 1024))
    segs.append(torch.from_numpy(O.synthetic_unet_masks(i)).to(dev))
names = [f"img{i}" for i in range(n_images)]
head = SAMMaskDecoderHead(None, "vit_b", str(dev), store, sam_model=sam)
refiner = SAMSegRefiner("SAM", str(dev), [["box"], ["pos_points", "neg_points"]], sam_predictor=head)
batch = torch.stack(segs)
low = torch.randn((256, 1, 256, 256), device=dev)
masks = torch.from_numpy(np.stack([O.synthetic_unet_masks(i % 16) for i in range(256)])).to(dev)
from samcarriestheburden_b200.segment_anything.utils.transforms import ResizeLongestSide  # noqa: E402
native = torch.randint(0, 256, (2570, 2040, 3), dtype=torch.uint8, device=dev)
ingest = ResizeLongestSide(1024)


def roi():
    if what in ("all", "refine"):
        refiner.refine_batch(batch, names)
    if what in ("all", "stages"):
        upscale_masks(low, (1024, 1024), (1024, 1024), small_size=(384, 224))
        upscale_masks(low, (1024, 653), (1182, 754), small_size=(384, 224))
        extract_seeds_boxes(masks)
        ingest.apply_image_cuda(native)


roi()
torch.cuda.synchronize()
torch.cuda.profiler.start()
roi()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ok")
