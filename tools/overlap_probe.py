"""Where does the staged embed + refine pipeline spend its time?  Stage size sweep, with and without the second stream."""
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from samcarriestheburden_b200 import synthetic as O  # noqa: E402
from samcarriestheburden_b200.scripts import pipelines as P  # noqa: E402
from samcarriestheburden_b200.segment_anything import sam_model_registry  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 192
dev = "cuda:0"
sam = sam_model_registry["vit_h"]()
sam.load_state_dict(O.random_state_dict("vit_h", seed=0), strict=True)
sam = sam.to(dev)
bases = [O.synthetic_radiograph(200 + k) for k in range(4)]
rng = np.random.default_rng(5)
imgs = [np.roll(bases[i % 4], (int(rng.integers(0, 1024)), int(rng.integers(0, 1024))), axis=(0, 1)) for i in range(n)]
pbase = [torch.from_numpy(O.synthetic_unet_probs(k)).pin_memory() for k in range(8)]
probs = [pbase[i % 8] for i in range(n)]
names = [f"s{i}" for i in range(n)]


def two_phase():
    store, _ = P.generate_img_embeddings(sam, imgs, names, batch=8)
    P.refine_segmentations(sam, store, probs, names, batch=8, ccl_selection="highest_probability")


def timed(fn, label):
    fn()
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    fn()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"{label:50s} {n / dt:7.1f} images/s  ({1e3 * dt / n:.2f} ms/image)", flush=True)


timed(two_phase, "two phases (encode all, then refine all)")
for stage in (n, 64, 32):
    for overlap in (False, True):
        timed(lambda: P.embed_and_refine(sam, imgs, probs, names, batch=8, stage=stage, overlap=overlap,
                                         ccl_selection="highest_probability"),
              f"staged, stage={stage}, second stream={overlap}")
