"""Wall-clock phases of the end-to-end pipeline (diagnostic): embeddings / refinement / host copy, three repeats."""
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402

from bench import build_model  # noqa: E402
from samcarriestheburden_b200 import synthetic as O  # noqa: E402  (synthetic inputs)
from samcarriestheburden_b200.scripts.pipelines import generate_img_embeddings, refine_segmentations  # noqa: E402

dev = torch.device("cuda", 0)
sam = build_model("vit_h", dev)
n_images, batch = 32, 8
imgs = [O.synthetic_radiograph(100 + i) for i in range(n_images)]
names = [f"p{i}" for i in range(n_images)]
probs = [torch.from_numpy(O.synthetic_unet_probs(i % 8)).pin_memory() for i in range(n_images)]


def t():
    torch.cuda.synchronize()
    return time.perf_counter()


for rep in range(4):
    t0 = t()
    store, _ = generate_img_embeddings(sam, imgs, names, batch=batch)
    t1 = t()
    results, _ = refine_segmentations(sam, store, probs, names, batch=batch, ccl_selection="highest_probability")
    t2 = t()
    host = [r[1].cpu() for r in results]
    t3 = t()
    print(f"rep {rep}: embed {1e3 * (t1 - t0):.1f} ms  refine {1e3 * (t2 - t1):.1f} ms  d2h {1e3 * (t3 - t2):.1f} ms  "
          f"total {1e3 * (t3 - t0) / n_images:.2f} ms/img", flush=True)
import cProfile, pstats
pr = cProfile.Profile()
pr.enable()
store, _ = generate_img_embeddings(sam, imgs, names, batch=batch)
results, _ = refine_segmentations(sam, store, probs, names, batch=batch, ccl_selection="highest_probability")
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(25)
