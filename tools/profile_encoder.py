"""Profiling driver: N encoder forwards of a batch (ViT-H by default) and optionally one refinement, no timing.
Used under `ncu` (launch list / --set full captures); numbers printed under a profiler are never bench values."""
import argparse
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402

from bench import build_model, synthetic_batch  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--model", default="vit_h")
ap.add_argument("--batch", type=int, default=8)
ap.add_argument("--iters", type=int, default=2)
ap.add_argument("--refine", action="store_true")
a = ap.parse_args()
dev = torch.device("cuda", 0)
sam = build_model(a.model, dev)
x = synthetic_batch(a.batch, 0).to(dev)
for _ in range(a.iters):
    emb = sam.encode_image(x)
torch.cuda.synchronize()
if a.refine:
    from bench import refine_throughput
    print(refine_throughput(sam, dev, n_images=2))
print("done", float(emb.float().abs().mean()))
