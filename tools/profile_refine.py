"""Profiling driver: SAMSegRefiner.refine on a few images from resident embeddings (for ncu launch lists)."""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402

from bench import build_model, refine_throughput  # noqa: E402

dev = torch.device("cuda", 0)
sam = build_model("vit_b", dev)  # decoder is identical across model sizes
print(refine_throughput(sam, dev, n_images=int(sys.argv[1]) if len(sys.argv) > 1 else 3))
