"""Profiling driver: window + global attention kernels on ViT-H shapes (for ncu)."""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402

from samcarriestheburden_b200 import _lib  # noqa: E402

lib = _lib.load()
B, heads, hd = int(sys.argv[1]) if len(sys.argv) > 1 else 2, 16, 80
D = heads * hd
dev = "cuda"
qkv = torch.randn((B * 4096, 3 * D), device=dev).bfloat16()
bias = torch.randn((3 * D,), device=dev).bfloat16()
out = torch.empty((B * 4096, D), dtype=torch.bfloat16, device=dev)
for it in range(3):
    for glob, S in ((0, 14), (1, 64), (2, 64)):
        rel_h = (0.02 * torch.randn((2 * S - 1, hd), device=dev)).bfloat16()
        rel_w = (0.02 * torch.randn((2 * S - 1, hd), device=dev)).bfloat16()
        _lib.check(lib.b200sam_encoder_attention(qkv.data_ptr(), bias.data_ptr(), rel_h.data_ptr(), rel_w.data_ptr(),
                                                 out.data_ptr(), B, heads, hd, glob, _lib.current_stream()))
torch.cuda.synchronize()
print("done")
