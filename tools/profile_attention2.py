"""Profiling driver: one attention mode on ViT-H shapes (for ncu).  usage: profile_attention2.py <batch> <mode>"""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402

from samcarriestheburden_b200 import _lib  # noqa: E402

lib = _lib.load()
B, mode = int(sys.argv[1]), int(sys.argv[2])
heads, hd = 16, 80
D = heads * hd
dev = "cuda"
S = 64 if mode in (1, 2) else 14
qkv = torch.randn((B * 4096, 3 * D), device=dev).bfloat16()
bias = torch.randn((3 * D,), device=dev).bfloat16()
out = torch.empty((B * 4096, D), dtype=torch.bfloat16, device=dev)
rel_h = (0.02 * torch.randn((2 * S - 1, hd), device=dev)).bfloat16()
rel_w = (0.02 * torch.randn((2 * S - 1, hd), device=dev)).bfloat16()
for it in range(3):
    _lib.check(lib.b200sam_encoder_attention(qkv.data_ptr(), bias.data_ptr(), rel_h.data_ptr(), rel_w.data_ptr(),
                                             out.data_ptr(), B, heads, hd, mode, _lib.current_stream()))
torch.cuda.synchronize()
print("done")
