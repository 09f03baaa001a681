"""Profiling driver: window + global attention kernels (tcgen05) on ViT-H shapes, fp16 operands (for ncu)."""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402

from samcarriestheburden_b200 import _lib  # noqa: E402

lib = _lib.load()
B, heads, hd = int(sys.argv[1]) if len(sys.argv) > 1 else 8, 16, 80
fmt = (sys.argv[2] if len(sys.argv) > 2 else "fp16")
dt, of = (torch.float16, 1) if fmt == "fp16" else (torch.bfloat16, 0)
D = heads * hd
dev = "cuda"
qkv = torch.randn((B * 4096, 3 * D), device=dev).to(dt)
bias = torch.randn((3 * D,), device=dev).to(dt)
out = torch.empty((B * 4096, D), dtype=dt, device=dev)
for it in range(3):
    for glob, S in ((0, 14), (1, 64)):
        rel_h = (0.02 * torch.randn((2 * S - 1, hd), device=dev)).to(dt)
        rel_w = (0.02 * torch.randn((2 * S - 1, hd), device=dev)).to(dt)
        _lib.check(lib.b200sam_encoder_attention(qkv.data_ptr(), bias.data_ptr(), rel_h.data_ptr(), rel_w.data_ptr(),
                                                 out.data_ptr(), B, heads, hd, glob, of, _lib.current_stream()))
torch.cuda.synchronize()
print("done")
