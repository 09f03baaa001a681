"""Profiling driver: the four linear shapes of one ViT block through the tcgen05 GEMM (for ncu --set full)."""
import argparse
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402

from samcarriestheburden_b200 import _lib  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=8)
ap.add_argument("--dim", type=int, default=1280)
ap.add_argument("--iters", type=int, default=3)
a = ap.parse_args()
lib = _lib.load()
D, M = a.dim, a.batch * 4096
dev = "cuda"
for it in range(a.iters):
    for name, N, K, out_bf16, gelu in [("qkv", 3 * D, D, 1, 0), ("proj", D, D, 0, 0), ("lin1", 4 * D, D, 1, 1),
                                       ("lin2", D, 4 * D, 0, 0)]:
        A = torch.randn((M, K), device=dev).bfloat16()
        W = (torch.randn((N, K), device=dev) / K ** 0.5).bfloat16()
        b = torch.randn((N,), device=dev)
        out = torch.zeros((M, N), dtype=torch.bfloat16 if out_bf16 else torch.float32, device=dev)
        _lib.check(lib.b200sam_gemm_bf16(A.data_ptr(), W.data_ptr(), out.data_ptr(), b.data_ptr(),
                                         out.data_ptr() if not out_bf16 else None, M, N, K, K, K, N, N, 0, gelu,
                                         out_bf16, 0, _lib.current_stream()))
torch.cuda.synchronize()
print("done")
