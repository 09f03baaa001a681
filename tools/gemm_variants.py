"""Micro-benchmark: epilogue variants of the tcgen05 GEMM on one shape (CUDA events, L2 flushed between launches)."""
import argparse
import statistics
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402

from samcarriestheburden_b200 import _lib  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--M", type=int, default=32768)
ap.add_argument("--N", type=int, default=1280)
ap.add_argument("--K", type=int, default=1280)
a = ap.parse_args()
lib = _lib.load()
dev = "cuda"
M, N, K = a.M, a.N, a.K
A = torch.randn((M, K), device=dev).bfloat16()
W = (torch.randn((N, K), device=dev) / K ** 0.5).bfloat16()
b = torch.randn((N,), device=dev)
out32 = torch.zeros((M, N), device=dev)
res32 = torch.randn((M, N), device=dev)
out16 = torch.zeros((M, N), dtype=torch.bfloat16, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def run(name, out, bias, res, gelu, out_bf16, reps=10, max_ctas=0):
    ms = []
    for i in range(reps + 3):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(lib.b200sam_gemm_bf16(A.data_ptr(), W.data_ptr(), out.data_ptr(), _lib.ptr(bias), _lib.ptr(res), M, N,
                                         K, K, K, N, N, 0, gelu, out_bf16, max_ctas, _lib.current_stream()))
        e1.record()
        e1.synchronize()
        if i >= 3:
            ms.append(e0.elapsed_time(e1))
    t = statistics.mean(ms)
    print(f"{name:32s} {t*1e3:8.1f} us  {2.0*M*N*K/t/1e9:8.1f} TF/s")


run("bf16 out, bias", out16, b, None, 0, 1)
run("bf16 out, bias [1-CTA]", out16, b, None, 0, 1, max_ctas=-1)
run("bf16 out, bias, gelu", out16, b, None, 1, 1)
run("bf16 out, no bias", out16, None, None, 0, 1)
run("f32 out, bias", out32, b, None, 0, 0)
run("f32 out, bias, residual(other)", out32, b, res32, 0, 0)
run("f32 out, bias, residual(inplace)", out32, b, out32, 0, 0)
run("f32 out, bias, residual(inplace) [1-CTA]", out32, b, out32, 0, 0, max_ctas=-1)
