"""A/B timing of the CTA-pair GEMM (tcgen05 cta_group::2) against the single-CTA kernel on the ViT-H linear shapes at batch
8, with the epilogues the encoder uses (fp16 operands, LayerNorm folding).  Diagnostic (CUDA events, back-to-back launches)."""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402

from samcarriestheburden_b200 import _lib  # noqa: E402

lib = _lib.load()
dev = "cuda"
print("max co-resident CTA pairs:", lib.b200sam_gemm_pair_max_clusters(), "SMs:",
      torch.cuda.get_device_properties(0).multi_processor_count, flush=True)
D, M = 1280, 8 * 4096
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
st = _lib.current_stream()
g = torch.Generator(device="cpu").manual_seed(0)
x = torch.randn((M, D), generator=g).to(dev)
x16 = torch.empty((M, D), dtype=torch.float16, device=dev)
stat = torch.empty((M, D // 64, 2), device=dev)
h16 = torch.randn((M, 4 * D), generator=g).to(dev).half()


def w(n, k):
    return (torch.randn((n, k), generator=g) / k ** 0.5).to(dev).half()


Wq, Wp, W1, W2 = w(3 * D, D), w(D, D), w(4 * D, D), w(D, 4 * D)
bq, bp, b1, b2 = (torch.randn((n,), generator=g).to(dev) for n in (3 * D, D, 4 * D, D))
sq, s1 = Wq.double().sum(1).float().contiguous(), W1.double().sum(1).float().contiguous()
att = torch.randn((M, D), generator=g).to(dev).half()
qkv = torch.empty((M, 3 * D), dtype=torch.float16, device=dev)
hid = torch.empty((M, 4 * D), dtype=torch.float16, device=dev)


def proj():
    _lib.check(lib.b200sam_gemm_ln_residual(att.data_ptr(), Wp.data_ptr(), bp.data_ptr(), x.data_ptr(), x.data_ptr(),
                                            x16.data_ptr(), stat.data_ptr(), M, D, D, 1, st))


def qkv_():
    _lib.check(lib.b200sam_gemm_ln_folded(x16.data_ptr(), Wq.data_ptr(), bq.data_ptr(), sq.data_ptr(), stat.data_ptr(),
                                          D // 64, 1e-6, qkv.data_ptr(), M, 3 * D, D, 0, 1, st))


def lin1():
    _lib.check(lib.b200sam_gemm_ln_folded(x16.data_ptr(), W1.data_ptr(), b1.data_ptr(), s1.data_ptr(), stat.data_ptr(),
                                          D // 64, 1e-6, hid.data_ptr(), M, 4 * D, D, 1, 1, st))


def lin2():
    _lib.check(lib.b200sam_gemm_ln_residual(h16.data_ptr(), W2.data_ptr(), b2.data_ptr(), x.data_ptr(), x.data_ptr(),
                                            x16.data_ptr(), stat.data_ptr(), M, D, 4 * D, 1, st))


def plain(N, K, A, W, out):
    return lambda: _lib.check(lib.b200sam_gemm_f16(A.data_ptr(), W.data_ptr(), out.data_ptr(), None, None, M, N, K, K, K, N,
                                                   0, 0, 0, 1, 0, st))


x32 = torch.empty((M, D), device=dev)


def gen(N, K, A, W, out, res, out16):
    return lambda: _lib.check(lib.b200sam_gemm_f16(A.data_ptr(), W.data_ptr(), out.data_ptr(), None, _lib.ptr(res), M, N, K,
                                                   K, K, N, N, 0, 0, out16, 0, st))


cases = [("proj+res+stats", proj, 2.0 * M * D * D),
         ("proj fp32+res", gen(D, D, att, Wp, x, x, 0), 2.0 * M * D * D),
         ("proj fp32 plain", gen(D, D, att, Wp, x32, None, 0), 2.0 * M * D * D),
         ("proj fp16 plain", gen(D, D, att, Wp, x16, None, 1), 2.0 * M * D * D), ("qkv folded", qkv_, 2.0 * M * 3 * D * D),
         ("lin1 folded+gelu", lin1, 2.0 * M * 4 * D * D), ("lin2+res+stats", lin2, 2.0 * M * D * 4 * D),
         ("qkv plain", plain(3 * D, D, x16, Wq, qkv), 2.0 * M * 3 * D * D),
         ("lin2 plain16", plain(D, 4 * D, h16, W2, x16), 2.0 * M * D * 4 * D)]
proj()
torch.cuda.synchronize()
for name, fn, flop in cases:
    res = {}
    for mode in (0, 1):
        lib.b200sam_set_gemm_pair(mode)
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        res[mode] = e0.elapsed_time(e1) / reps
    print(f"{name:18s} single {res[0] * 1e3:7.1f} us ({flop / res[0] / 1e9:6.0f} TF/s)   pair {res[1] * 1e3:7.1f} us "
          f"({flop / res[1] / 1e9:6.0f} TF/s)   pair/single {res[1] / res[0]:.3f}", flush=True)
lib.b200sam_set_gemm_pair(-1)

# the library GEMM (cuBLASLt through torch.matmul, fp16 in / fp16 out, fp32 accumulate) on the same four shapes, same timing:
# the reference point for "what does a plain GEMM of this shape reach on this board" (no bias, LayerNorm, GELU, residual)
for name, A, W, flop in (("qkv   [32768 x 1280] x [1280 x 3840]", x16, Wq, 2.0 * M * 3 * D * D),
                         ("proj  [32768 x 1280] x [1280 x 1280]", att, Wp, 2.0 * M * D * D),
                         ("lin1  [32768 x 1280] x [1280 x 5120]", x16, W1, 2.0 * M * 4 * D * D),
                         ("lin2  [32768 x 5120] x [5120 x 1280]", h16, W2, 2.0 * M * D * 4 * D)):
    Wt = W.t()
    for _ in range(3):
        torch.matmul(A, Wt)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        torch.matmul(A, Wt)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"cuBLAS fp16 {name}: {ms * 1e3:7.1f} us ({flop / ms / 1e9:6.0f} TF/s)", flush=True)
