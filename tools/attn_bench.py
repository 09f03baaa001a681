"""Micro-benchmark of the encoder attention kernels (CUDA events)."""
import statistics
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402

from samcarriestheburden_b200 import _lib  # noqa: E402

lib = _lib.load()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
heads, hd = 16, 80
D = heads * hd
dev = "cuda"
qkv = torch.randn((B * 4096, 3 * D), device=dev).bfloat16()
bias = torch.randn((3 * D,), device=dev).bfloat16()
out = torch.empty((B * 4096, D), dtype=torch.bfloat16, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for name, glob, S, flop in (("window (tcgen05)", 0, 14, 4 * heads * 4096 * 210 * hd), ("window (tcgen05 v2)", 4, 14, 4 * heads * 4096 * 210 * hd), ("window (tcgen05 v3)", 5, 14, 4 * heads * 4096 * 210 * hd), ("window (mma.sync)", 3, 14, 4 * heads * 4096 * 210 * hd), ("global (tcgen05)", 1, 64, 4 * heads * 4096 * 4160 * hd),
                            ("global (mma.sync)", 2, 64, 4 * heads * 4096 * 4160 * hd)):
    rel_h = (0.02 * torch.randn((2 * S - 1, hd), device=dev)).bfloat16()
    rel_w = (0.02 * torch.randn((2 * S - 1, hd), device=dev)).bfloat16()
    ms = []
    for i in range(8):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(lib.b200sam_encoder_attention(qkv.data_ptr(), bias.data_ptr(), rel_h.data_ptr(), rel_w.data_ptr(),
                                                 out.data_ptr(), B, heads, hd, glob, _lib.current_stream()))
        e1.record()
        e1.synchronize()
        if i >= 3:
            ms.append(e0.elapsed_time(e1))
    t = statistics.mean(ms)
    print(f"{name:20s} B={B}: {t*1e3:9.1f} us  {B*flop/t/1e9:8.1f} TF/s (algorithmic)")
