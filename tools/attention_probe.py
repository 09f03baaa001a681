"""Times the two encoder attention kernels alone (CUDA events, ViT-H shapes) and optionally compares the output with one
saved by an earlier run (A/B of kernel variants selected by environment variables, one process per variant).

usage: attention_probe.py [B] [fp16|bf16] [save|check] [file]"""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402

from samcarriestheburden_b200 import _lib  # noqa: E402

lib = _lib.load()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
fmt = sys.argv[2] if len(sys.argv) > 2 else "fp16"
mode = sys.argv[3] if len(sys.argv) > 3 else ""
path = sys.argv[4] if len(sys.argv) > 4 else "gpurun_out/attention_probe.pt"
heads, hd = 16, 80
dt, of = (torch.float16, 1) if fmt == "fp16" else (torch.bfloat16, 0)
D = heads * hd
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(1)
qkv = torch.randn((B * 4096, 3 * D), device=dev, generator=g).to(dt)
bias = torch.randn((3 * D,), device=dev, generator=g).to(dt)
outs = {}
for glob, S, name in ((0, 14, "window"), (1, 64, "global")):
    rel_h = (0.5 * torch.randn((2 * S - 1, hd), device=dev, generator=g)).to(dt)
    rel_w = (0.5 * torch.randn((2 * S - 1, hd), device=dev, generator=g)).to(dt)
    out = torch.zeros((B * 4096, D), dtype=dt, device=dev)

    def run():
        _lib.check(lib.b200sam_encoder_attention(qkv.data_ptr(), bias.data_ptr(), rel_h.data_ptr(), rel_w.data_ptr(),
                                                 out.data_ptr(), B, heads, hd, glob, of, _lib.current_stream()))
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    n = 20
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
    ev[0].record()
    for i in range(n):
        run()
        ev[i + 1].record()
    torch.cuda.synchronize()
    ts = sorted(ev[i].elapsed_time(ev[i + 1]) * 1e3 for i in range(n))
    print(f"{name:7s} attention B={B} {fmt}: median {ts[n // 2]:8.1f} us   min {ts[0]:8.1f} us", flush=True)
    outs[name] = out.float().cpu()
if mode == "save":
    torch.save(outs, path)
elif mode == "check":
    ref = torch.load(path)
    for k in outs:
        d = (outs[k] - ref[k]).abs()
        print(f"{k}: max |diff| vs saved {float(d.max()):.3e}  (ref max {float(ref[k].abs().max()):.3f}, "
              f"nan {int(torch.isnan(outs[k]).sum())})")
