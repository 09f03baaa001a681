"""Per-CTA timeline of the windowed attention kernel from its clock64 trace (b200sam_debug_window_trace).

Prints, over all CTAs of one ViT-H launch (interior 14x14 windows and edge windows separately), the median SM-cycle offset
of every stamp from the CTA's start, the median CTA lifetime and the launch duration."""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402

from samcarriestheburden_b200 import _lib  # noqa: E402

lib = _lib.load()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
heads, hd, D = 16, 80, 1280
dev = "cuda"
dt = torch.float16
qkv = torch.randn((B * 4096, 3 * D), device=dev).to(dt)
bias = torch.randn((3 * D,), device=dev).to(dt)
rel_h = (0.5 * torch.randn((27, hd), device=dev)).to(dt)
rel_w = (0.5 * torch.randn((27, hd), device=dev)).to(dt)
out = torch.zeros((B * 4096, D), dtype=dt, device=dev)
n_cta = 25 * heads * B
trace = torch.zeros((n_cta, 32), dtype=torch.int64, device=dev)


def run():
    _lib.check(lib.b200sam_encoder_attention(qkv.data_ptr(), bias.data_ptr(), rel_h.data_ptr(), rel_w.data_ptr(),
                                             out.data_ptr(), B, heads, hd, 0, 1, _lib.current_stream()))


for _ in range(3):
    run()
torch.cuda.synchronize()
lib.b200sam_debug_window_trace(trace.data_ptr())
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
run()
e1.record()
torch.cuda.synchronize()
lib.b200sam_debug_window_trace(None)
print(f"launch {e0.elapsed_time(e1) * 1e3:.1f} us (traced)")
t = trace.cpu().view(B, heads, 25, 32)
names = {0: "entry (control)", 31: "entry (softmax)", 1: "setup done (barriers, TMEM)", 2: "Q + table landed, tails zeroed",
         8: "softmax: Q tails zeroed", 9: "softmax: T MMAs retired", 10: "softmax: bias gathered", 3: "control: pre_done",
         11: "softmax: K patched", 4: "control: K landed + patched -> QK0", 12: "tile0 scores ready", 13: "tile0 pass 1 done",
         14: "tile0 chunk0 P", 15: "tile0 chunk1 P", 16: "tile0 chunk2 P", 17: "tile0 chunk3 P", 5: "control: P0 ready",
         6: "control: V landed", 18: "tile0 O ready", 19: "tile0 stored", 20: "tile1 scores ready", 21: "tile1 pass 1 done",
         22: "tile1 chunk0 P", 23: "tile1 chunk1 P", 24: "tile1 chunk2 P", 25: "tile1 chunk3 P", 26: "tile1 O ready",
         27: "tile1 stored", 7: "TMEM freed (end)"}
win = torch.arange(25)
interior = ((win // 5) < 4) & ((win % 5) < 4)
for label, mask in (("interior 14x14 windows", interior), ("edge windows", ~interior)):
    tt = t[:, :, mask, :].reshape(-1, 32)
    rel = tt - tt[:, :1]
    print(f"--- {label}: {tt.shape[0]} CTAs, cycles from CTA entry (median / p90)")
    order = sorted(names, key=lambda i: float(rel[:, i].float().median()))
    for i in order:
        v = rel[:, i].float()
        v = v[tt[:, i] != 0]
        if v.numel():
            print(f"  {names[i]:36s} {v.median():9.0f} {v.quantile(0.9):9.0f}")
