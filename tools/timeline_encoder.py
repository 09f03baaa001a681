"""In-run kernel timeline of the encoder forward (CUPTI through torch.profiler, no serialisation): how much of a
steady-state step is kernel time, how much is gaps between dependent launches.  Diagnostic only (never a bench value)."""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

from bench import build_model, synthetic_batch  # noqa: E402

dev = torch.device("cuda", 0)
batch = int(sys.argv[1]) if len(sys.argv) > 1 else 8
sam = build_model("vit_h", dev)
x = synthetic_batch(batch, 0).to(dev)
for _ in range(6):
    sam.encode_image(x)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        sam.encode_image(x)
    torch.cuda.synchronize()
evs = sorted((e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA),
             key=lambda e: e.time_range.start)
span = evs[-1].time_range.end - evs[0].time_range.start
busy = sum(e.time_range.end - e.time_range.start for e in evs)
gaps = [b.time_range.start - a.time_range.end for a, b in zip(evs, evs[1:])]
by = {}
for e in evs:
    k = e.name.split("<")[0].split("::")[-1][:40]
    d = by.setdefault(k, [0, 0.0])
    d[0] += 1
    d[1] += e.time_range.end - e.time_range.start
print(f"kernels {len(evs)}  span {span / 1e3:.2f} ms  busy {busy / 1e3:.2f} ms  idle {(span - busy) / 1e3:.2f} ms "
      f"({100 * (span - busy) / span:.1f} %)  median gap {sorted(gaps)[len(gaps) // 2]:.2f} us  max gap {max(gaps):.1f} us")
for k, (n, t) in sorted(by.items(), key=lambda kv: -kv[1][1]):
    print(f"  {k:40s} {n:5d}  {t / 1e3:8.3f} ms  {100 * t / busy:5.1f} %  avg {t / n:7.1f} us")
