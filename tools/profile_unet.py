"""Profiling driver: U-Net forward at batch B (no timing); run under ncu for the launch list."""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402

from samcarriestheburden_b200.custom_arcitecture.classic_u_net import UNet  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
torch.manual_seed(0)
m = UNet(1, 17).to("cuda")
x = torch.randn((B, 1, 384, 224), device="cuda")
for _ in range(2):
    m.predict_proba(x)
torch.cuda.synchronize()
torch.cuda.profiler.start()
m.predict_proba(x)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ok")
