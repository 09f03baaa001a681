"""Golden vectors for the U-Net, generated in the build container:

    python tests/golden/make_golden_unet.py     (needs /root/reference; writes tests/golden/unet_golden.npz)

The reference's own `UNet` class (custom_arcitecture/classic_u_net.py) is imported unmodified, loaded with the oracle's
seeded state dict (strict) and run on two small inputs; the logits are stored as float32."""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import unet_oracle as U  # noqa: E402


def main():
    sys.path.insert(0, "/root/reference")
    from custom_arcitecture.classic_u_net import UNet
    torch.set_num_threads(8)
    model = UNet(1, 17, n_last_channel=64).eval()
    sd = U.random_unet_state_dict(0)
    model.load_state_dict(sd, strict=True)
    out = {}
    for seed, (H, W) in enumerate([(64, 48), (32, 80)]):
        x = U.synthetic_radiograph_small(seed, H, W)
        with torch.inference_mode():
            y = model(x)
        out[f"x{seed}"] = x.numpy()
        out[f"y{seed}"] = y.numpy().astype(np.float32)
    np.savez_compressed(ROOT / "tests" / "golden" / "unet_golden.npz", **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
