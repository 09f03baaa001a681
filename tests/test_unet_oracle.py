"""CPU: the U-Net oracle against golden logits of the reference's own UNet class (tests/golden/make_golden_unet.py)."""
from pathlib import Path

import numpy as np
import torch

from oracle import unet_oracle as U

GOLD = Path(__file__).parent / "golden" / "unet_golden.npz"


def test_unet_oracle_matches_reference_class():
    g = np.load(GOLD)
    sd = U.random_unet_state_dict(0)
    for i in range(2):
        y = U.unet_forward(sd, torch.from_numpy(g[f"x{i}"])).numpy()
        ref = g[f"y{i}"]
        assert y.shape == ref.shape
        err = np.abs(y - ref).max()
        assert err < 1e-4 * max(1.0, float(np.abs(ref).max())), err


def test_unet_state_dict_keys_match_mirror():
    from samcarriestheburden_b200.custom_arcitecture.classic_u_net import UNet
    m = UNet(1, 17)
    sd = U.random_unet_state_dict(0)
    assert set(m.state_dict()) == set(sd)
    m.load_state_dict(sd, strict=True)
    assert all(tuple(m.state_dict()[k].shape) == tuple(v.shape) for k, v in sd.items())
    try:
        m(torch.zeros((1, 1, 32, 32)))
        raise AssertionError("the CPU path must not exist")
    except RuntimeError as e:  # B200SamError
        assert "no CPU path" in str(e)
