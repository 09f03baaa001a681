"""The synthetic input generators are deterministic, have the reference's parameter names / shapes, and the
oracle re-exports the very same objects (the product arm of bench.py imports them from the package, not the oracle)."""
import numpy as np
import torch

from oracle import sam_oracle as O
from oracle import unet_oracle as U
from samcarriestheburden_b200 import synthetic as S


def test_oracle_reexports_the_package_generators():
    for name in ("VIT_CONFIGS", "random_state_dict", "synthetic_radiograph", "synthetic_unet_masks", "synthetic_unet_probs"):
        assert getattr(O, name) is getattr(S, name)
    assert U.random_unet_state_dict is S.random_unet_state_dict and U.synthetic_radiograph_small is S.synthetic_radiograph_small


def test_generators_are_seeded_and_shaped():
    a, b = S.synthetic_radiograph(7), S.synthetic_radiograph(7)
    assert a.dtype == np.uint8 and a.shape == (1024, 1024, 3) and np.array_equal(a, b)
    assert np.array_equal(a[..., 0], a[..., 1]) and not np.array_equal(a, S.synthetic_radiograph(8))  # gray -> RGB
    m = S.synthetic_unet_masks(3)
    assert m.dtype == bool and m.shape == (17, 384, 224) and 1 <= int((m.reshape(17, -1).sum(1) == 0).sum()) <= 2
    p = S.synthetic_unet_probs(3)
    assert p.dtype == np.float32 and p.shape == (17, 384, 224) and 0.0 <= float(p.min()) and float(p.max()) <= 1.0
    sd1, sd2 = S.random_state_dict("vit_b", seed=0), S.random_state_dict("vit_b", seed=0)
    assert sd1.keys() == sd2.keys() and all(torch.equal(sd1[k], sd2[k]) for k in sd1)
    assert float(sd1["image_encoder.pos_embed"].abs().max()) > 0  # zero at init in the reference: randomised here
    assert sd1["image_encoder.blocks.0.attn.rel_pos_h"].shape == (27, 64)


def test_state_dict_loads_strictly_into_the_mirror():
    from samcarriestheburden_b200.segment_anything import sam_model_registry
    sam = sam_model_registry["vit_b"]()
    sam.load_state_dict(S.random_state_dict("vit_b", seed=1), strict=True)
    from samcarriestheburden_b200.custom_arcitecture.classic_u_net import UNet
    UNet(1, 17).load_state_dict(S.random_unet_state_dict(0), strict=True)
