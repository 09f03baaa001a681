"""GPU: U-Net through the C ABI against the CPU oracle (fp32) and the reference goldens."""
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import unet_oracle as U
from test_unet_oracle import GOLD

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(scope="module")
def unet():
    from samcarriestheburden_b200.custom_arcitecture.classic_u_net import UNet
    m = UNet(1, 17)
    m.load_state_dict(U.random_unet_state_dict(0), strict=True)
    return m.to(DEV)


def test_unet_matches_reference_golden(unet):
    g = np.load(GOLD)
    for i in range(2):
        y = unet(torch.from_numpy(g[f"x{i}"]).to(DEV)).cpu().numpy()
        ref = g[f"y{i}"]
        err = np.abs(y - ref).max()
        print(f"unet golden {i}: max |dlogit| {err:.2e} (max |logit| {np.abs(ref).max():.2f})")
        assert err < 5e-3, err


def test_unet_full_size_batch_masks(unet):
    """384 x 224 (the pipeline's U-Net grid), batch 3: logits vs the fp32 oracle, masks (p > 0.5) Dice >= 0.999."""
    x = torch.cat([U.synthetic_radiograph_small(10 + i) for i in range(3)])
    ref = U.unet_forward(U.random_unet_state_dict(0), x)
    logits = unet(x.to(DEV)).cpu()
    probs = unet.predict_proba(x.to(DEV)).cpu()
    err = float((logits - ref).abs().max())
    assert err < 5e-3, err
    assert torch.allclose(probs, torch.sigmoid(logits), atol=1e-6)
    a, b = probs > 0.5, torch.sigmoid(ref) > 0.5
    dice = 2.0 * float((a & b).sum()) / max(float(a.sum() + b.sum()), 1.0)
    mism = int((a != b).sum())
    print(f"unet 384x224 x3: max |dlogit| {err:.2e}, mask dice {dice:.6f}, {mism} mismatched px of {a.numel()}")
    assert dice >= 0.999
    # batch == single
    one = unet(x[1:2].to(DEV)).cpu()
    assert torch.equal(one, logits[1:2])


@pytest.mark.parametrize("shape", [(1, 32, 32), (1, 64, 48), (2, 96, 160), (5, 32, 272)])
def test_unet_other_geometries(unet, shape):
    """Sizes whose zero-bordered pixel grid (implicit convolution) does not line up with the GEMM's 128 / 256 row tiles,
    down to a 2 x 2 bottleneck (32 x 32; the reference rejects 16 x 16: InstanceNorm over a single element)."""
    B, H, W = shape
    g = torch.Generator().manual_seed(100 + H + W)
    x = torch.randn((B, 1, H, W), generator=g)
    ref = U.unet_forward(U.random_unet_state_dict(0), x)
    y = unet(x.to(DEV)).cpu()
    err = float((y - ref).abs().max())
    print(f"unet {B}x{H}x{W}: max |dlogit| {err:.2e} (max |logit| {float(ref.abs().max()):.2f})")
    assert err < 5e-3, err


def test_unet_feeds_the_refinement_pipeline(unet):
    """U-Net probabilities -> SegEnhance (CCL) -> SAM refinement, all on the device."""
    from samcarriestheburden_b200.segment_anything import sam_model_registry
    from samcarriestheburden_b200.segment_anything.sam_mask_decoder_head import EmbeddingStore, SAMMaskDecoderHead
    from samcarriestheburden_b200.utils.seg_refinement import SAMSegRefiner, SegEnhance
    from oracle import sam_oracle as O
    sam = sam_model_registry["vit_b"]()
    sam.load_state_dict(O.random_state_dict("vit_b", seed=0), strict=True)
    sam = sam.to(DEV)
    store = EmbeddingStore()
    g = torch.Generator().manual_seed(3)
    store.add("u0", torch.randn((1, 256, 64, 64), generator=g).to(DEV), (1024, 1024), (1024, 1024))
    head = SAMMaskDecoderHead(None, "vit_b", DEV, store, sam_model=sam)
    enh = SegEnhance(SAMSegRefiner("SAM", DEV, [["box"], ["pos_points", "neg_points"]], sam_predictor=head),
                     "highest_probability", "dilation", "square", 0, DEV)
    probs = unet.predict_proba(U.synthetic_radiograph_small(4).to(DEV))[0]
    if int(((probs > 0.5).flatten(1).sum(1) > 0).sum()) >= 2:
        seg, dice = enh.enhance(probs, "u0")
        assert seg.shape == (17, 384, 224) and seg.dtype == torch.bool and dice.shape == (17,)
