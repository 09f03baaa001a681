"""GPU: connected-component selection / SegEnhance through the C ABI against the oracle and the reference goldens."""
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import sam_oracle as O
from test_ccl_oracle import GOLD, golden_case

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
@pytest.mark.parametrize("sel", ["highest_probability", "largest"])
def test_ccl_matches_reference_golden(seed, sel):
    from samcarriestheburden_b200.utils.segmentation_preprocessing import remove_all_but_one_connected_component
    prob, ref = golden_case(np.load(GOLD), seed, sel)
    got = remove_all_but_one_connected_component(torch.from_numpy(prob).to(DEV), sel, num_iter=max(prob.shape[-2:]))
    assert np.array_equal(got.cpu().numpy(), ref)


def test_ccl_batch_edge_cases_and_shapes():
    from samcarriestheburden_b200.utils.segmentation_preprocessing import remove_all_but_one_connected_component_batch
    rng = np.random.default_rng(3)
    for (N, C, H, W) in [(1, 1, 1, 1), (2, 3, 7, 5), (3, 4, 64, 33), (8, 17, 384, 224)]:
        prob = rng.random((N, C, H, W)).astype(np.float32)
        prob[prob < 0.35] = 0.0
        if N > 1:
            prob[1, 0] = 0.0           # empty plane
            for n in (0, N - 1):       # lone pixel at index 0 of an image's first class: label 0 == background in the
                prob[n, 0, 0, 0] = 0.99  # reference, for EVERY image (the reference labels one image per call)
                if W > 1:
                    prob[n, 0, 0, 1] = 0.0
                if H > 1:
                    prob[n, 0, 1, :2] = 0.0
        for sel in ("largest", "highest_probability"):
            got = remove_all_but_one_connected_component_batch(torch.from_numpy(prob).to(DEV), sel).cpu().numpy()
            ref = np.stack([O.remove_all_but_one_connected_component(prob[n], sel) for n in range(N)])
            assert np.array_equal(got, ref), (N, C, H, W, sel)
    # a long serpentine component (geodesic length >> max(H, W)): one component after convergence
    H, W = 33, 33
    snake = np.zeros((1, 1, H, W), np.float32)
    for r in range(0, H, 2):
        snake[0, 0, r, :] = 0.8
        if r + 1 < H:
            snake[0, 0, r + 1, (W - 1) if (r // 2) % 2 == 0 else 0] = 0.8
    got = remove_all_but_one_connected_component_batch(torch.from_numpy(snake).to(DEV), "largest").cpu().numpy()
    assert np.array_equal(got, snake)


def test_seg_enhance_pipeline_equals_oracle_preprocessing():
    """SegEnhance.enhance / enhance_batch == refine(oracle-CCL'd probabilities); morphology of square / disk footprints."""
    import torch.nn.functional as F
    from samcarriestheburden_b200.segment_anything import sam_model_registry
    from samcarriestheburden_b200.segment_anything.sam_mask_decoder_head import EmbeddingStore, SAMMaskDecoderHead
    from samcarriestheburden_b200.utils.seg_refinement import SAMSegRefiner, SegEnhance
    from samcarriestheburden_b200.utils import segmentation_preprocessing as sp
    sam = sam_model_registry["vit_b"]()
    sam.load_state_dict(O.random_state_dict("vit_b", seed=0), strict=True)
    sam = sam.to(DEV)
    store = EmbeddingStore()
    g = torch.Generator().manual_seed(11)
    for i in range(2):
        store.add(f"e{i}", torch.randn((1, 256, 64, 64), generator=g).to(DEV), (1024, 1024), (1024, 1024))
    head = SAMMaskDecoderHead(None, "vit_b", DEV, store, sam_model=sam)
    refiner = SAMSegRefiner("SAM", DEV, [["box"], ["pos_points", "neg_points"]], sam_predictor=head)
    enh = SegEnhance(refiner, "highest_probability", "dilation", "square", 8, DEV)
    probs = [O.synthetic_unet_probs(20 + i) for i in range(2)]
    seg0, dice0 = enh.enhance(torch.from_numpy(probs[0]).to(DEV), "e0")
    pre = O.remove_all_but_one_connected_component(probs[0], "highest_probability")
    ref_seg, ref_dice = refiner.refine(torch.from_numpy(pre).to(DEV), "e0")
    assert torch.equal(seg0, ref_seg)
    assert torch.equal(torch.nan_to_num(dice0, nan=-1.0), torch.nan_to_num(ref_dice, nan=-1.0))
    # dilation by an 8x8 square == max-pool with the footprint anchored at (4, 4)
    want = F.max_pool2d(F.pad(torch.from_numpy(pre)[None], (4, 3, 4, 3), value=-1e4), 8, 1)[0]
    assert torch.equal(enh.last_preprocessed_seg.cpu(), want)
    segb, diceb = enh.enhance_batch(torch.from_numpy(np.stack(probs)).to(DEV), ["e0", "e1"])
    assert torch.equal(segb[0], seg0)
    # erosion with a disk footprint against a direct evaluation
    k = sp.structuring_element("disk", 2)
    x = torch.from_numpy(probs[1][:3]).to(DEV)
    er = sp.morph_flat(x, k, dilate=False).cpu()
    xp = F.pad(x.cpu()[None], (2, 2, 2, 2), value=1e4)[0]
    ref = torch.full_like(er, 1e4)
    for dy in range(5):
        for dx in range(5):
            if k[dy, dx]:
                ref = torch.minimum(ref, xp[:, dy:dy + x.shape[1], dx:dx + x.shape[2]])
    assert torch.equal(er, ref)
