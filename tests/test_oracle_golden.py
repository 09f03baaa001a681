"""Pin oracle/sam_oracle.py against fixtures produced by the reference's own code (tests/golden/make_golden.py)."""
import numpy as np
import pytest
import torch

from oracle import sam_oracle as O


def _sd(golden):
    return {k[3:]: torch.from_numpy(golden[k]) for k in golden.files if k.startswith("sd/")}


TINY = dict(depth=2, num_heads=2, global_attn_indexes=(1,), window_size=14)


def test_encoder_matches_reference(golden):
    sd = _sd(golden)
    img = torch.from_numpy(golden["enc/image"].astype(np.float32))
    x = O.preprocess(img, img_size=256)[None]
    feats = O.image_encoder(sd, x, **TINY)
    ref = torch.from_numpy(golden["enc/features"])
    assert feats.shape == ref.shape
    assert torch.allclose(feats, ref, atol=2e-5, rtol=1e-5), float((feats - ref).abs().max())


def test_prompt_encoder_and_decoder_match_reference(golden):
    sd = _sd(golden)
    feats = torch.from_numpy(golden["enc/features"])
    pe = O.dense_pe(sd, size=16)
    assert torch.allclose(pe, torch.from_numpy(golden["dec/dense_pe"]), atol=1e-6)
    boxes = torch.from_numpy(golden["dec/boxes"])
    sp, de = O.prompt_encoder(sd, None, boxes, None, emb_size=16, img_size=256)
    assert torch.allclose(sp, torch.from_numpy(golden["dec/sparse1"]), atol=1e-6)
    low1, iou1 = O.mask_decoder(sd, feats, pe, sp, de, multimask_output=False, heads=2)
    assert torch.allclose(low1, torch.from_numpy(golden["dec/low1"]), atol=1e-5)
    assert torch.allclose(iou1, torch.from_numpy(golden["dec/iou1"]), atol=1e-5)
    pts, labs = torch.from_numpy(golden["dec/points"]), torch.from_numpy(golden["dec/labels"])
    sp2, de2 = O.prompt_encoder(sd, (pts, labs), None, low1, emb_size=16, img_size=256)
    assert torch.allclose(sp2, torch.from_numpy(golden["dec/sparse2"]), atol=1e-6)
    assert torch.allclose(de2, torch.from_numpy(golden["dec/dense2"]), atol=1e-5)
    low2, iou2 = O.mask_decoder(sd, feats, pe, sp2, de2, multimask_output=False, heads=2)
    assert torch.allclose(low2, torch.from_numpy(golden["dec/low2"]), atol=1e-5)
    assert torch.allclose(iou2, torch.from_numpy(golden["dec/iou2"]), atol=1e-5)
    low3, iou3 = O.mask_decoder(sd, feats, pe, sp2, de2, multimask_output=True, heads=2)
    assert low3.shape[1] == 3
    assert torch.allclose(low3, torch.from_numpy(golden["dec/low3"]), atol=1e-5)
    assert torch.allclose(iou3, torch.from_numpy(golden["dec/iou3"]), atol=1e-5)


@pytest.mark.parametrize("i", [0, 1, 2])
def test_postprocess_matches_reference(golden, i):
    logits = torch.from_numpy(golden["post/logits"])
    orig, inp = tuple(golden[f"post/{i}/orig"]), tuple(golden[f"post/{i}/inp"])
    assert O.get_preprocess_shape(orig[0], orig[1]) == tuple(int(v) for v in inp)
    m = O.postprocess_masks(logits, inp, orig)
    assert torch.equal(m[:, :, ::37, ::41], torch.from_numpy(golden[f"post/{i}/sample"]))
    assert np.array_equal(np.packbits((m > 0).numpy()), golden[f"post/{i}/mask"])


@pytest.mark.parametrize("i", range(6))
def test_prompt_extraction_bit_exact(golden, i):
    masks = np.unpackbits(golden[f"pe/{i}/masks"])[: 17 * 384 * 224].reshape(17, 384, 224).astype(bool)
    prompts = O.prompt_extract(masks)
    assert [p.class_idx for p in prompts] == golden[f"pe/{i}/classes"].tolist()
    assert np.array_equal(np.stack([p.pos_seeds for p in prompts]), golden[f"pe/{i}/pos"])
    assert np.array_equal(np.stack([p.neg_seeds for p in prompts]), golden[f"pe/{i}/neg"])
    assert np.array_equal(np.stack([p.box for p in prompts]), golden[f"pe/{i}/box"])
    _, _, boxes, has_box = O.extract_seeds_boxes(masks)
    assert np.array_equal(boxes * has_box[:, None], golden[f"pe/{i}/boxes_all"])
    p0 = prompts[0]
    got = O.scale_coords(torch.from_numpy(p0.neg_seeds), p0.img_size, (1024, 653)).numpy()
    assert np.array_equal(got, golden[f"pe/{i}/pos_scaled"])
    gotb = O.scale_coords(torch.from_numpy(p0.box).reshape(-1, 2), p0.img_size, (1024, 653)).reshape(-1, 4).numpy()
    assert np.array_equal(gotb, golden[f"pe/{i}/box_scaled"])


def test_prompt_extraction_edge_cases():
    # empty input, single-class (reference raises on torch.cat of an empty list), all-overlap class
    s, hs, b, hb = O.extract_seeds_boxes(np.zeros((0, 4, 4), bool))
    assert s.shape == (0, 2) and hb.shape == (0,)
    one = np.zeros((3, 8, 8), bool)
    one[1, 2:4, 2:6] = True
    with pytest.raises(ValueError):
        O.prompt_extract(one)
    assert O.prompt_extract(np.zeros((3, 8, 8), bool)) == []
    # round-half-even: mean exactly x.5
    m = np.zeros((2, 4, 4), bool)
    m[0, 0, 0] = m[0, 1, 1] = True  # mean (0.5, 0.5) -> (0, 0)
    m[1, 1, 1] = m[1, 2, 2] = True  # overlaps class 0 at (1,1) -> seed from (2,2) only
    seeds, has, _, _ = O.extract_seeds_boxes(m)
    assert seeds[0].tolist() == [0, 0] and seeds[1].tolist() == [2, 2] and has.all()
