"""End-to-end parity of the CUDA path behind the reference API against the CPU oracle (same seeded inputs)."""
import numpy as np
import pytest
import torch

from oracle import sam_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda"


def dice(a, b):
    a, b = np.asarray(a, bool), np.asarray(b, bool)
    den = a.sum() + b.sum()
    return 1.0 if den == 0 else 2.0 * (a & b).sum() / den


@pytest.fixture(scope="module")
def vit_b():
    from samcarriestheburden_b200.segment_anything import sam_model_registry
    sd = O.random_state_dict("vit_b", seed=0)
    sam = sam_model_registry["vit_b"]()
    sam.load_state_dict(sd, strict=True)
    return sam.to(DEV), sd


@pytest.fixture(scope="module")
def embedding(vit_b):
    """(image, oracle fp32 embedding, CUDA embedding) for one non-square synthetic radiograph."""
    from samcarriestheburden_b200.segment_anything import SamPredictor
    sam, sd = vit_b
    img = O.synthetic_radiograph(3, 754, 589)
    pred = SamPredictor(sam)
    pred.set_image(img)
    resized = pred.transform.apply_image(img)
    x = O.preprocess(torch.from_numpy(resized).permute(2, 0, 1).float())[None]
    ref = O.image_encoder(sd, x, **O.VIT_CONFIGS["vit_b"])
    return img, ref, pred


# Embedding gates = ~1.5-2 x the measured error of each operand format against the fp32 oracle (ViT-B/L/H all measure
# 5.1e-3 .. 5.5e-3 with bf16 operands and 6e-4 .. 7e-4 with fp16 operands).
REL_GATE = {"fp16": 1.5e-3, "bf16": 9e-3}
COS_GATE = {"fp16": 0.999998, "bf16": 0.99995}


def _rel_cos(got, ref):
    rel = float((got - ref).norm() / ref.norm())
    cos = float(torch.nn.functional.cosine_similarity(got.flatten().double(), ref.flatten().double(), dim=0))
    return rel, cos


def test_encoder_embedding_tolerance(embedding):
    """16-bit operands / fp32 accumulate against the fp32 oracle, default format (fp16) and LayerNorm folding."""
    _, ref, pred = embedding
    enc = pred.model.image_encoder
    assert enc.operand_format == "fp16" and enc.ln_fused
    got = pred.get_image_embedding().float().cpu()
    assert got.shape == (1, 256, 64, 64)
    rel, cos = _rel_cos(got, ref)
    print(f"encoder vit_b fp16 rel_l2={rel:.3e} cos={cos:.7f}")
    assert rel <= REL_GATE["fp16"] and cos >= COS_GATE["fp16"], (rel, cos)


@pytest.mark.parametrize("fmt,fused", [("fp16", False), ("bf16", True), ("bf16", False)])
def test_encoder_operand_formats_and_unfused_layernorm(vit_b, embedding, fmt, fused):
    """The other three (operand format, LayerNorm folding) combinations of the encoder against the same oracle embedding."""
    from samcarriestheburden_b200.segment_anything import SamPredictor
    sam, _ = vit_b
    img, ref, _ = embedding
    enc = sam.image_encoder
    try:
        enc.set_precision(fmt, fused)
        pred = SamPredictor(sam)
        pred.set_image(img)
        got = pred.get_image_embedding().float().cpu()
    finally:
        enc.set_precision("fp16", True)
    rel, cos = _rel_cos(got, ref)
    print(f"encoder vit_b {fmt} ln_fused={fused} rel_l2={rel:.3e} cos={cos:.7f}")
    assert rel <= REL_GATE[fmt] and cos >= COS_GATE[fmt], (fmt, fused, rel, cos)


def _e2e_refine(sam, sd, model, seed, native):
    """image -> CUDA encoder -> CUDA two-pass refine, against the oracle's fp32 encoder -> oracle refine (the reference
    path: predictor.py:34-90 features into utils/seg_refinement.py:99-116)."""
    from samcarriestheburden_b200.segment_anything import SamPredictor
    from samcarriestheburden_b200.segment_anything.sam_mask_decoder_head import EmbeddingStore, SAMMaskDecoderHead
    from samcarriestheburden_b200.utils.seg_refinement import SAMSegRefiner
    img = O.synthetic_radiograph(seed, *native)
    pred = SamPredictor(sam)
    pred.set_image(img)
    resized = pred.transform.apply_image(img)
    ref_emb = O.image_encoder(sd, O.preprocess(torch.from_numpy(resized).permute(2, 0, 1).float())[None], **O.VIT_CONFIGS[model])
    rel, _ = _rel_cos(pred.features.float().cpu(), ref_emb)
    store = EmbeddingStore()
    store.add("img", pred.features, native, pred.input_size)
    head = SAMMaskDecoderHead(None, model, DEV, store, sam_model=sam)
    refiner = SAMSegRefiner("SAM", DEV, [["box"], ["pos_points", "neg_points"]], sam_predictor=head)
    seg = O.synthetic_unet_masks(seed)
    got_seg, got_dice = refiner.refine(torch.from_numpy(seg.copy()), "img")
    ref_seg, ref_dice, ref_native, _ = O.refine(sd, ref_emb, seg, pred.input_size, native)
    # native-resolution masks of every refined class (second pass) through the reference-shaped predict_mask API
    from samcarriestheburden_b200.segment_anything.utils.prompt_utils import PromptExtractor
    ds, mism, total = [], 0, 0
    for p in PromptExtractor(torch.from_numpy(seg).to(DEV)).extract():
        _, _, low1 = head.predict_mask("img", p, ["box"])
        m2, _, _ = head.predict_mask("img", p, ["pos_points", "neg_points"], low1)
        got = m2[0, 0].cpu().numpy()
        ds.append(dice(got, ref_native[p.class_idx]))
        mism += int((got != ref_native[p.class_idx]).sum())
        total += got.size
    got_small = got_seg.cpu().numpy()
    ds_small = [dice(got_small[c], ref_seg[c]) for c in range(seg.shape[0])]
    return dict(rel=rel, min_dice=min(ds), mean_dice=float(np.mean(ds)), mismatched=mism, total=total,
                min_dice_small=min(ds_small), est_dice_err=float(np.nanmax(np.abs(got_dice.numpy() - ref_dice))))


@pytest.mark.parametrize("native", [(754, 589), (1024, 1024)])
def test_e2e_mask_parity_vit_b(vit_b, native):
    """north_star's end-to-end bar: CUDA encoder -> CUDA refine masks vs the fp32 reference path at Dice >= 0.999 per
    class, mismatched-pixel count reported, embedding rel-L2 beside it (VERDICT r1 missing #1)."""
    sam, sd = vit_b
    r = _e2e_refine(sam, sd, "vit_b", 3, native)
    print(f"e2e vit_b fp16 {native}: emb rel-L2 {r['rel']:.2e}  native masks min Dice {r['min_dice']:.5f} mean {r['mean_dice']:.5f}  "
          f"mismatched {r['mismatched']} / {r['total']} px  384x224 min Dice {r['min_dice_small']:.5f}")
    assert r["min_dice"] >= 0.999 and r["min_dice_small"] >= 0.998, r
    assert r["est_dice_err"] < 2e-3


def test_e2e_mask_parity_bf16_operands_reported(vit_b):
    """The same end-to-end comparison with bf16 operands: NOT at the 0.999 bar at random init (logits hug the threshold,
    SURVEY section 7) -- which is why fp16 is the default.  Gated as a regression bound only; the number is printed."""
    sam, sd = vit_b
    try:
        sam.image_encoder.set_precision("bf16")
        r = _e2e_refine(sam, sd, "vit_b", 3, (754, 589))
    finally:
        sam.image_encoder.set_precision("fp16")
    print(f"e2e vit_b bf16: emb rel-L2 {r['rel']:.2e}  native masks min Dice {r['min_dice']:.5f}  "
          f"mismatched {r['mismatched']} / {r['total']} px")
    assert r["min_dice"] >= 0.994, r


def test_encoder_batch_matches_single(vit_b):
    sam, _ = vit_b
    imgs = torch.stack([torch.from_numpy(O.synthetic_radiograph(s)).permute(2, 0, 1) for s in (11, 12, 13)]).to(DEV)
    batch = sam.encode_image(imgs)
    single = torch.cat([sam.encode_image(imgs[i:i + 1]) for i in range(3)])
    torch.cuda.synchronize()
    assert torch.equal(batch, single)


def test_predict_box_matches_oracle(vit_b, embedding):
    """Decode stage from IDENTICAL fp32 embeddings: SamPredictor.predict(box)."""
    sam, sd = vit_b
    img, ref_emb, pred = embedding
    saved = pred.features
    pred.features = ref_emb.to(DEV)
    box = np.array([100.0, 150.0, 400.0, 600.0])
    masks, iou, low = pred.predict(box=box, multimask_output=False)
    m3, iou3, low3 = pred.predict(point_coords=np.array([[200.0, 300.0], [50.0, 60.0]]), point_labels=np.array([1, 0]),
                                  mask_input=low, multimask_output=True)
    pred.features = saved
    bt = torch.from_numpy(pred.transform.apply_boxes(box[None], (754, 589))).float()
    sp, de = O.prompt_encoder(sd, None, bt, None)
    low_o, iou_o = O.mask_decoder(sd, ref_emb, O.dense_pe(sd), sp, de, False)
    assert np.abs(low - low_o[0].numpy()).max() < 5e-4, np.abs(low - low_o[0].numpy()).max()
    assert np.abs(iou - iou_o[0].numpy()).max() < 5e-4
    m_o = (O.postprocess_masks(low_o, pred.input_size, (754, 589)) > 0)[0].numpy()
    d = dice(masks, m_o)
    print(f"predict(box): dice={d:.6f} mismatched={int((masks != m_o).sum())}")
    assert d >= 0.999
    pts = torch.from_numpy(pred.transform.apply_coords(np.array([[200.0, 300.0], [50.0, 60.0]]), (754, 589))).float()[None]
    sp, de = O.prompt_encoder(sd, (pts, torch.tensor([[1, 0]])), None, low_o)
    low3_o, iou3_o = O.mask_decoder(sd, ref_emb, O.dense_pe(sd), sp, de, True)
    assert low3.shape == (3, 256, 256)
    assert np.abs(low3 - low3_o[0].numpy()).max() < 1e-3
    assert np.abs(iou3 - iou3_o[0].numpy()).max() < 1e-3


def test_dense_pe_matches_oracle(vit_b):
    sam, sd = vit_b
    pe = sam.prompt_encoder.get_dense_pe().cpu()
    assert pe.shape == (1, 256, 64, 64)
    assert (pe - O.dense_pe(sd)).abs().max() < 2e-5


@pytest.mark.parametrize("orig", [(754, 589), (1024, 1024)])
def test_refine_matches_oracle(vit_b, orig):
    """Full two-pass refinement (box, then pos/neg points + previous logits), all prompts batched, against the
    oracle's per-class B=1 loop from the same fp32 embedding: Dice >= 0.999 per class, mismatches reported."""
    from samcarriestheburden_b200.segment_anything.sam_mask_decoder_head import EmbeddingStore, SAMMaskDecoderHead
    from samcarriestheburden_b200.utils.seg_refinement import SAMSegRefiner
    sam, sd = vit_b
    g = torch.Generator().manual_seed(orig[0])
    feats = torch.randn((1, 256, 64, 64), generator=g)
    inp = O.get_preprocess_shape(*orig)
    store = EmbeddingStore()
    store.add("img", feats.to(DEV), orig, inp)
    head = SAMMaskDecoderHead(None, "vit_b", DEV, store, sam_model=sam)
    refiner = SAMSegRefiner("SAM", DEV, [["box"], ["pos_points", "neg_points"]], sam_predictor=head)
    seg = O.synthetic_unet_masks(7)
    got_seg, got_dice = refiner.refine(torch.from_numpy(seg.copy()), "img")
    ref_seg, ref_dice, ref_native, _ = O.refine(sd, feats, seg, inp, orig)
    got_seg = got_seg.cpu().numpy()
    ds = [dice(got_seg[c], ref_seg[c]) for c in range(seg.shape[0])]
    mism = int((got_seg != ref_seg).sum())
    print(f"refine {orig}: min dice={min(ds):.6f} mismatched px={mism} of {got_seg.size}")
    assert min(ds) >= 0.999, ds
    assert np.allclose(got_dice.numpy(), ref_dice, atol=1e-3, equal_nan=True)
    # native-resolution masks of the second pass through the reference-shaped predict_mask API
    from samcarriestheburden_b200.segment_anything.utils.prompt_utils import PromptExtractor
    prompts = PromptExtractor(torch.from_numpy(seg).to(DEV)).extract()
    p = prompts[0]
    m1, s1, low1 = head.predict_mask("img", p, ["box"])
    m2, s2, _ = head.predict_mask("img", p, ["pos_points", "neg_points"], low1)
    assert m2.shape == (1, 1) + tuple(orig) and m2.dtype == torch.bool
    d = dice(m2[0, 0].cpu().numpy(), ref_native[p.class_idx])
    print(f"predict_mask native {orig}: dice={d:.6f}")
    assert d >= 0.999


def test_pipeline_drivers_match_single_image_api(vit_b):
    """generate_img_embeddings / refine_segmentations drivers (batched, mixed native sizes) == per-image API."""
    from samcarriestheburden_b200.scripts.pipelines import generate_img_embeddings, refine_segmentations
    from samcarriestheburden_b200.segment_anything import SamPredictor
    sam, sd = vit_b
    imgs = [O.synthetic_radiograph(21), O.synthetic_radiograph(22, 754, 589), O.synthetic_radiograph(23),
            O.synthetic_radiograph(24, 754, 589), O.synthetic_radiograph(25, 600, 1000)]
    names = [f"im{i}" for i in range(len(imgs))]
    store, gathered = generate_img_embeddings(sam, imgs, names, batch=2, gather=True)
    assert gathered.shape == (5, 256, 64, 64)
    pred = SamPredictor(sam)
    for i, img in enumerate(imgs):
        pred.set_image(img)
        assert torch.equal(pred.features[0], gathered[i]), i
        assert list(store[names[i]].attrs["original_size"]) == list(img.shape[:2])
        assert tuple(store[names[i]].attrs["input_size"]) == tuple(pred.input_size)
    segs = [torch.from_numpy(O.synthetic_unet_masks(30 + i)) for i in range(len(imgs))]
    results, allseg = refine_segmentations(sam, store, segs, names, gather=True)
    assert allseg.shape == (5, 17, 384, 224)
    # decode stage vs oracle from the SAME (CUDA-produced) embedding for one non-square image
    i = 1
    ref_seg, ref_dice, _, _ = O.refine(sd, gathered[i:i + 1].cpu(), segs[i].numpy(), store[names[i]].attrs["input_size"],
                                       store[names[i]].attrs["original_size"])
    got = allseg[i].bool().cpu().numpy()
    ds = [dice(got[c], ref_seg[c]) for c in range(17)]
    print(f"pipeline refine: min dice={min(ds):.6f} mismatched={int((got != ref_seg).sum())}")
    assert min(ds) >= 0.999


def test_overlapped_pipeline_equals_two_phase_drivers(vit_b):
    """embed_and_refine (encoder of stage s+1 and refinement of stage s on two CUDA streams) returns bit-identical
    embeddings and masks to generate_img_embeddings followed by refine_segmentations."""
    from samcarriestheburden_b200.scripts.pipelines import embed_and_refine, generate_img_embeddings, refine_segmentations
    sam, _ = vit_b
    imgs = [O.synthetic_radiograph(40 + i, *((754, 589) if i % 3 == 1 else (1024, 1024))) for i in range(7)]
    names = [f"ov{i}" for i in range(len(imgs))]
    probs = [torch.from_numpy(O.synthetic_unet_probs(i)) for i in range(len(imgs))]
    store, emb = generate_img_embeddings(sam, imgs, names, batch=2, gather=True)
    res, seg = refine_segmentations(sam, store, probs, names, gather=True, batch=2, ccl_selection="highest_probability")
    for _ in range(2):  # twice: the second pass reuses warm engines and exercises stream reuse
        store2, res2, emb2, seg2 = embed_and_refine(sam, imgs, probs, names, batch=2, stage=3, gather=True,
                                                    ccl_selection="highest_probability")
        torch.cuda.synchronize()
        assert torch.equal(emb, emb2)
        assert torch.equal(seg, seg2)
        assert [i for i, _, _ in res2] == [i for i, _, _ in res]
        for (_, a, da), (_, b, db) in zip(res, res2):
            assert torch.equal(a, b) and torch.equal(torch.nan_to_num(da, nan=-1.0), torch.nan_to_num(db, nan=-1.0))


def test_refine_batch_equals_per_image(vit_b):
    """Multi-image ragged decode (images with different numbers of classes / native sizes in ONE launch sequence)
    is bit-identical to the per-image calls: absent token slots never act as attention keys."""
    from samcarriestheburden_b200.segment_anything.sam_mask_decoder_head import EmbeddingStore, SAMMaskDecoderHead
    from samcarriestheburden_b200.utils.seg_refinement import SAMSegRefiner
    sam, _ = vit_b
    store = EmbeddingStore()
    g = torch.Generator().manual_seed(5)
    sizes = [(1024, 1024), (754, 589), (1024, 1024), (600, 1000)]
    segs = []
    for i, orig in enumerate(sizes):
        store.add(f"b{i}", torch.randn((1, 256, 64, 64), generator=g).to(DEV), orig, O.get_preprocess_shape(*orig))
        seg = torch.from_numpy(O.synthetic_unet_masks(50 + i))
        if i == 1:
            seg[[0, 4, 9, 11]] = False   # fewer classes -> fewer negative points than the other images
        if i == 2:
            seg[:] = False               # an image without any prompt
        segs.append(seg)
    head = SAMMaskDecoderHead(None, "vit_b", DEV, store, sam_model=sam)
    refiner = SAMSegRefiner("SAM", DEV, [["box"], ["pos_points", "neg_points"]], sam_predictor=head)
    names = [f"b{i}" for i in range(len(sizes))]
    got_seg, got_dice = refiner.refine_batch(torch.stack(segs), names)
    ks = set()
    for i in range(len(sizes)):
        one_seg, one_dice = refiner.refine(segs[i].clone(), names[i])
        ks.add(int((~torch.isnan(one_dice)).sum()))
        assert torch.equal(got_seg[i], one_seg), i
        assert torch.equal(torch.nan_to_num(got_dice[i], nan=-1.0), torch.nan_to_num(one_dice, nan=-1.0)), i
    assert len(ks) >= 3, ks  # the batch really was ragged
    # single-pass configuration (box + points in one pass, no self-refinement)
    r1 = SAMSegRefiner("SAM", DEV, ["pos_points", "neg_points", "box"], sam_predictor=head)
    b_seg, _ = r1.refine_batch(torch.stack(segs), names)
    for i in (0, 1):
        assert torch.equal(b_seg[i], r1.refine(segs[i].clone(), names[i])[0])


def test_vit_l_encoder_batch16_matches_oracle():
    """BASELINE.json configs[3]: ViT-L (hd = 64, depth 24, global blocks 5/11/17/23), batch 16 per GPU."""
    from samcarriestheburden_b200.segment_anything import sam_model_registry
    sd = O.random_state_dict("vit_l", seed=1)
    sam = sam_model_registry["vit_l"]()
    sam.load_state_dict(sd, strict=True)
    sam = sam.to(DEV)
    imgs = torch.stack([torch.from_numpy(O.synthetic_radiograph(40 + i)).permute(2, 0, 1) for i in range(16)]).to(DEV)
    emb = sam.encode_image(imgs)
    torch.cuda.synchronize()
    assert emb.shape == (16, 256, 64, 64)
    for i in (5, 15):
        ref = O.image_encoder(sd, O.preprocess(imgs[i].cpu().float())[None], **O.VIT_CONFIGS["vit_l"])
        rel, cos = _rel_cos(emb[i:i + 1].float().cpu(), ref)
        print(f"encoder vit_l image {i} rel_l2={rel:.3e} cos={cos:.7f}")
        assert rel <= REL_GATE["fp16"] and cos >= COS_GATE["fp16"], (i, rel, cos)
    for i in range(0, 16, 4):  # every batch position class: the same image alone gives the same embedding
        assert torch.equal(sam.encode_image(imgs[i:i + 1]), emb[i:i + 1]), i
    del sam
    torch.cuda.empty_cache()


def test_vit_h_encoder_batch8_matches_oracle():
    """BASELINE.json configs[1], the headline configuration: ViT-H (D 1280, hd 80, depth 32, global blocks 7/15/23/31) at
    the bench's batch of 8 (so the boustrophedon traversal and the multi-image attention grids are exercised), one image
    two images of the batch against the fp32 CPU oracle (a few seconds each on the box's host cores), every image of the
    batch against its own single-image run (bit-identical), then the end-to-end mask parity on the headline model."""
    from samcarriestheburden_b200.segment_anything import sam_model_registry
    sd = O.random_state_dict("vit_h", seed=2)
    sam = sam_model_registry["vit_h"]()
    sam.load_state_dict(sd, strict=True)
    sam = sam.to(DEV)
    imgs = torch.stack([torch.from_numpy(O.synthetic_radiograph(60 + i)).permute(2, 0, 1) for i in range(8)]).to(DEV)
    emb = sam.encode_image(imgs)
    torch.cuda.synchronize()
    assert emb.shape == (8, 256, 64, 64) and bool(torch.isfinite(emb).all())
    for i in (0, 6):
        ref = O.image_encoder(sd, O.preprocess(imgs[i].cpu().float())[None], **O.VIT_CONFIGS["vit_h"])
        rel, cos = _rel_cos(emb[i:i + 1].float().cpu(), ref)
        print(f"encoder vit_h image {i} rel_l2={rel:.3e} cos={cos:.7f}")
        assert rel <= REL_GATE["fp16"] and cos >= COS_GATE["fp16"], (i, rel, cos)
    # every image alone gives the same embedding (batch / traversal order independence)
    for i in range(8):
        assert torch.equal(sam.encode_image(imgs[i:i + 1]), emb[i:i + 1]), i
    r = _e2e_refine(sam, sd, "vit_h", 5, (1182, 754))
    print(f"e2e vit_h fp16 (1182, 754): emb rel-L2 {r['rel']:.2e}  native masks min Dice {r['min_dice']:.5f} mean "
          f"{r['mean_dice']:.5f}  mismatched {r['mismatched']} / {r['total']} px  384x224 min Dice {r['min_dice_small']:.5f}")
    assert r["min_dice"] >= 0.999, r
    del sam
    torch.cuda.empty_cache()


def test_engine_follows_weight_reloads_and_inplace_edits(vit_b):
    """`sam.load_state_dict(...)` on the PARENT module and in-place parameter edits after the first forward
    must re-pack the cached engines (packed tensors never alias live parameters)."""
    from samcarriestheburden_b200.segment_anything import sam_model_registry
    _, sd = vit_b
    img = torch.from_numpy(O.synthetic_radiograph(2)).permute(2, 0, 1)[None].to(DEV)
    sam = sam_model_registry["vit_b"]()
    sam.load_state_dict(sd, strict=True)
    sam = sam.to(DEV)
    e0 = sam.encode_image(img).clone()
    sd2 = O.random_state_dict("vit_b", seed=9)
    sam.load_state_dict(sd2, strict=True)          # parent-level reload after the first forward
    e1 = sam.encode_image(img).clone()
    fresh = sam_model_registry["vit_b"]()
    fresh.load_state_dict(sd2, strict=True)
    fresh = fresh.to(DEV)
    assert torch.equal(e1, fresh.encode_image(img)) and not torch.equal(e0, e1)
    with torch.no_grad():                           # in-place edit of one weight
        sam.image_encoder.blocks[3].mlp.lin2.bias.add_(0.5)
        fresh.image_encoder.blocks[3].mlp.lin2.bias.add_(0.5)
    e2 = sam.encode_image(img)
    assert not torch.equal(e1, e2)
    fresh2 = sam_model_registry["vit_b"]()
    fresh2.load_state_dict(fresh.state_dict(), strict=True)
    assert torch.equal(e2, fresh2.to(DEV).encode_image(img))
    # decoder engine: dense PE follows a reload of the prompt encoder's gaussian matrix
    pe_before = sam.prompt_encoder.get_dense_pe().clone()
    sam.load_state_dict(sd, strict=True)
    assert not torch.equal(pe_before, sam.prompt_encoder.get_dense_pe())
    del sam, fresh, fresh2
    torch.cuda.empty_cache()


def test_model_on_second_device_while_first_is_current():
    """the C ABI launches on the CURRENT device; every wrapper must enter the device that owns its tensors."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from samcarriestheburden_b200.segment_anything import sam_model_registry
    sd = O.random_state_dict("vit_b", seed=0)
    img = torch.from_numpy(O.synthetic_radiograph(2)).permute(2, 0, 1)[None]
    sam0 = sam_model_registry["vit_b"]()
    sam0.load_state_dict(sd, strict=True)
    e0 = sam0.to("cuda:0").encode_image(img.to("cuda:0"))
    sam1 = sam_model_registry["vit_b"]()
    sam1.load_state_dict(sd, strict=True)
    torch.cuda.set_device(0)
    e1 = sam1.to("cuda:1").encode_image(img.to("cuda:1"))
    torch.cuda.synchronize(1)
    assert torch.equal(e0.cpu(), e1.cpu())


def test_sam_forward_and_return_logits(vit_b, embedding):
    """Upstream batched API `Sam.forward` (sam.py:53-131) and `predict(return_logits=True)`."""
    sam, sd = vit_b
    img, ref_emb, pred = embedding
    resized = torch.from_numpy(pred.transform.apply_image(img)).permute(2, 0, 1).contiguous().to(DEV)
    box = torch.tensor([[100.0, 150.0, 400.0, 600.0]], device=DEV)
    boxes_in = pred.transform.apply_boxes_torch(box, (754, 589))
    out = sam([{"image": resized, "original_size": (754, 589), "boxes": boxes_in}], multimask_output=False)
    assert out[0]["masks"].shape == (1, 1, 754, 589) and out[0]["masks"].dtype == torch.bool
    m1, iou1, low1 = pred.predict(box=box[0].cpu().numpy(), multimask_output=False)
    assert np.array_equal(out[0]["masks"][0].cpu().numpy(), m1)
    logits, _, _ = pred.predict(box=box[0].cpu().numpy(), multimask_output=False, return_logits=True)
    assert logits.dtype == np.float32 and np.array_equal(logits > 0, m1)
    # error behaviour of the reference API
    from samcarriestheburden_b200.segment_anything import SamPredictor
    fresh = SamPredictor(sam)
    with pytest.raises(RuntimeError):
        fresh.predict(box=box[0].cpu().numpy())
    with pytest.raises(RuntimeError):
        fresh.get_image_embedding()
    with pytest.raises(AssertionError):
        fresh.set_image(img, image_format="XYZ")


def test_standalone_prompt_encoder_and_mask_decoder_forward(vit_b, embedding):
    """The reference's module-level API (prompt_encoder.py:128-168, mask_decoder.py:71-110), tensors in / tensors out,
    against the oracle: points + box, points only (pad point), mask input; then MaskDecoder.forward on those embeddings."""
    sam, sd = vit_b
    _, ref_emb, _ = embedding
    pe, md = sam.prompt_encoder, sam.mask_decoder
    g = torch.Generator().manual_seed(3)
    pts = torch.rand((3, 2, 2), generator=g) * 1000
    labs = torch.tensor([[1, 0], [1, 1], [0, 1]])
    boxes = torch.tensor([[10.0, 20.0, 300.0, 400.0], [50.0, 60.0, 900.0, 1000.0], [0.0, 0.0, 512.0, 512.0]])
    masks = torch.randn((3, 1, 256, 256), generator=g)
    for points, bx, mk in [((pts, labs), boxes, None), ((pts, labs), None, masks), (None, boxes, None), (None, None, masks)]:
        sp, de = pe(tuple(t.to(DEV) for t in points) if points else None, bx.to(DEV) if bx is not None else None,
                    mk.to(DEV) if mk is not None else None)
        sp_o, de_o = O.prompt_encoder(sd, points, bx, mk)
        assert sp.shape == sp_o.shape and de.shape[1:] == (256, 64, 64)
        assert (sp.cpu() - sp_o).abs().max() < 2e-5 if sp_o.numel() else True
        assert (de.cpu() - de_o.expand_as(de)).abs().max() < 2e-5
        low, iou = md(ref_emb.to(DEV), pe.get_dense_pe(), sp, de, multimask_output=(bx is None))
        low_o, iou_o = O.mask_decoder(sd, ref_emb, O.dense_pe(sd), sp_o, de_o, bx is None)
        assert low.shape == low_o.shape and iou.shape == iou_o.shape
        assert (low.cpu() - low_o).abs().max() < 1e-3 and (iou.cpu() - iou_o).abs().max() < 1e-3
    with pytest.raises(NotImplementedError):
        md(ref_emb.to(DEV), torch.zeros((1, 256, 64, 64), device=DEV), sp, de, multimask_output=False)
