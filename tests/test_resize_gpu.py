"""GPU image ingest: b200sam_resize_u8 through the Python mirror against the oracle restatement of
Pillow's antialiased bilinear resample and the committed Pillow golden vectors.  Integer arithmetic: bit-exact."""
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import sam_oracle as O

pytestmark = pytest.mark.gpu
GOLD = np.load(Path(__file__).parent / "golden" / "resize_golden.npz")
N_CASES = sum(1 for k in GOLD.files if k.startswith("in"))
DEV = "cuda"


def _resize(img: np.ndarray, oh: int, ow: int, chw: bool) -> np.ndarray:
    from samcarriestheburden_b200.segment_anything.utils.transforms import resize_u8_cuda
    out = resize_u8_cuda(torch.from_numpy(img).cuda(), oh, ow, chw=chw)
    torch.cuda.synchronize()
    return out.cpu().numpy()


@pytest.mark.parametrize("i", range(N_CASES))
def test_resize_matches_pillow_golden(i):
    img, want = GOLD[f"in{i}"], GOLD[f"out{i}"]
    got = _resize(img, want.shape[0], want.shape[1], chw=False)
    assert np.array_equal(got, want)
    got_chw = _resize(img, want.shape[0], want.shape[1], chw=True)
    assert np.array_equal(got_chw, np.transpose(want, (2, 0, 1)))


@pytest.mark.parametrize("shape", [(1182, 754), (881, 578), (2570, 2040), (640, 1024), (300, 300), (1024, 1024)])
def test_apply_image_cuda_native_sizes_bit_exact(shape):
    """ResizeLongestSide.apply_image_cuda == the oracle (== Pillow) on CVAT-like native radiograph sizes."""
    from samcarriestheburden_b200.segment_anything.utils.transforms import ResizeLongestSide
    H, W = shape
    img = O.synthetic_radiograph(3, H, W)
    img[::7, ::5] = np.random.default_rng(0).integers(0, 256, size=img[::7, ::5].shape, dtype=np.uint8)  # sharp detail
    tr = ResizeLongestSide(1024)
    want = O.apply_image(img, 1024)
    got = tr.apply_image_cuda(img, chw=False).cpu().numpy()
    assert got.shape == want.shape
    assert np.array_equal(got, want)
    assert np.array_equal(tr.apply_image(img), want)  # the host path (Pillow itself) agrees as well


def test_resize_extreme_ratios():
    """Many taps per output sample (50x down-scaling: 101 taps) and strong up-scaling, both axes different."""
    rng = np.random.default_rng(9)
    img = rng.integers(0, 256, size=(700, 900, 3), dtype=np.uint8)
    assert np.array_equal(_resize(img, 14, 21, chw=False), O.resize_bilinear_u8(img, 14, 21))
    small = rng.integers(0, 256, size=(7, 5, 3), dtype=np.uint8)
    assert np.array_equal(_resize(small, 333, 129, chw=False), O.resize_bilinear_u8(small, 333, 129))
    assert np.array_equal(_resize(small, 1, 1, chw=True), np.transpose(O.resize_bilinear_u8(small, 1, 1), (2, 0, 1)))


def test_resize_identity_and_single_channel_and_errors():
    from samcarriestheburden_b200 import _lib
    from samcarriestheburden_b200.segment_anything.utils.transforms import resize_u8_cuda
    rng = np.random.default_rng(1)
    img = rng.integers(0, 256, size=(37, 53, 1), dtype=np.uint8)
    same = _resize(img, 37, 53, chw=True)  # no pass needed: layout conversion only
    assert np.array_equal(same, np.transpose(img, (2, 0, 1)))
    assert np.array_equal(_resize(img, 37, 53, chw=False), img)
    up = _resize(img, 80, 53, chw=False)  # height only
    assert np.array_equal(up, O.resize_bilinear_u8(img, 80, 53))
    with pytest.raises(AssertionError):
        resize_u8_cuda(torch.zeros((4, 4), dtype=torch.uint8, device="cuda"), 2, 2)
    lib = _lib.load()
    x = torch.zeros((4, 4, 3), dtype=torch.uint8, device="cuda")
    y = torch.zeros((2, 4, 3), dtype=torch.uint8, device="cuda")
    # height changes but no vertical table: loud error, no silent fallback
    rc = lib.b200sam_resize_u8(x.data_ptr(), 4, 4, 3, None, None, 0, None, None, 0, 2, 4, None, y.data_ptr(), 0, None)
    assert rc != 0 and b"vertical" in lib.b200sam_last_error()


def test_set_image_uses_gpu_resize_and_matches_host_resize():
    """SamPredictor.set_image on a non-square native image == set_torch_image of the Pillow-resized image."""
    from samcarriestheburden_b200.segment_anything import SamPredictor, sam_model_registry
    sam = sam_model_registry["vit_b"]()
    sam.load_state_dict(O.random_state_dict("vit_b", seed=0), strict=True)
    sam = sam.to("cuda")
    pred = SamPredictor(sam)
    img = O.synthetic_radiograph(11, 591, 377)
    pred.set_image(img)
    a = pred.get_image_embedding().clone()
    assert pred.original_size == (591, 377) and pred.input_size == (1024, 653)
    host = torch.from_numpy(O.apply_image(img, 1024)).permute(2, 0, 1).contiguous()[None].cuda()
    pred.set_torch_image(host, (591, 377))
    assert torch.equal(a, pred.get_image_embedding())


def test_cv2_linear_resize_matches_cv2_golden_and_oracle():
    """U-Net ingest (scripts/save_refined_segmentations.py:63-67): the GPU restatement of cv2.resize(INTER_LINEAR) on uint8 is
    bit-exact against goldens written by cv2 itself, and the fused normalisation equals the reference's fp32 operations."""
    from test_cv2resize_oracle import golden_cases, make_image
    from samcarriestheburden_b200.segment_anything.utils.transforms import cv_resize_linear_cuda
    g, seeds = golden_cases()
    mean, std = 0.3505533917353781, 0.22763733675869177
    for seed in seeds:
        H, W = (int(v) for v in g[f"shape_{seed}"])
        img = make_image(seed, H, W)
        got = cv_resize_linear_cuda(torch.from_numpy(img).to(DEV), 384, 224)
        assert got.dtype == torch.uint8 and np.array_equal(got.cpu().numpy(), g[f"out_{seed}"]), (seed, H, W)
        norm = cv_resize_linear_cuda(torch.from_numpy(img).to(DEV), 384, 224, normalize=(mean, std)).cpu()
        ref = (torch.from_numpy(g[f"out_{seed}"]).float() / 255 - mean) / std      # the reference's :64 and :67
        assert torch.equal(norm, ref), (seed, float((norm - ref).abs().max()))
    # batched same-size images and other target sizes against the oracle
    rng = np.random.default_rng(3)
    batch = rng.integers(0, 256, (3, 301, 517), dtype=np.uint8)
    got = cv_resize_linear_cuda(torch.from_numpy(batch).to(DEV), 97, 131).cpu().numpy()
    for i in range(3):
        assert np.array_equal(got[i], O.cv2_resize_linear_u8(batch[i], 97, 131))


def test_medsam_ingest_matches_oracle_and_embeds():
    """MedSAM branch (scripts/generate_img_embeddings.py:49-64): cubic resize + min-max normalisation bit-exact against the
    oracle (pinned to cv2 goldens), and the embeddings of the `medsam` driver against the oracle encoder on that input."""
    from test_cv2resize_oracle import golden_cases, make_image
    from samcarriestheburden_b200.scripts.pipelines import generate_img_embeddings
    from samcarriestheburden_b200.segment_anything import sam_model_registry
    from samcarriestheburden_b200.segment_anything.utils.transforms import medsam_preprocess_cuda
    g, _ = golden_cases()
    for seed in (20, 22):
        H, W = (int(v) for v in g[f"cubic_shape_{seed}"])
        gray = make_image(seed, H, W)
        out, resized = medsam_preprocess_cuda(torch.from_numpy(gray).to(DEV), 1024, return_resized=True)
        assert np.array_equal(resized.cpu().numpy()[::8], g[f"cubic_rows_{seed}"])
        assert torch.equal(out.cpu(), O.medsam_preprocess(gray))
    flat = np.full((50, 70), 9, np.uint8)  # max == min: the reference divides by clip(0, 1e-8) -> all zeros
    assert float(medsam_preprocess_cuda(torch.from_numpy(flat).to(DEV)).abs().max()) == 0.0
    sd = O.random_state_dict("vit_b", seed=0)
    sam = sam_model_registry["vit_b"]()
    sam.load_state_dict(sd, strict=True)
    sam = sam.to(DEV)
    imgs = [np.repeat(make_image(40 + i, 300 + 50 * i, 200)[:, :, None], 3, axis=2) for i in range(3)]
    store, emb = generate_img_embeddings(sam, imgs, ["a", "b", "c"], batch=2, gather=True, sam_type="medsam")
    assert emb.shape == (3, 256, 64, 64)
    assert tuple(store["b"].attrs["original_size"]) == (350, 200) and tuple(store["b"].attrs["input_size"]) == (1024, 1024)
    ref = O.image_encoder(sd, O.medsam_preprocess(imgs[1][..., 0]), **O.VIT_CONFIGS["vit_b"])
    rel = float((emb[1:2].cpu() - ref).norm() / ref.norm())
    print(f"medsam embedding rel-L2 {rel:.2e}")
    assert rel < 1.5e-3
