"""CPU: the asynchronous result writer (storage.py) in its npy-directory layout (h5py is not part of this image) and the
reader that turns it back into an EmbeddingStore; the h5 layout is exercised only where h5py is importable."""
import json

import numpy as np
import pytest
import torch

from samcarriestheburden_b200.storage import AsyncResultWriter, open_embeddings


def test_embedding_writer_roundtrip(tmp_path):
    g = torch.Generator().manual_seed(0)
    feats = {f"img{i}": torch.randn((1, 256, 64, 64), generator=g) for i in range(5)}
    out = tmp_path / "emb"
    with AsyncResultWriter(out, "embedding", {"checkpoint": "sam_vit_h_4b8939.pth", "img_encoder_img_size": 1024}, depth=2) as w:
        for i, (k, v) in enumerate(feats.items()):
            w.put_embedding(k, v, (1182 + i, 754), (1024, 653))
    assert w.records == 5
    assert json.loads((out / "attrs.json").read_text())["img_encoder_img_size"] == 1024
    store = open_embeddings(out)
    assert store.attrs["checkpoint"] == "sam_vit_h_4b8939.pth"
    for i, (k, v) in enumerate(feats.items()):
        assert torch.equal(store[k]["features"], v)
        assert store[k].attrs["original_size"].tolist() == [1182 + i, 754]
        assert store[k].attrs["input_size"].tolist() == [1024, 653]
    with pytest.raises(FileExistsError):  # like h5py's 'x' mode in the reference: never overwrite
        AsyncResultWriter(out, "embedding", {})


def test_mask_writer_layout(tmp_path):
    seg = torch.rand((17, 384, 224)) > 0.5
    dice = torch.rand(17)
    dice[3] = float("nan")
    with AsyncResultWriter(tmp_path / "masks", "mask", {"labels": "{}", "refine_params": "{}"}) as w:
        w.put_masks("0001_0523", seg, dice)
    m = np.load(tmp_path / "masks" / "0001_0523.segmentation_mask.npy")
    assert m.dtype == np.bool_ and np.array_equal(m, seg.numpy())
    assert np.array_equal(np.load(tmp_path / "masks" / "0001_0523.estimated_dice.npy"), dice.numpy(), equal_nan=True)


def test_h5_layout_matches_reference_when_h5py_exists(tmp_path):
    h5py = pytest.importorskip("h5py")
    if getattr(h5py, "__version__", None) is None:
        pytest.skip("a stand-in h5py module is installed in sys.modules by the oracle-vs-reference tests")
    f = torch.randn((1, 256, 64, 64))
    with AsyncResultWriter(tmp_path / "e.h5", "embedding", {"checkpoint": "c.pth", "img_encoder_img_size": 1024}) as w:
        w.put_embedding("a", f, (10, 20), (1024, 512))
    with h5py.File(tmp_path / "e.h5") as h:
        assert h.attrs["checkpoint"] == "c.pth"
        assert np.array_equal(h["img_embedding/a/features"][:], f.numpy())
        assert h["img_embedding/a"].attrs["input_size"].tolist() == [1024, 512]


def test_writer_collects_in_memory_without_a_path():
    """path=None: records end up as numpy arrays in writer.backend.records (the host hand-over of the pipeline drivers)."""
    from samcarriestheburden_b200.storage import AsyncResultWriter
    w = AsyncResultWriter(None, "mask")
    seg = torch.zeros((3, 5, 7), dtype=torch.bool)
    seg[1, 2, 3] = True
    dice = torch.tensor([0.25, float("nan"), 0.75])
    for i in range(5):
        w.put_masks(f"m{i}", seg, dice + i)
    assert w.close() == 5
    rec = w.backend.records
    assert sorted(rec) == [f"m{i}" for i in range(5)]
    assert rec["m3"]["segmentation_mask"].dtype == np.bool_ and rec["m3"]["segmentation_mask"][1, 2, 3]
    assert np.allclose(rec["m3"]["estimated_dice"], (dice + 3).numpy(), equal_nan=True)

