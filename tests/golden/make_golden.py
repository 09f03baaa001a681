"""Generate golden fixtures by running the REFERENCE's own code (imported from /root/reference).

Run in the build container only:  python tests/golden/make_golden.py
The reference has no tests / golden vectors of its own, so these fixtures — inputs plus
the reference's outputs on them — are what pins oracle/sam_oracle.py.  A tiny SAM (same module classes, small
dims, 256-px input) keeps the fixture small; prompt extraction / scaling / post-processing use real sizes.
Storage / pre-processing deps the reference imports but that are absent here (h5py, kornia, skimage, pyamg,
cv2 is present) are stubbed: they carry no arithmetic of this path.
"""
import sys
import types
from pathlib import Path

import numpy as np
import torch

REF = "/root/reference"
OUT = Path(__file__).resolve().parent
sys.path.insert(0, REF)
sys.path.insert(0, str(OUT.parents[1]))

for name in ("h5py", "kornia", "kornia.contrib", "kornia.morphology", "skimage", "skimage.morphology", "pyamg"):
    if name not in sys.modules:
        sys.modules[name] = types.ModuleType(name)
sys.modules["kornia.contrib"].connected_components = None
for fn in ("dilation", "erosion"):
    setattr(sys.modules["kornia.morphology"], fn, None)
for fn in ("square", "disk", "diamond", "star"):
    setattr(sys.modules["skimage.morphology"], fn, None)

from functools import partial  # noqa: E402

from segment_anything.modeling import ImageEncoderViT, MaskDecoder, PromptEncoder, Sam, TwoWayTransformer  # noqa: E402
from segment_anything.utils.prompt_utils import PromptExtractor, scale_box, scale_coords  # noqa: E402
from segment_anything.utils.transforms import ResizeLongestSide  # noqa: E402

from oracle import sam_oracle as O  # noqa: E402


def tiny_sam(seed=0):
    torch.manual_seed(seed)
    D, img, patch = 32, 256, 16
    emb = img // patch
    sam = Sam(
        image_encoder=ImageEncoderViT(depth=2, embed_dim=D, img_size=img, mlp_ratio=4,
                                      norm_layer=partial(torch.nn.LayerNorm, eps=1e-6), num_heads=2, patch_size=patch,
                                      qkv_bias=True, use_rel_pos=True, global_attn_indexes=[1], window_size=14,
                                      out_chans=32),
        prompt_encoder=PromptEncoder(embed_dim=32, image_embedding_size=(emb, emb), input_image_size=(img, img),
                                     mask_in_chans=16),
        mask_decoder=MaskDecoder(num_multimask_outputs=3,
                                 transformer=TwoWayTransformer(depth=2, embedding_dim=32, mlp_dim=64, num_heads=2),
                                 transformer_dim=32, iou_head_depth=3, iou_head_hidden_dim=32),
    ).eval()
    with torch.no_grad():  # zero at init in the reference -> randomise so the rel-pos path is exercised
        sam.image_encoder.pos_embed.normal_(0, 0.02)
        for blk in sam.image_encoder.blocks:
            blk.attn.rel_pos_h.normal_(0, 0.02)
            blk.attn.rel_pos_w.normal_(0, 0.02)
        for m in sam.modules():
            if isinstance(m, torch.nn.LayerNorm) or m.__class__.__name__ == "LayerNorm2d":
                m.weight.normal_(1.0, 0.1)
                m.bias.normal_(0.0, 0.1)
    return sam


@torch.no_grad()
def main():
    out = {}
    sam = tiny_sam()
    sd = sam.state_dict()
    for k, v in sd.items():
        out["sd/" + k] = v.numpy()

    # ---- encoder + preprocess on a non-square input (exercises zero padding after normalisation)
    rng = np.random.default_rng(0)
    img = rng.integers(0, 256, size=(3, 200, 256)).astype(np.float32)
    x = sam.preprocess(torch.from_numpy(img))
    feats = sam.image_encoder(x[None])
    out["enc/image"] = img.astype(np.uint8)
    out["enc/features"] = feats.numpy()

    # ---- prompt encoder + decoder: box pass, then point(+pad)+mask pass, then multimask
    boxes = torch.tensor([[30.0, 40.0, 180.0, 150.0], [5.0, 5.0, 100.0, 220.0]])
    sp, de = sam.prompt_encoder(points=None, boxes=boxes, masks=None)
    low1, iou1 = sam.mask_decoder(image_embeddings=feats, image_pe=sam.prompt_encoder.get_dense_pe(),
                                  sparse_prompt_embeddings=sp, dense_prompt_embeddings=de, multimask_output=False)
    out["dec/boxes"] = boxes.numpy(); out["dec/sparse1"] = sp.numpy(); out["dec/low1"] = low1.numpy(); out["dec/iou1"] = iou1.numpy()
    out["dec/dense_pe"] = sam.prompt_encoder.get_dense_pe().numpy()
    pts = torch.tensor([[[60.0, 70.0], [10.0, 200.0], [128.5, 3.25]], [[200.0, 100.0], [90.0, 90.0], [1.0, 1.0]]])
    labs = torch.tensor([[1, 0, 0], [1, 0, 0]], dtype=torch.int)
    sp2, de2 = sam.prompt_encoder(points=(pts, labs), boxes=None, masks=low1)
    low2, iou2 = sam.mask_decoder(image_embeddings=feats, image_pe=sam.prompt_encoder.get_dense_pe(),
                                  sparse_prompt_embeddings=sp2, dense_prompt_embeddings=de2, multimask_output=False)
    low3, iou3 = sam.mask_decoder(image_embeddings=feats, image_pe=sam.prompt_encoder.get_dense_pe(),
                                  sparse_prompt_embeddings=sp2, dense_prompt_embeddings=de2, multimask_output=True)
    out["dec/points"] = pts.numpy(); out["dec/labels"] = labs.numpy()
    out["dec/sparse2"] = sp2.numpy(); out["dec/dense2"] = de2.numpy()
    out["dec/low2"] = low2.numpy(); out["dec/iou2"] = iou2.numpy(); out["dec/low3"] = low3.numpy(); out["dec/iou3"] = iou3.numpy()

    # ---- post-processing (real sizes: 256-logits -> 1024 -> crop -> native) + threshold
    logits = torch.from_numpy(rng.normal(0, 1, size=(2, 1, 256, 256)).astype(np.float32))
    out["post/logits"] = logits.numpy()
    real = Sam.__new__(Sam)  # only needs image_encoder.img_size for postprocess_masks
    torch.nn.Module.__init__(real)
    real.image_encoder = types.SimpleNamespace(img_size=1024)
    for i, (orig, inp) in enumerate([((1182, 754), None), ((578, 881), None), ((1024, 1024), None)]):
        inp = ResizeLongestSide.get_preprocess_shape(orig[0], orig[1], 1024)
        m = Sam.postprocess_masks(real, logits, inp, orig)
        out[f"post/{i}/orig"] = np.array(orig); out[f"post/{i}/inp"] = np.array(inp)
        out[f"post/{i}/mask"] = np.packbits((m > 0.0).numpy())
        out[f"post/{i}/sample"] = m[:, :, ::37, ::41].numpy()
        small = torch.nn.functional.interpolate((m > 0.0).float(), size=(384, 224), mode="nearest-exact")
        out[f"post/{i}/small"] = np.packbits(small.numpy() > 0.5)

    # ---- prompt extraction + coordinate scaling on synthetic U-Net masks (17 x 384 x 224)
    for i in range(6):
        masks = O.synthetic_unet_masks(i)
        if i == 4:
            masks[3] = masks[5]  # class fully overlapped by another -> no seed but a box
        if i == 5:
            masks[:, ::2, :] &= rng.random(masks[:, ::2, :].shape) < 0.5  # ragged masks, .5 rounding cases
        ex = PromptExtractor(torch.from_numpy(masks))
        prompts = ex.extract()
        out[f"pe/{i}/masks"] = np.packbits(masks)
        out[f"pe/{i}/classes"] = np.array([p.class_idx for p in prompts], np.int32)
        out[f"pe/{i}/pos"] = np.stack([p.pos_seeds.numpy() for p in prompts]).astype(np.int32)
        out[f"pe/{i}/neg"] = np.stack([p.neg_seeds.numpy() for p in prompts]).astype(np.int32)
        out[f"pe/{i}/box"] = np.stack([p.box.numpy() for p in prompts]).astype(np.int32)
        out[f"pe/{i}/boxes_all"] = np.stack([(ex._extract_box(c) if ex._extract_box(c) is not None
                                              else torch.zeros(4, dtype=torch.int)).numpy() for c in range(17)])
        p0 = prompts[0]
        out[f"pe/{i}/pos_scaled"] = scale_coords(p0.neg_seeds, p0.img_size, (1024, 653)).numpy()
        out[f"pe/{i}/box_scaled"] = scale_box(p0.box.unsqueeze(0), p0.img_size, (1024, 653)).numpy()
    np.savez_compressed(OUT / "reference_golden.npz", **out)
    print("wrote", OUT / "reference_golden.npz", sum(v.nbytes for v in out.values()) / 1e6, "MB raw")


if __name__ == "__main__":
    main()
