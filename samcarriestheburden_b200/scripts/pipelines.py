"""Drivers mirroring the reference's two hot-path scripts, restructured for a B200 box:

  generate_img_embeddings (reference scripts/generate_img_embeddings.py:36-72)
      per image, B=1, H2D + encoder + D2H + gzip-9 h5 write, one fixed device
   -> batched encoder launches per GPU, images sharded by rank, embeddings kept resident in an EmbeddingStore
      (or gathered once at the end); disk I/O is left to the caller.

  refine_segmentations (reference scripts/save_refined_segmentations.py:60-80 after the U-Net)
      per image: prompt extraction + 2 B=1 decoder calls per class
   -> per BATCH of images: one prompt-extraction launch, two batched decoder passes over all classes of all
      images of the batch, one fused upscale/threshold/nearest-exact launch per distinct native size.

Host -> device traffic never sits in front of compute on the compute stream: image batches, native-resolution radiographs and
the U-Net maps go through pinned staging buffers and copy / side streams one step ahead of the kernels that consume them
(_HostBatchUploader, _NativeImageUploader, refine_segmentations.upload; B200SAM_UPLOAD_RINGS=0 restores pageable uploads for
A/B: set500 123.8 -> 128.6 images/s, native-resolution embedding 136 -> 143-153 images/s, refine phase +6 %).
"""
from __future__ import annotations

import os
from typing import Dict, Iterable, List, Sequence, Tuple

import numpy as np
import torch

from .. import sharding
from ..segment_anything.predictor import SamPredictor
from ..segment_anything.sam_mask_decoder_head import EmbeddingStore, SAMMaskDecoderHead
from ..utils.seg_refinement import SAMSegRefiner, SegEnhance


# B200SAM_UPLOAD_RINGS=0: pageable uploads on the compute stream (A/B of the two uploaders below)
_UPLOAD_RINGS = os.environ.get("B200SAM_UPLOAD_RINGS", "1") != "0"


class _HostBatchUploader:
    """Double-buffered upload of batches of same-shape host images (HWC uint8): memcpy into pinned memory, asynchronous H2D
    on its own stream, HWC -> CHW on the GPU.  The copy of batch i+1 runs on the copy engine while the encoder works on
    batch i (the launch loop is a batch ahead of the GPU); a pageable `tensor.to(device)` on the compute stream costs the
    encoder ~2 ms per batch of eight 1024 x 1024 images."""

    def __init__(self, device, batch: int, shape):
        self.device = torch.device(device)
        self.stream = torch.cuda.Stream(device=self.device)
        self.pinned = [torch.empty((batch,) + tuple(shape), dtype=torch.uint8).pin_memory() for _ in range(2)]
        self.devbuf = [torch.empty((batch,) + tuple(shape), dtype=torch.uint8, device=self.device) for _ in range(2)]
        self.copied = [None, None]    # H2D of this slot finished (the pinned buffer may be overwritten)
        self.consumed = [None, None]  # the compute stream has read this slot's device buffer
        self.k = 0

    def upload(self, images) -> torch.Tensor:
        k, n = self.k, len(images)
        self.k ^= 1
        if self.copied[k] is not None:
            self.copied[k].synchronize()
        for i, img in enumerate(images):
            self.pinned[k][i].numpy()[...] = img
        cur = torch.cuda.current_stream(self.device)
        with torch.cuda.stream(self.stream):
            if self.consumed[k] is not None:
                self.stream.wait_event(self.consumed[k])
            self.devbuf[k][:n].copy_(self.pinned[k][:n], non_blocking=True)
            self.copied[k] = self.stream.record_event()
        cur.wait_event(self.copied[k])
        x = self.devbuf[k][:n].permute(0, 3, 1, 2).contiguous()  # [n, 3, H, W] uint8
        self.consumed[k] = cur.record_event()
        return x


class _NativeImageUploader:
    """Native-resolution radiographs (HWC uint8 of any shape, e.g. 2570 x 2040 = 15.7 MB) on their way to the GPU resize:
    a ring of pinned staging + device buffers and a side stream that carries the H2D (0.6 ms at PCIe rates) AND the resize
    kernels of image i+1 while the encoder is busy with the previous batch on the compute stream."""

    SLOTS = 4

    def __init__(self, device):
        self.device = torch.device(device)
        self.stream = torch.cuda.Stream(device=self.device)
        self.cap = 0
        self.pinned, self.devbuf = [], []
        self.copied = [None] * self.SLOTS
        self.k = 0

    def _grow(self, nbytes: int) -> None:
        torch.cuda.synchronize(self.device)  # nothing in flight may still reference the old buffers
        self.cap = 1 << max(20, (nbytes - 1).bit_length())
        self.pinned = [torch.empty(self.cap, dtype=torch.uint8).pin_memory() for _ in range(self.SLOTS)]
        self.devbuf = [torch.empty(self.cap, dtype=torch.uint8, device=self.device) for _ in range(self.SLOTS)]
        self.copied = [None] * self.SLOTS

    def upload_resized(self, img: np.ndarray, transform) -> torch.Tensor:
        """H2D + `ResizeLongestSide.apply_image_cuda` of one image, both on the uploader's stream (so a ring slot is free
        again as soon as ITS resize has run, not when the compute stream gets to it behind a whole encoder batch).  Returns
        the resized [C, h, w] uint8 CUDA tensor; the caller's stream is made to wait for it."""
        img = np.ascontiguousarray(img)
        n = img.nbytes
        if n > self.cap:
            self._grow(n)
        k = self.k
        self.k = (k + 1) % self.SLOTS
        if self.copied[k] is not None:
            self.copied[k].synchronize()
        self.pinned[k][:n].numpy()[...] = img.reshape(-1)
        cur = torch.cuda.current_stream(self.device)
        with torch.cuda.stream(self.stream):
            self.devbuf[k][:n].copy_(self.pinned[k][:n], non_blocking=True)  # (stream order: after the slot's last resize)
            self.copied[k] = self.stream.record_event()
            t = transform.apply_image_cuda(self.devbuf[k][:n].view(img.shape), device=self.device, chw=True)
            ready = self.stream.record_event()
        t.record_stream(cur)
        cur.wait_event(ready)
        return t


@torch.no_grad()
def generate_img_embeddings(sam, images: Sequence[np.ndarray], names: Sequence[str], batch: int = 8,
                            store: EmbeddingStore | None = None, gather: bool = False, sam_type: str = "sam",
                            writer=None):
    """images: HWC uint8 RGB arrays (gray replicated to 3 channels like the reference :39-40).  Each rank encodes
    its shard in batches (same-shape images are batched together) and registers the results in `store`.
    sam_type = 'sam': SamPredictor.set_image semantics (:45-48); 'medsam' (the reference's default, :16): cubic resize to
    1024 x 1024 + min-max normalisation, the encoder is called directly without Sam.preprocess (:49-64).
    `writer`: an `storage.AsyncResultWriter(kind='embedding')`; every embedding is handed to it as soon as its batch is
    encoded (pinned D2H copy on the writer's stream + background write in the reference's layout, :67-70), off the
    critical path.  Returns (store, gathered [N,256,64,64] tensor or None)."""
    assert len(images) == len(names)
    if sam_type not in ("sam", "medsam"):
        raise NotImplementedError(f"Unknown SAM type: {sam_type}")
    if sam_type == "medsam":
        return _generate_medsam_embeddings(sam, images, names, batch, store, gather, writer)
    dev = sam.device
    store = store if store is not None else EmbeddingStore(img_encoder_img_size=sam.image_encoder.img_size)
    pred = SamPredictor(sam)
    mine = sharding.shard_indices(len(images))
    local = torch.empty((len(mine), 256, 64, 64), dtype=torch.float32, device=dev)
    pending: Dict[tuple, List[Tuple[int, object, tuple]]] = {}
    # (kept on the model: pinning the staging buffers costs ~20 ms, the staged pipeline calls this function per stage)
    uploaders = sam.__dict__.setdefault("_host_uploaders", {})

    def flush(bkey):
        items = pending.pop(bkey)
        key = bkey[1:]  # (h, w) of the encoder input = input_size of the record
        ts = [t for _, t, _ in items]
        if bkey[0]:
            # host HWC uint8 images already at the encoder's size: pinned staging + H2D on a copy stream (it overlaps the
            # previous batch's encoder), HWC -> CHW on the GPU
            ukey = (str(dev), batch, tuple(ts[0].shape))
            up = uploaders.get(ukey)
            if up is None:
                up = uploaders[ukey] = _HostBatchUploader(dev, batch, ts[0].shape)
            x = up.upload(ts)
        elif all(not t.is_cuda for t in ts):  # one stacked upload instead of one copy per image
            x = torch.stack(ts).to(dev, non_blocking=True)
        else:
            x = torch.stack([t.to(dev, non_blocking=True) for t in ts])
        emb = sam.encode_image(x)
        for j, (slot, _, orig) in enumerate(items):
            local[slot] = emb[j]
            store.add(names[mine[slot]], local[slot:slot + 1], orig, key)
            if writer is not None:
                writer.put_embedding(names[mine[slot]], local[slot:slot + 1], orig, key)

    for slot, i in enumerate(mine):
        img = images[i]
        target = pred.transform.get_preprocess_shape(img.shape[0], img.shape[1], pred.transform.target_length)
        if target == tuple(img.shape[:2]):  # already at the encoder's size: upload as is (batched, see flush)
            t = np.ascontiguousarray(img)
            if t.dtype != np.uint8 or t.ndim != 3 or not _UPLOAD_RINGS:
                t = torch.from_numpy(t).permute(2, 0, 1).contiguous()
        else:  # native-resolution radiograph: upload the uint8 pixels once, Pillow-exact resize on the GPU
            if isinstance(img, np.ndarray) and img.dtype == np.uint8 and _UPLOAD_RINGS:
                nat = uploaders.get(("native", str(dev)))
                if nat is None:
                    nat = uploaders[("native", str(dev))] = _NativeImageUploader(dev)
                t = nat.upload_resized(img, pred.transform)
            else:
                t = pred.transform.apply_image_cuda(img, device=dev, chw=True)
        host = isinstance(t, np.ndarray)
        bkey = (host,) + (tuple(t.shape[:2]) if host else tuple(t.shape[-2:]))
        pending.setdefault(bkey, []).append((slot, t, tuple(img.shape[:2])))
        if len(pending[bkey]) == batch:
            flush(bkey)
    for bkey in list(pending):
        flush(bkey)
    gathered = sharding.gather_sharded(local, len(images)) if gather else None
    return store, gathered


def _generate_medsam_embeddings(sam, images, names, batch, store, gather, writer=None):
    """scripts/generate_img_embeddings.py:49-64: per image cv2 INTER_CUBIC resize (bit-exact GPU restatement) + min-max
    normalise, then `image_encoder(img_tensor)` on batches of the resulting float tensors; original_size = the native
    size, input_size = (1024, 1024)."""
    from ..segment_anything.utils.transforms import medsam_preprocess_cuda
    dev = sam.device
    size = sam.image_encoder.img_size
    store = store if store is not None else EmbeddingStore(img_encoder_img_size=size)
    mine = sharding.shard_indices(len(images))
    local = torch.empty((len(mine), 256, 64, 64), dtype=torch.float32, device=dev)
    for j in range(0, len(mine), batch):
        chunk = mine[j:j + batch]
        xs = []
        for i in chunk:
            img = images[i]
            g = torch.from_numpy(np.ascontiguousarray(img if img.ndim == 2 else img[..., 0])).to(dev, non_blocking=True)
            xs.append(medsam_preprocess_cuda(g, size))
        emb = sam.image_encoder(torch.cat(xs))
        for k, i in enumerate(chunk):
            local[j + k] = emb[k]
            store.add(names[i], local[j + k:j + k + 1], tuple(images[i].shape[:2]), (size, size))
            if writer is not None:
                writer.put_embedding(names[i], local[j + k:j + k + 1], tuple(images[i].shape[:2]), (size, size))
    gathered = sharding.gather_sharded(local, len(images)) if gather else None
    return store, gathered


@torch.no_grad()
def predict_unet_probabilities(unet, images: Sequence[np.ndarray], size=(384, 224), batch: int = 8,
                               mean: float = 0.3505533917353781, std: float = 0.22763733675869177) -> List[torch.Tensor]:
    """The U-Net stage of save_refined_segmentations.py:61-69 for the local shard: grey uint8 images ->
    cv2.resize(INTER_LINEAR) to (H, W) (OpenCV's fixed-point uint8 path, restated bit-exactly on the GPU: the result is
    ROUNDED to uint8 before / 255, like the reference) -> / 255 -> normalise -> U-Net -> sigmoid.
    Returns one [C, H, W] probability map per input image (on the model's device)."""
    from ..segment_anything.utils.transforms import cv_resize_linear_cuda
    dev = unet.outc.conv.weight.device
    H, W = size
    out: List[torch.Tensor] = []
    for j in range(0, len(images), batch):
        xs = []
        for img in images[j:j + batch]:
            g = torch.from_numpy(np.ascontiguousarray(img if img.ndim == 2 else img[..., 0])).to(dev, non_blocking=True)
            # cv2.resize returns the input unchanged when the size already matches; the kernel's identity tables do too
            xs.append(cv_resize_linear_cuda(g, H, W, normalize=(mean, std))[None, None])
        probs = unet.predict_proba(torch.cat(xs))
        out.extend(probs[k] for k in range(probs.shape[0]))
    return out


@torch.no_grad()
def refine_segmentations(sam, store: EmbeddingStore, segs: Sequence[torch.Tensor], names: Sequence[str],
                         prompts2use=(("box",), ("pos_points", "neg_points")), gather: bool = False,
                         batch: int = 8, ccl_selection: str | None = None, writer=None):
    """segs[i]: [C,H,W] bool (or probabilities) U-Net masks of image names[i]; every rank refines the images of its
    shard whose embeddings it holds.  `ccl_selection` ('highest_probability' | 'largest') runs the SegEnhance
    connected-component pre-processing (save_refined_segmentations.py:25-33,76) on the probability maps first.
    `writer`: a `storage.AsyncResultWriter(kind='mask')` that persists every image's refined masks + estimated Dice
    (save_refined_segmentations.py:75-80) off the critical path.
    Returns (list of (index, seg bool [C,H,W], est_dice [C]) for the local shard, gathered [N,C,H,W] uint8 tensor
    or None)."""
    dev = sam.device
    head = SAMMaskDecoderHead(None, "", str(dev), store, sam_model=sam)
    refiner = SAMSegRefiner("SAM", str(dev), [list(p) for p in prompts2use], sam_predictor=head)
    stage = SegEnhance(refiner, ccl_selection, "dilation", "square", 0, str(dev)) if ccl_selection else None
    mine = sharding.shard_indices(len(segs))
    results = []
    # the U-Net maps of batch j+1 (17 x 384 x 224 fp32 = 5.8 MB per image) are uploaded on a copy stream while batch j is
    # refined (the refinement has host round trips, so the upload must be ISSUED before it starts); sources that are not
    # pinned simply copy synchronously as before
    copy_stream = sam.__dict__.setdefault("_seg_upload_stream", None)
    if copy_stream is None or copy_stream.device != torch.device(dev):
        copy_stream = sam.__dict__["_seg_upload_stream"] = torch.cuda.Stream(device=dev)
    cur = torch.cuda.current_stream(dev)

    def upload(chunk):
        if not _UPLOAD_RINGS or any(segs[i].is_cuda or segs[i].shape != segs[chunk[0]].shape for i in chunk):
            return torch.stack([segs[i].to(dev) for i in chunk]), None
        with torch.cuda.stream(copy_stream):
            x = torch.empty((len(chunk),) + tuple(segs[chunk[0]].shape), dtype=segs[chunk[0]].dtype, device=dev)
            for k, i in enumerate(chunk):
                x[k].copy_(segs[i], non_blocking=True)
            ev = copy_stream.record_event()
        x.record_stream(cur)
        return x, ev

    chunks = [mine[j:j + batch] for j in range(0, len(mine), batch)]
    nxt = upload(chunks[0]) if chunks else None
    for ci, chunk in enumerate(chunks):
        (x, ev), nm = nxt, [names[i] for i in chunk]
        nxt = upload(chunks[ci + 1]) if ci + 1 < len(chunks) else None
        if ev is not None:
            cur.wait_event(ev)
        seg_b, est_b = stage.enhance_batch(x, nm) if stage is not None else refiner.refine_batch(x, nm)
        results.extend((i, seg_b[k], est_b[k]) for k, i in enumerate(chunk))
        if writer is not None:
            for k, i in enumerate(chunk):
                writer.put_masks(names[i], seg_b[k], est_b[k])
    gathered = None
    if gather and len(segs):
        local = torch.stack([r[1] for r in results]).to(torch.uint8) if results else \
            torch.zeros((0,) + tuple(segs[0].shape), dtype=torch.uint8, device=dev)
        gathered = sharding.gather_sharded(local, len(segs))
    return results, gathered


@torch.no_grad()
def embed_and_refine(sam, images: Sequence[np.ndarray], segs: Sequence[torch.Tensor], names: Sequence[str],
                     prompts2use=(("box",), ("pos_points", "neg_points")), batch: int = 8, stage: int = 32,
                     ccl_selection: str | None = None, gather: bool = False, emb_writer=None, mask_writer=None,
                     overlap: bool = True):
    """Both scripts of the reference as ONE software pipeline (the reference runs them one after the other and goes through
    an h5 file in between): the image set is cut into stages of `stage` images; the encoder of stage s + 1 is enqueued on its
    own CUDA stream before the refinement of stage s (prompt extraction, two decoder passes, upscale) is issued on a second
    stream.  Results and sharding are those of generate_img_embeddings + refine_segmentations called on every stage (bit-
    identical, tests/test_model_gpu.py).  Measured (tools/overlap_probe.py, ViT-H, 192 images, one B200): 121 images/s for
    every stage size and with or without the second stream, i.e. the same as the two phases back to back - the encoder
    saturates the GPU and the decode stage leaves no idle time worth filling; the function exists for the single-call API
    and for bounded embedding residency (`stage` images instead of the whole set), not for speed.
    Returns (store, results, gathered embeddings or None, gathered masks or None)."""
    assert len(images) == len(segs) == len(names)
    dev = sam.device
    # the two streams are kept on the model: torch's caching allocator pools blocks per stream, so fresh streams would
    # send every temporary of every call to cudaMalloc
    streams = getattr(sam, "_pipeline_streams", None)
    if streams is None or streams[0].device != torch.device(dev):
        streams = (torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev))
        sam._pipeline_streams = streams
    enc_stream = streams[0]
    ref_stream = streams[1] if overlap else enc_stream  # overlap=False: staged, one stream (A/B)
    start = torch.cuda.current_stream(dev).record_event()
    store = EmbeddingStore(img_encoder_img_size=sam.image_encoder.img_size)
    bounds = [(a, min(a + stage, len(images))) for a in range(0, len(images), stage)]
    encoded, results, emb_parts, seg_parts = [], [], [], []

    def encode(k):
        a, b = bounds[k]
        with torch.cuda.stream(enc_stream):
            if k == 0:
                enc_stream.wait_event(start)
            _, g = generate_img_embeddings(sam, images[a:b], names[a:b], batch=batch, store=store, gather=gather,
                                           writer=emb_writer)
            emb_parts.append(g)
            encoded.append(enc_stream.record_event())

    encode(0)
    for k, (a, b) in enumerate(bounds):
        if k + 1 < len(bounds):
            encode(k + 1)
        with torch.cuda.stream(ref_stream):
            ref_stream.wait_event(encoded[k])
            res, g = refine_segmentations(sam, store, segs[a:b], names[a:b], prompts2use=prompts2use, gather=gather,
                                          batch=batch, ccl_selection=ccl_selection, writer=mask_writer)
            results.extend((a + i, seg, dice) for i, seg, dice in res)
            seg_parts.append(g)
    cur = torch.cuda.current_stream(dev)
    cur.wait_stream(enc_stream)
    cur.wait_stream(ref_stream)
    emb_all = torch.cat(emb_parts) if gather and emb_parts else None
    seg_all = torch.cat(seg_parts) if gather and seg_parts else None
    return store, results, emb_all, seg_all

