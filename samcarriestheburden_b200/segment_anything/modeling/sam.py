"""`Sam` container (reference: segment_anything/modeling/sam.py) with the fused CUDA entry points."""
from __future__ import annotations

import ctypes as C
import weakref
from typing import Any, Dict, List, Optional, Tuple

import torch
from torch import nn

from ... import _lib
from .image_encoder import ImageEncoderViT
from .mask_decoder import MaskDecoder
from .prompt_encoder import PromptEncoder


def _pack_decoder_weight(sd: Dict[str, torch.Tensor], spec: str) -> torch.Tensor:
    key, _, packing = spec.partition("|")
    if packing == "cat4":  # point_embeddings.{0..3}.weight -> [4, 256]
        return torch.cat([sd[f"{key}.{i}.weight"] for i in range(4)], dim=0).float().contiguous()
    t = sd[key].detach().float()
    if packing == "convT":  # ConvTranspose2d [Cin, Cout, 2, 2] -> [(dy*2+dx)*Cout + co, ci]
        return t.permute(2, 3, 1, 0).reshape(-1, t.shape[0]).contiguous()
    if packing == "repeat4":
        return t.repeat(4).contiguous()
    if packing:
        raise ValueError(f"unknown packing {packing}")
    return t.contiguous().clone()  # never alias the live parameter


class _DecoderEngine:
    """Packed fp32 decoder weights + the C-side handle + workspaces for one device."""

    def __init__(self, sam: "Sam", device: torch.device) -> None:
        lib = _lib.load()
        self.lib, self.device = lib, device
        sd = {k: v.to(device) for k, v in sam.state_dict().items()
              if k.startswith("prompt_encoder.") or k.startswith("mask_decoder.")}
        n = lib.b200sam_decoder_weight_count()
        self.packed = [_pack_decoder_weight(sd, lib.b200sam_decoder_weight_name(i).decode()) for i in range(n)]
        arr = (C.c_void_p * n)(*[t.data_ptr() for t in self.packed])
        handle = C.c_void_p()
        _lib.run(device, lib.b200sam_decoder_create, arr, n, C.byref(handle), what="b200sam_decoder_create")
        self.handle = handle
        self._ws: Optional[torch.Tensor] = None
        self._pe: Optional[torch.Tensor] = None

    def dense_pe(self) -> torch.Tensor:
        if self._pe is None:
            tok = torch.empty((4096, 256), dtype=torch.float32, device=self.device)
            _lib.run(self.device, self.lib.b200sam_decoder_copy_dense_pe, self.handle, tok.data_ptr(), what="b200sam_decoder_copy_dense_pe")
            self._pe = tok.view(64, 64, 256).permute(2, 0, 1).unsqueeze(0)
        return self._pe

    def decode(self, embedding: torch.Tensor, coords: Optional[torch.Tensor], labels: Optional[torch.Tensor],
               mask_prev: Optional[torch.Tensor], multimask: bool,
               image_of: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
        """embedding [n_img,256,64,64] (or one image in any [.., 256,64,64] view); coords [NB,Np,2] fp32; labels
        [NB,Np] int32 (-2 = absent trailing slot); mask_prev [NB,1,256,256]; image_of [NB] int32 image index of
        every prompt (required when n_img > 1)."""
        emb = embedding.reshape(-1, 256, 64, 64).float().contiguous()
        n_img = emb.shape[0]
        NB = coords.shape[0] if coords is not None else (mask_prev.shape[0] if mask_prev is not None else 1)
        Np = coords.shape[1] if coords is not None else 0
        nm = 3 if multimask else 1
        low = torch.empty((NB, nm, 256, 256), dtype=torch.float32, device=self.device)
        iou = torch.empty((NB, nm), dtype=torch.float32, device=self.device)
        need = self.lib.b200sam_decoder_workspace_bytes_batch(n_img, NB, Np)
        if self._ws is None or self._ws.numel() < need + 256:
            self._ws = None  # release before growing
            self._ws = torch.empty(need + 256, dtype=torch.uint8, device=self.device)
        base = (self._ws.data_ptr() + 255) & ~255
        if coords is not None:
            coords = coords.float().contiguous()
            labels = labels.to(torch.int32).contiguous()
        if mask_prev is not None:
            mask_prev = mask_prev.reshape(NB, 256, 256).float().contiguous()
        if image_of is not None:
            image_of = image_of.to(device=self.device, dtype=torch.int32).contiguous()
            assert image_of.numel() == NB, "image_of must name one image per prompt"
        elif n_img != 1:
            raise ValueError("image_of is required when decoding prompts of more than one image")
        _lib.run(self.device, self.lib.b200sam_decode_batch, self.handle, emb.data_ptr(), n_img, _lib.ptr(image_of), NB, Np,
                                                 _lib.ptr(coords), _lib.ptr(labels), _lib.ptr(mask_prev),
                                                 int(multimask), low.data_ptr(), iou.data_ptr(), base,
                                                 self._ws.numel() - (base - self._ws.data_ptr()), what="b200sam_decode_batch")
        return low, iou

    def _workspace(self, n_img: int, NB: int, Np: int):
        need = self.lib.b200sam_decoder_workspace_bytes_batch(n_img, NB, Np)
        if self._ws is None or self._ws.numel() < need + 256:
            self._ws = None  # release before growing
            self._ws = torch.empty(need + 256, dtype=torch.uint8, device=self.device)
        base = (self._ws.data_ptr() + 255) & ~255
        return base, self._ws.numel() - (base - self._ws.data_ptr())

    def prompt_encode(self, coords: Optional[torch.Tensor], labels: Optional[torch.Tensor], mask_prev: Optional[torch.Tensor],
                      batch: int) -> Tuple[torch.Tensor, torch.Tensor]:
        """PromptEncoder.forward alone: -> (sparse [B,Np,256], dense [B,256,64,64] as an NCHW view of token-major data)."""
        NB = batch
        Np = coords.shape[1] if coords is not None else 0
        sparse = torch.empty((NB, Np, 256), dtype=torch.float32, device=self.device)
        dense_tok = torch.empty((NB, 4096, 256), dtype=torch.float32, device=self.device)
        tmp = torch.empty((NB, 5 + Np, 256), dtype=torch.float32, device=self.device) if Np else None
        ntok = torch.empty((NB,), dtype=torch.int32, device=self.device) if Np else None
        if coords is not None:
            coords = coords.float().contiguous()
            labels = labels.to(torch.int32).contiguous()
        if mask_prev is not None:
            mask_prev = mask_prev.reshape(NB, 256, 256).float().contiguous()
        _lib.run(self.device, self.lib.b200sam_prompt_encode, self.handle, _lib.ptr(coords), _lib.ptr(labels), NB, Np,
                 _lib.ptr(mask_prev), _lib.ptr(tmp), _lib.ptr(ntok), _lib.ptr(sparse), dense_tok.data_ptr(),
                 what="b200sam_prompt_encode")
        return sparse, dense_tok.view(NB, 64, 64, 256).permute(0, 3, 1, 2)

    def decode_embedded(self, embedding: torch.Tensor, sparse: torch.Tensor, dense: torch.Tensor,
                        multimask: bool) -> Tuple[torch.Tensor, torch.Tensor]:
        """MaskDecoder.predict_masks + slicing on caller-supplied sparse [B,N,256] / dense [B,256,64,64] embeddings."""
        emb = embedding.reshape(-1, 256, 64, 64).float().contiguous()
        n_img, NB, Ns = emb.shape[0], sparse.shape[0], sparse.shape[1]
        if n_img not in (1, NB):
            raise ValueError("image_embeddings must hold one image or one image per prompt (mask_decoder.py:125)")
        image_of = torch.arange(NB, dtype=torch.int32, device=self.device) if n_img > 1 else None
        sp = sparse.float().contiguous()
        dt = dense.float().expand(NB, 256, 64, 64).permute(0, 2, 3, 1).reshape(NB, 4096, 256).contiguous()
        nm = 3 if multimask else 1
        low = torch.empty((NB, nm, 256, 256), dtype=torch.float32, device=self.device)
        iou = torch.empty((NB, nm), dtype=torch.float32, device=self.device)
        base, size = self._workspace(n_img, NB, Ns)
        _lib.run(self.device, self.lib.b200sam_decode_embedded, self.handle, emb.data_ptr(), n_img, _lib.ptr(image_of), NB,
                 Ns, _lib.ptr(sp) if Ns else None, dt.data_ptr(), int(multimask), low.data_ptr(), iou.data_ptr(), base,
                 size, what="b200sam_decode_embedded")
        return low, iou

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                self.lib.b200sam_decoder_destroy(self.handle)
        except Exception:
            pass


def assemble_prompt_points(point_coords: Optional[torch.Tensor], point_labels: Optional[torch.Tensor],
                           boxes: Optional[torch.Tensor]):
    """[points | pad point iff no box | box corners] + labels (-1 pad, 0/1 points, 2/3 corners), as the prompt encoder
    consumes them (prompt_encoder.py:73-100,155).  Built on the inputs' own device."""
    coords, labels = [], []
    if point_coords is not None:
        pc = point_coords.float()
        pl = point_labels.to(device=pc.device, dtype=torch.int32)
        coords.append(pc)
        labels.append(pl)
        if boxes is None:  # pad point, label -1 (prompt_encoder.py:81-85)
            coords.append(torch.zeros((pc.shape[0], 1, 2), device=pc.device))
            labels.append(-torch.ones((pc.shape[0], 1), dtype=torch.int32, device=pc.device))
    if boxes is not None:
        b = boxes.float().reshape(-1, 2, 2)
        if coords:
            b = b.to(coords[0].device)
        coords.append(b)
        labels.append(torch.tensor([[2, 3]], dtype=torch.int32, device=b.device).expand(b.shape[0], 2))
    if not coords:
        return None, None
    return torch.cat(coords, dim=1), torch.cat(labels, dim=1)


def upscale_masks(low_res: torch.Tensor, input_size, original_size, img_size: int = 1024, threshold: float = 0.0,
                  return_logits: bool = False, small_size=None):
    """Fused postprocess_masks (+ threshold, + nearest-exact tap) on [B,C,256,256] logits via the C ABI."""
    lib = _lib.load()
    B, Cn, L, _ = low_res.shape
    low = low_res.float().contiguous()
    oh, ow = int(original_size[0]), int(original_size[1])
    dev = low.device
    mask = None if return_logits else torch.empty((B, Cn, oh, ow), dtype=torch.bool, device=dev)
    logits = torch.empty((B, Cn, oh, ow), dtype=torch.float32, device=dev) if return_logits else None
    small = None
    sh = sw = 0
    if small_size is not None:
        sh, sw = int(small_size[0]), int(small_size[1])
        small = torch.empty((B, Cn, sh, sw), dtype=torch.bool, device=dev)
    _lib.run(dev, lib.b200sam_upscale_threshold, low.data_ptr(), B * Cn, L, img_size, int(input_size[0]), int(input_size[1]),
                                             oh, ow, float(threshold), _lib.ptr(mask), _lib.ptr(logits), _lib.ptr(small),
                                             sh, sw, what="b200sam_upscale_threshold")
    out = logits if return_logits else mask
    return (out, small) if small_size is not None else out


class Sam(nn.Module):
    mask_threshold: float = 0.0
    image_format: str = "RGB"

    def __init__(self, image_encoder: ImageEncoderViT, prompt_encoder: PromptEncoder, mask_decoder: MaskDecoder,
                 pixel_mean: List[float] = [123.675, 116.28, 103.53],
                 pixel_std: List[float] = [58.395, 57.12, 57.375]) -> None:
        super().__init__()
        self.image_encoder = image_encoder
        self.prompt_encoder = prompt_encoder
        self.mask_decoder = mask_decoder
        self.register_buffer("pixel_mean", torch.Tensor(pixel_mean).view(-1, 1, 1), False)
        self.register_buffer("pixel_std", torch.Tensor(pixel_std).view(-1, 1, 1), False)
        self._pixel_mean_host = tuple(float(v) for v in pixel_mean)
        self._pixel_std_host = tuple(float(v) for v in pixel_std)
        self._dec_engine: Optional[_DecoderEngine] = None
        self._dec_versions: Optional[tuple] = None
        self.prompt_encoder._owner = weakref.ref(self)
        self.mask_decoder._owner = weakref.ref(self)
        self.register_load_state_dict_post_hook(lambda module, incompatible: module._invalidate())

    def _invalidate(self) -> None:
        self._dec_engine = None

    @property
    def device(self) -> Any:
        return self.pixel_mean.device

    def _apply(self, fn, *a, **k):
        self._invalidate()
        return super()._apply(fn, *a, **k)

    def decoder_engine(self) -> _DecoderEngine:
        dev = self.pixel_mean.device
        if dev.type != "cuda":
            raise _lib.B200SamError("b200sam has no CPU path: move the model to a CUDA device")
        versions = tuple(p._version for m in (self.prompt_encoder, self.mask_decoder) for p in m.parameters())
        if self._dec_engine is None or self._dec_engine.device != dev or versions != self._dec_versions:
            self._dec_engine = None
            self._dec_engine = _DecoderEngine(self, dev)
            self._dec_versions = versions
        return self._dec_engine

    # ---- fused entry points -------------------------------------------------------------------------
    @torch.no_grad()
    def encode_image(self, transformed_image: torch.Tensor) -> torch.Tensor:
        """preprocess (sam.py:164-174) + image_encoder (image_encoder.py:106-116) in one CUDA pipeline.
        transformed_image: [B,3,h,w] uint8 or float, un-normalised, long side == 1024."""
        return self.image_encoder.forward_raw(transformed_image.to(self.device), self._pixel_mean_host,
                                              self._pixel_std_host)

    @torch.no_grad()
    def decode_prompts(self, features: torch.Tensor, point_coords: Optional[torch.Tensor],
                       point_labels: Optional[torch.Tensor], boxes: Optional[torch.Tensor],
                       mask_input: Optional[torch.Tensor], multimask_output: bool) -> Tuple[torch.Tensor, torch.Tensor]:
        """PromptEncoder.forward + MaskDecoder.forward for a batch of prompts of ONE image
        (prompt_encoder.py:128-168, mask_decoder.py:71-110).  Returns (low_res [B,1|3,256,256], iou [B,1|3])."""
        dev = self.device
        # assemble [points | pad point | box corners] + labels on the inputs' own device (CPU tensors stay on the
        # host and cross in ONE copy each), then hand raw pointers to the C ABI
        c, l = assemble_prompt_points(point_coords, point_labels, boxes)
        c = c.to(dev, non_blocking=True) if c is not None else None
        l = l.to(dev, non_blocking=True) if l is not None else None
        m = mask_input.to(dev) if mask_input is not None else None
        return self.decoder_engine().decode(features, c, l, m, multimask_output)

    # ---- reference API ------------------------------------------------------------------------------
    @torch.no_grad()
    def forward(self, batched_input: List[Dict[str, Any]], multimask_output: bool) -> List[Dict[str, torch.Tensor]]:
        """Batched end-to-end prediction (reference sam.py:53-131).  Images of the same (already transformed) shape are
        encoded together in ONE encoder launch sequence (the reference stacks all of them, :97-98); prompts are decoded per
        record like the reference's loop (:101-124)."""
        groups: Dict[Tuple[int, ...], List[int]] = {}
        for i, rec in enumerate(batched_input):
            groups.setdefault(tuple(rec["image"].shape), []).append(i)
        embeddings: List[Optional[torch.Tensor]] = [None] * len(batched_input)
        for idx in groups.values():
            emb = self.encode_image(torch.stack([batched_input[i]["image"] for i in idx]))
            for k, i in enumerate(idx):
                embeddings[i] = emb[k:k + 1]
        outputs = []
        for rec, emb in zip(batched_input, embeddings):
            low, iou = self.decode_prompts(emb, rec.get("point_coords"), rec.get("point_labels"), rec.get("boxes"),
                                           rec.get("mask_inputs"), multimask_output)
            masks = upscale_masks(low, rec["image"].shape[-2:], rec["original_size"], self.image_encoder.img_size,
                                  self.mask_threshold)
            outputs.append({"masks": masks, "iou_predictions": iou, "low_res_logits": low})
        return outputs

    def postprocess_masks(self, masks: torch.Tensor, input_size: Tuple[int, ...],
                          original_size: Tuple[int, ...]) -> torch.Tensor:
        """Upscale to 1024, crop the padding, resize to the original size (reference sam.py:133-162) — fused."""
        return upscale_masks(masks, input_size, original_size, self.image_encoder.img_size, return_logits=True)

    def preprocess(self, x: torch.Tensor) -> torch.Tensor:
        """Normalize pixel values and pad to a square input (reference sam.py:164-174).  Kept for API parity;
        the product path fuses this into the encoder (encode_image)."""
        x = (x - self.pixel_mean) / self.pixel_std
        h, w = x.shape[-2:]
        return torch.nn.functional.pad(x, (0, self.image_encoder.img_size - w, 0, self.image_encoder.img_size - h))
