"""Resize helpers (reference: segment_anything/utils/transforms.py).  Semantics kept: PIL bilinear (antialiased)
uint8 resize so the long side is `target_length`, `int(x * scale + 0.5)` rounding.  `apply_image` is the reference's
host path; `apply_image_cuda` / `resize_u8_cuda` run the same arithmetic (Pillow's 22-bit fixed-point resample,
bit-exact) on the GPU through `b200sam_resize_u8` so that native-resolution radiographs are uploaded once as uint8
and never resized on the host."""
import ctypes as C
from copy import deepcopy
from typing import Dict, Tuple

import numpy as np
import torch
from torch.nn import functional as F

_TABLES: Dict[Tuple[int, int, str], Tuple[torch.Tensor, torch.Tensor, int]] = {}


def _resize_tables(in_size: int, out_size: int, device: torch.device):
    """Fixed-point coefficient tables of one axis, built by the library's host function and cached on the device."""
    from ... import _lib
    key = (in_size, out_size, str(device))
    hit = _TABLES.get(key)
    if hit is None:
        lib = _lib.load()
        ksize = lib.b200sam_resize_ksize(in_size, out_size)
        bounds = np.zeros((out_size, 2), np.int32)
        kk = np.zeros((out_size, ksize), np.int32)
        _lib.check(lib.b200sam_resize_coeffs_host(in_size, out_size, bounds.ctypes.data_as(C.c_void_p),
                                                  kk.ctypes.data_as(C.c_void_p)), "resize_coeffs_host")
        hit = (torch.from_numpy(bounds).to(device), torch.from_numpy(kk).to(device), ksize)
        _TABLES[key] = hit
    return hit


def resize_u8_cuda(image: torch.Tensor, out_h: int, out_w: int, chw: bool = False) -> torch.Tensor:
    """image: [H, W, C] uint8 CUDA tensor -> `PIL.Image.resize((out_w, out_h), BILINEAR)` of it, bit-exact, as
    [out_h, out_w, C] (chw=False) or [C, out_h, out_w] (chw=True, the encoder's input layout)."""
    from ... import _lib
    assert image.is_cuda and image.dtype == torch.uint8 and image.dim() == 3, "expected a [H, W, C] uint8 CUDA tensor"
    image = image.contiguous()
    H, W, Cc = image.shape
    lib = _lib.load()
    xb = xk = yb = yk = None
    kx = ky = 0
    if out_w != W:
        xb, xk, kx = _resize_tables(W, out_w, image.device)
    if out_h != H:
        yb, yk, ky = _resize_tables(H, out_h, image.device)
    out = torch.empty((Cc, out_h, out_w) if chw else (out_h, out_w, Cc), dtype=torch.uint8, device=image.device)
    tmp = None
    if xb is not None and (yb is not None or chw):
        tmp = torch.empty((H, out_w, Cc), dtype=torch.uint8, device=image.device)
    _lib.run(image.device, lib.b200sam_resize_u8, _lib.ptr(image), H, W, Cc, _lib.ptr(xb), _lib.ptr(xk), kx, _lib.ptr(yb),
             _lib.ptr(yk), ky, out_h, out_w, _lib.ptr(tmp), _lib.ptr(out), int(chw), what="resize_u8")
    return out


_CV_TABLES = {}


def _cv_tables(in_size: int, out_size: int, clamp: bool, device: torch.device):
    """OpenCV INTER_LINEAR tap / weight tables of one axis (library host function), cached on the device."""
    from ... import _lib
    key = (in_size, out_size, clamp, str(device))
    hit = _CV_TABLES.get(key)
    if hit is None:
        lib = _lib.load()
        idx = np.zeros((out_size, 2), np.int32)
        w = np.zeros((out_size, 2), np.int32)
        _lib.check(lib.b200sam_cvresize_coeffs_host(in_size, out_size, int(clamp), idx.ctypes.data_as(C.c_void_p),
                                                    w.ctypes.data_as(C.c_void_p)), "cvresize_coeffs_host")
        hit = (torch.from_numpy(idx).to(device), torch.from_numpy(w).to(device))
        _CV_TABLES[key] = hit
    return hit


def cv_resize_linear_cuda(image: torch.Tensor, out_h: int, out_w: int, normalize=None):
    """image: [H, W] or [N, H, W] uint8 CUDA tensor -> `cv2.resize(img, (out_w, out_h), interpolation=cv2.INTER_LINEAR)`
    of every image, bit-exact (the U-Net ingest of scripts/save_refined_segmentations.py:63).  normalize = (mean, std):
    return float32 ((u8 / 255) - mean) / std instead (save_refined_segmentations.py:64-67 fused)."""
    from ... import _lib
    assert image.is_cuda and image.dtype == torch.uint8 and image.dim() in (2, 3), "expected a [N,] H, W uint8 CUDA tensor"
    squeeze = image.dim() == 2
    img = image.contiguous().view(-1, image.shape[-2], image.shape[-1])
    n, H, W = img.shape
    lib = _lib.load()
    xi, xw = _cv_tables(W, out_w, True, img.device)
    yi, yw = _cv_tables(H, out_h, False, img.device)
    if normalize is None:
        out = torch.empty((n, out_h, out_w), dtype=torch.uint8, device=img.device)
        o8, of, mean, std = out.data_ptr(), None, 0.0, 1.0
    else:
        out = torch.empty((n, out_h, out_w), dtype=torch.float32, device=img.device)
        o8, of, mean, std = None, out.data_ptr(), float(normalize[0]), float(normalize[1])
    _lib.run(img.device, lib.b200sam_cvresize_linear_u8, img.data_ptr(), n, H, W, xi.data_ptr(), xw.data_ptr(),
             yi.data_ptr(), yw.data_ptr(), out_h, out_w, o8, of, mean, std, what="b200sam_cvresize_linear_u8")
    return out[0] if squeeze else out


def _cv_cubic_tables(in_size: int, out_size: int, device: torch.device):
    from ... import _lib
    key = ("cubic", in_size, out_size, str(device))
    hit = _CV_TABLES.get(key)
    if hit is None:
        idx = np.zeros((out_size, 4), np.int32)
        w = np.zeros((out_size, 4), np.int32)
        _lib.check(_lib.load().b200sam_cvresize_cubic_coeffs_host(in_size, out_size, idx.ctypes.data_as(C.c_void_p),
                                                                   w.ctypes.data_as(C.c_void_p)), "cvresize_cubic_coeffs_host")
        hit = (torch.from_numpy(idx).to(device), torch.from_numpy(w).to(device))
        _CV_TABLES[key] = hit
    return hit


def medsam_preprocess_cuda(gray: torch.Tensor, size: int = 1024, return_resized: bool = False):
    """The MedSAM branch of scripts/generate_img_embeddings.py:39-40,49-62 on the GPU: grey uint8 [H, W] CUDA tensor (the
    reference replicates it to RGB) -> cv2 INTER_CUBIC resize to size x size -> min-max normalise -> float32
    [1, 3, size, size] in [0, 1], the tensor the reference feeds straight into `image_encoder` (no Sam.preprocess)."""
    from ... import _lib
    assert gray.is_cuda and gray.dtype == torch.uint8 and gray.dim() == 2, "expected a [H, W] uint8 CUDA tensor"
    g = gray.contiguous()
    H, W = g.shape
    xi, xw = _cv_cubic_tables(W, size, g.device)
    yi, yw = _cv_cubic_tables(H, size, g.device)
    resized = torch.empty((size, size), dtype=torch.uint8, device=g.device)
    minmax = torch.empty((2,), dtype=torch.int32, device=g.device)
    out = torch.empty((1, 3, size, size), dtype=torch.float32, device=g.device)
    _lib.run(g.device, _lib.load().b200sam_medsam_preprocess, g.data_ptr(), H, W, xi.data_ptr(), xw.data_ptr(), yi.data_ptr(),
             yw.data_ptr(), size, resized.data_ptr(), minmax.data_ptr(), out.data_ptr(), what="b200sam_medsam_preprocess")
    return (out, resized) if return_resized else out


class ResizeLongestSide:
    def __init__(self, target_length: int) -> None:
        self.target_length = target_length

    def apply_image(self, image: np.ndarray) -> np.ndarray:
        """HxWxC uint8 -> resized uint8 (reference transforms.py:26-31: torchvision `resize(to_pil_image(...))`)."""
        from PIL import Image
        newh, neww = self.get_preprocess_shape(image.shape[0], image.shape[1], self.target_length)
        if (newh, neww) == image.shape[:2]:
            return np.array(image)
        return np.array(Image.fromarray(image).resize((neww, newh), resample=Image.BILINEAR))

    def apply_image_cuda(self, image, device=None, chw: bool = True) -> torch.Tensor:
        """`apply_image` on the GPU: HxWxC uint8 (numpy array, pinned / pageable host tensor or CUDA tensor) ->
        resized uint8 CUDA tensor ([C, h, w] for chw=True).  Bit-identical to the host path."""
        from ... import _lib
        if isinstance(image, np.ndarray):
            image = torch.from_numpy(np.ascontiguousarray(image))
        if not image.is_cuda:
            image = image.to(_lib.require_cuda(device), non_blocking=True)
        newh, neww = self.get_preprocess_shape(image.shape[0], image.shape[1], self.target_length)
        return resize_u8_cuda(image, newh, neww, chw=chw)

    def apply_coords(self, coords: np.ndarray, original_size: Tuple[int, ...]) -> np.ndarray:
        old_h, old_w = original_size
        new_h, new_w = self.get_preprocess_shape(original_size[0], original_size[1], self.target_length)
        coords = deepcopy(coords).astype(float)
        coords[..., 0] = coords[..., 0] * (new_w / old_w)
        coords[..., 1] = coords[..., 1] * (new_h / old_h)
        return coords

    def apply_boxes(self, boxes: np.ndarray, original_size: Tuple[int, ...]) -> np.ndarray:
        boxes = self.apply_coords(boxes.reshape(-1, 2, 2), original_size)
        return boxes.reshape(-1, 4)

    def apply_image_torch(self, image: torch.Tensor) -> torch.Tensor:
        target_size = self.get_preprocess_shape(image.shape[2], image.shape[3], self.target_length)
        return F.interpolate(image, target_size, mode="bilinear", align_corners=False, antialias=True)

    def apply_coords_torch(self, coords: torch.Tensor, original_size: Tuple[int, ...]) -> torch.Tensor:
        old_h, old_w = original_size
        new_h, new_w = self.get_preprocess_shape(original_size[0], original_size[1], self.target_length)
        coords = deepcopy(coords).to(torch.float)
        coords[..., 0] = coords[..., 0] * (new_w / old_w)
        coords[..., 1] = coords[..., 1] * (new_h / old_h)
        return coords

    def apply_boxes_torch(self, boxes: torch.Tensor, original_size: Tuple[int, ...]) -> torch.Tensor:
        boxes = self.apply_coords_torch(boxes.reshape(-1, 2, 2), original_size)
        return boxes.reshape(-1, 4)

    @staticmethod
    def get_preprocess_shape(oldh: int, oldw: int, long_side_length: int) -> Tuple[int, int]:
        scale = long_side_length * 1.0 / max(oldh, oldw)
        newh, neww = oldh * scale, oldw * scale
        return (int(newh + 0.5), int(neww + 0.5))
