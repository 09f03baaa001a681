"""U-Net mirror (reference: custom_arcitecture/classic_u_net.py, from milesial/Pytorch-UNet).

Same constructor, sub-module tree and `state_dict` keys as the reference (`inc.double_conv.0.weight`, ...,
`up1.up.weight`, `outc.conv.bias`), so its checkpoints load with `strict=True`; `forward` runs the whole network as one
launch sequence of the CUDA library (csrc/unet.cu: im2col + tcgen05 GEMMs on 3-way bf16 split operands, fp32 statistics).
Only the reference default `bilinear=False` (transposed-convolution up-sampling) is implemented; there is no CPU path."""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch
import torch.nn as nn

from .. import _lib


class DoubleConv(nn.Module):
    """(convolution => InstanceNorm => LeakyReLU) * 2  (reference :9-27); parameter container only."""

    def __init__(self, in_channels, out_channels, mid_channels=None):
        super().__init__()
        if not mid_channels:
            mid_channels = out_channels
        self.double_conv = nn.Sequential(
            nn.Conv2d(in_channels, mid_channels, kernel_size=3, padding=1, bias=False),
            nn.InstanceNorm2d(mid_channels, affine=True),
            nn.LeakyReLU(inplace=True),
            nn.Conv2d(mid_channels, out_channels, kernel_size=3, padding=1, bias=False),
            nn.InstanceNorm2d(out_channels, affine=True),
            nn.LeakyReLU(inplace=True))

    def forward(self, x):
        raise NotImplementedError("sub-modules are parameter containers; call UNet.forward (fused CUDA pipeline)")


class Down(nn.Module):
    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.maxpool_conv = nn.Sequential(nn.MaxPool2d(2), DoubleConv(in_channels, out_channels))

    def forward(self, x):
        raise NotImplementedError("sub-modules are parameter containers; call UNet.forward (fused CUDA pipeline)")


class Up(nn.Module):
    def __init__(self, in_channels, out_channels, bilinear=True):
        super().__init__()
        if bilinear:
            raise NotImplementedError("bilinear up-sampling is not implemented (the reference default is bilinear=False)")
        self.up = nn.ConvTranspose2d(in_channels, in_channels // 2, kernel_size=2, stride=2)
        self.conv = DoubleConv(in_channels, out_channels)

    def forward(self, x1, x2):
        raise NotImplementedError("sub-modules are parameter containers; call UNet.forward (fused CUDA pipeline)")


class OutConv(nn.Module):
    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.conv = nn.Conv2d(in_channels, out_channels, kernel_size=1)

    def forward(self, x):
        raise NotImplementedError("sub-modules are parameter containers; call UNet.forward (fused CUDA pipeline)")


def _pack(sd, spec: str, lib) -> torch.Tensor:
    key, _, packing = spec.partition("|")
    t = sd[key].detach().float()
    if packing == "conv3x3_tap":  # [Cout, Cin, 3, 3] -> [Cout, Kp], column (ky*3+kx)*Cin + c
        cout, cin = t.shape[:2]
        flat = t.permute(0, 2, 3, 1).reshape(cout, 9 * cin)
        kp = lib.b200sam_unet_conv_kp(cin)
        out = torch.zeros((cout, kp), dtype=torch.float32, device=t.device)
        out[:, :9 * cin] = flat
        return out.contiguous()
    if packing == "convT":  # ConvTranspose2d [Cin, Cout, 2, 2] -> [(dy*2+dx)*Cout + co, ci]
        return t.permute(2, 3, 1, 0).reshape(-1, t.shape[0]).contiguous()
    if packing == "repeat4":
        return t.repeat(4).contiguous()
    if packing == "pad_rows8":
        t = t.reshape(t.shape[0], -1)
        rows = (t.shape[0] + 7) // 8 * 8
        out = torch.zeros((rows, t.shape[1]), dtype=torch.float32, device=t.device)
        out[:t.shape[0]] = t
        return out.contiguous()
    if packing == "pad8":
        n = (t.numel() + 7) // 8 * 8
        out = torch.zeros((n,), dtype=torch.float32, device=t.device)
        out[:t.numel()] = t
        return out.contiguous()
    if packing:
        raise ValueError(f"unknown packing {packing}")
    return t.contiguous().clone()  # never alias the live parameter


class _Engine:
    def __init__(self, model: "UNet", device: torch.device):
        lib = _lib.load()
        self.lib, self.device = lib, device
        sd = {k: v.to(device) for k, v in model.state_dict().items()}
        n = lib.b200sam_unet_weight_count()
        self.packed = [_pack(sd, lib.b200sam_unet_weight_name(i).decode(), lib) for i in range(n)]
        arr = (C.c_void_p * n)(*[t.data_ptr() for t in self.packed])
        handle = C.c_void_p()
        _lib.run(device, lib.b200sam_unet_create, model.n_channels, model.n_classes, model.n_last_channel, arr, n,
                                           C.byref(handle), what="b200sam_unet_create")
        self.handle = handle
        self._ws: Optional[torch.Tensor] = None

    def run(self, model: "UNet", x: torch.Tensor, want_logits: bool, want_probs: bool):
        B, Cn, H, W = x.shape
        assert Cn == model.n_channels, "unexpected number of input channels"
        xin = x.float().contiguous()
        need = self.lib.b200sam_unet_workspace_bytes(self.handle, B, H, W)
        if self._ws is None or self._ws.numel() < need + 256:
            self._ws = None
            self._ws = torch.empty(need + 256, dtype=torch.uint8, device=self.device)
        base = (self._ws.data_ptr() + 255) & ~255
        logits = torch.empty((B, model.n_classes, H, W), dtype=torch.float32, device=self.device) if want_logits else None
        probs = torch.empty((B, model.n_classes, H, W), dtype=torch.float32, device=self.device) if want_probs else None
        _lib.run(self.device, self.lib.b200sam_unet_forward, self.handle, xin.data_ptr(), B, H, W, _lib.ptr(logits), _lib.ptr(probs),
                                                 base, self._ws.numel() - (base - self._ws.data_ptr()), what="b200sam_unet_forward")
        return logits, probs

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                self.lib.b200sam_unet_destroy(self.handle)
        except Exception:
            pass


class UNet(nn.Module):
    def __init__(self, n_channels: int, n_classes: int, bilinear: bool = False, n_last_channel: int = 64):
        """classic U-Net (reference :81-106): same arguments; `config` mirrors the reference's @store_config_args."""
        super().__init__()
        self.config = {"n_channels": n_channels, "n_classes": n_classes, "bilinear": bilinear,
                       "n_last_channel": n_last_channel}
        self.n_channels = n_channels
        self.n_classes = n_classes
        self.bilinear = bilinear
        self.n_last_channel = n_last_channel
        self.inc = DoubleConv(n_channels, 64)
        self.down1 = Down(64, 128)
        self.down2 = Down(128, 256)
        self.down3 = Down(256, 512)
        self.down4 = Down(512, 1024)
        self.up1 = Up(1024, 512, bilinear)
        self.up2 = Up(512, 256, bilinear)
        self.up3 = Up(256, 128, bilinear)
        self.up4 = Up(128, n_last_channel, bilinear)
        self.outc = OutConv(n_last_channel, n_classes)
        self._engine: Optional[_Engine] = None
        self._engine_versions: Optional[tuple] = None
        self.register_load_state_dict_post_hook(lambda module, incompatible: module._invalidate())

    def _invalidate(self) -> None:
        self._engine = None

    def _apply(self, fn, *a, **k):
        self._invalidate()
        return super()._apply(fn, *a, **k)

    @classmethod
    def load(cls, path, device):
        """reference modelio.LoadableModel.load: {'config': ..., 'model_state': ...}"""
        ckpt = torch.load(path, map_location="cpu")
        model = cls(**ckpt["config"])
        model.load_state_dict(ckpt["model_state"], strict=False)
        return model.to(device)

    def _eng(self) -> _Engine:
        dev = self.outc.conv.weight.device
        if dev.type != "cuda":
            raise _lib.B200SamError("b200sam has no CPU path: move the U-Net to a CUDA device")
        versions = tuple(p._version for p in self.parameters())  # in-place weight edits bump the version: re-pack
        if self._engine is None or self._engine.device != dev or versions != self._engine_versions:
            self._engine = None
            self._engine = _Engine(self, dev)
            self._engine_versions = versions
        return self._engine

    @torch.no_grad()
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """x: [B, n_channels, H, W] normalised image -> logits [B, n_classes, H, W] (reference :108-119)."""
        return self._eng().run(self, x.to(self.outc.conv.weight.device), True, False)[0]

    @torch.no_grad()
    def predict_proba(self, x: torch.Tensor) -> torch.Tensor:
        """sigmoid(forward(x)) in the same launch sequence (save_refined_segmentations.py:68-69)."""
        return self._eng().run(self, x.to(self.outc.conv.weight.device), False, True)[1]
