"""Persistence of embeddings and refined masks OFF the critical path (the writer half).

The reference writes every result synchronously from the hot loop: `predictor.features.cpu().numpy()` followed by an
h5py `create_dataset(..., compression='gzip', compression_opts=9)` (scripts/generate_img_embeddings.py:46,67-70) and
`refined_sam_masks.cpu().numpy()` + gzip-9 dataset + `estimated_dice` attribute
(scripts/save_refined_segmentations.py:75-80) - the gzip-9 compression is the wall-clock sink of both scripts.

Here `put()` only enqueues: the device tensor is copied into a pinned host buffer on a dedicated copy stream (ordered
after the producing stream by an event) and a background thread waits for that copy and writes the record.  Layouts:

* `.h5` / `.hdf5` path and `h5py` importable -> exactly the reference's layout
    embeddings: file attrs `checkpoint`, `img_encoder_img_size`; group `img_embedding/<stem>` with dataset `features`
                (1 x 256 x 64 x 64 float32, gzip-9) and attrs `original_size`, `input_size`;
    masks:      dataset `segmentation_mask/<stem>` (C x H x W bool, gzip-9) with attr `estimated_dice`; file attrs as given.
* `path=None` -> the records are collected in host memory (`writer.backend.records`);
* otherwise (h5py is not part of this image) -> a directory with `attrs.json` and, per record, raw `.npy` files
  `<stem>.<field>.npy` with the same fields (`features`, `original_size`, `input_size` / `segmentation_mask`,
  `estimated_dice`).  Plain `.npy` on purpose: the bytes go to the kernel in one `write` that releases the GIL, whereas
  `np.savez` (zip + CRC in Python) keeps the interpreter busy on the writer thread.  Records are independent files, so
  this backend is written by `threads` workers.  `open_embeddings` reads either layout back into an `EmbeddingStore`.

Staging buffers are pinned once per size class (power-of-two bytes) and shared by every writer of the process: native-size
masks have a different shape per image, and pinning a fresh 4-13 MiB buffer per record (about a millisecond per MiB)
was what the first version of this writer spent its time on.  Measured (bench.py `pipeline.with_async_writer`, ViT-H, 32
images host-in -> host-out, one B200): 125 images/s without persistence, 118 with it (first version: 116 -> 78-86); the queue
is drained when the pipeline returns.
"""
from __future__ import annotations

import json
import queue
import threading
from pathlib import Path
from typing import Dict, Optional

import numpy as np
import torch


def _have_h5py() -> bool:
    try:
        import h5py
        return hasattr(h5py, "File") and getattr(h5py, "__version__", None) is not None  # not a test stand-in module
    except Exception:
        return False


class _Backend:
    def write(self, kind: str, name: str, arrays: Dict[str, np.ndarray]) -> None:
        raise NotImplementedError

    def close(self) -> None:
        pass


class _H5Backend(_Backend):
    def __init__(self, path: Path, file_attrs: dict, gzip: int):
        import h5py
        self.f = h5py.File(path, "x")  # like the reference: refuse to overwrite
        for k, v in file_attrs.items():
            self.f.attrs[k] = v
        self.kw = dict(compression="gzip", compression_opts=gzip) if gzip else {}

    def write(self, kind, name, arrays):
        if kind == "embedding":
            g = self.f.create_group(f"img_embedding/{name}")
            g.attrs["original_size"] = arrays["original_size"]
            g.attrs["input_size"] = arrays["input_size"]
            g.create_dataset("features", data=arrays["features"], **self.kw)
        else:
            d = self.f.create_dataset("segmentation_mask/" + name, data=arrays["segmentation_mask"], **self.kw)
            d.attrs["estimated_dice"] = arrays["estimated_dice"]

    def close(self):
        self.f.close()


class _NpyDirBackend(_Backend):
    def __init__(self, path: Path, file_attrs: dict):
        path.mkdir(parents=True, exist_ok=False)  # refuse to overwrite, like h5py's 'x' mode
        self.path = path
        (path / "attrs.json").write_text(json.dumps({k: (v if isinstance(v, (str, int, float)) else str(v))
                                                     for k, v in file_attrs.items()}))

    def write(self, kind, name, arrays):
        for field, arr in arrays.items():
            np.save(self.path / f"{name}.{field}.npy", arr, allow_pickle=False)


class _MemoryBackend(_Backend):
    """path=None: the records are collected in host memory (`writer.backend.records[name][field]` numpy arrays, copied out
    of the pinned staging buffers by the writer threads) - the "results on the host" hand-over without a blocking
    `tensor.cpu()` per image on the launch thread."""

    def __init__(self):
        self.records: Dict[str, Dict[str, np.ndarray]] = {}
        self._lock = threading.Lock()

    def write(self, kind, name, arrays):
        rec = {k: np.array(v, copy=True) for k, v in arrays.items()}
        with self._lock:
            self.records[name] = rec


# pinned staging buffers (uint8, power-of-two size classes) shared by all writers of the process
_PINNED_POOL: Dict[int, list] = {}
_PINNED_LOCK = threading.Lock()
_MIN_CLASS = 1 << 12


def _size_class(nbytes: int) -> int:
    return max(_MIN_CLASS, 1 << max(0, int(nbytes) - 1).bit_length())


class AsyncResultWriter:
    """Background writer for embeddings (`kind='embedding'`) or refined masks (`kind='mask'`).

    put() cost on the producing stream: one event record; the D2H copy runs on the writer's own stream into a pinned
    buffer from a small pool (back-pressure: put() blocks only when `depth` records are still being written)."""

    def __init__(self, path, kind: str, file_attrs: Optional[dict] = None, gzip: int = 9, depth: int = 64,
                 device: Optional[torch.device] = None, threads: int = 2):
        assert kind in ("embedding", "mask")
        self.kind = kind
        if path is None:
            self.backend = _MemoryBackend()
            path = Path(".")
        else:
            path = Path(path)
        if isinstance(getattr(self, "backend", None), _MemoryBackend):
            pass
        elif path.suffix in (".h5", ".hdf5") and _have_h5py():
            self.backend: _Backend = _H5Backend(path, file_attrs or {}, gzip)
            threads = 1  # one HDF5 file, one writer
        else:
            self.backend = _NpyDirBackend(path.with_suffix("") if path.suffix in (".h5", ".hdf5") else path, file_attrs or {})
        self.device = device
        self._stream: Optional[torch.cuda.Stream] = None
        self._q: "queue.Queue" = queue.Queue(maxsize=depth)
        self._pool, self._pool_lock = _PINNED_POOL, _PINNED_LOCK
        self._error: Optional[BaseException] = None
        self.records = 0
        self._rec_lock = threading.Lock()
        self._threads = [threading.Thread(target=self._run, name=f"b200sam-writer-{i}", daemon=True)
                         for i in range(max(1, int(threads)))]
        for t in self._threads:
            t.start()

    # ------------------------------------------------------------------ producer side
    def _host_buffer(self, nbytes: int) -> torch.Tensor:
        cls = _size_class(nbytes)
        with self._pool_lock:
            free = self._pool.get(cls)
            if free:
                return free.pop()
        return torch.empty(cls, dtype=torch.uint8, pin_memory=torch.cuda.is_available())

    def _to_host_async(self, t: torch.Tensor):
        if not t.is_cuda:
            return t, None
        if self._stream is None:
            self._stream = torch.cuda.Stream(device=t.device)
        ready = torch.cuda.Event()
        ready.record(torch.cuda.current_stream(t.device))
        nbytes = t.numel() * t.element_size()
        raw = self._host_buffer(nbytes)
        host = raw[:nbytes].view(t.dtype).reshape(t.shape)
        with torch.cuda.stream(self._stream):
            self._stream.wait_event(ready)
            host.copy_(t, non_blocking=True)
            done = torch.cuda.Event()
            done.record(self._stream)
        t.record_stream(self._stream)
        return host, (done, raw)

    def put_embedding(self, name: str, features: torch.Tensor, original_size, input_size) -> None:
        """features: [1,256,64,64] float32 (what `predictor.features` holds, generate_img_embeddings.py:46)."""
        assert self.kind == "embedding"
        self._put(name, {"features": features.reshape(1, *features.shape[-3:])},
                  {"original_size": np.asarray(original_size), "input_size": np.asarray(input_size)})

    def put_masks(self, name: str, seg: torch.Tensor, est_dice: torch.Tensor) -> None:
        """seg: [C,H,W] bool, est_dice: [C] float (save_refined_segmentations.py:75-80)."""
        assert self.kind == "mask"
        self._put(name, {"segmentation_mask": seg.bool(), "estimated_dice": est_dice.float()}, {})

    def _put(self, name: str, tensors: Dict[str, torch.Tensor], extra: Dict[str, np.ndarray]) -> None:
        if self._error is not None:
            raise RuntimeError("b200sam writer thread failed") from self._error
        staged = {k: self._to_host_async(v.contiguous()) for k, v in tensors.items()}
        self._q.put((name, staged, extra))

    # ------------------------------------------------------------------ consumer side
    def _run(self) -> None:
        while True:
            item = self._q.get()
            if item is None:
                return
            name, staged, extra = item
            try:
                arrays = dict(extra)
                for k, (host, pending) in staged.items():
                    if pending is not None:
                        pending[0].synchronize()
                    arrays[k] = host.numpy()
                self.backend.write(self.kind, name, arrays)
                with self._rec_lock:
                    self.records += 1
                with self._pool_lock:
                    for _, pending in staged.values():
                        if pending is not None:
                            self._pool.setdefault(pending[1].numel(), []).append(pending[1])
            except BaseException as e:  # surfaced on the next put() / close()
                self._error = e

    def close(self) -> int:
        """Drain the queue, close the file; returns the number of records written."""
        for _ in self._threads:
            self._q.put(None)
        for t in self._threads:
            t.join()
        self.backend.close()
        if self._error is not None:
            raise RuntimeError("b200sam writer thread failed") from self._error
        return self.records

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False


def open_embeddings(path, device=None):
    """Read embeddings written by AsyncResultWriter (npy directory) or by the reference (h5, needs h5py) into an
    `EmbeddingStore` (features stay on `device` when given, else on the host until first use)."""
    from .segment_anything.sam_mask_decoder_head import EmbeddingStore
    path = Path(path)
    if path.is_dir():
        attrs = json.loads((path / "attrs.json").read_text())
        store = EmbeddingStore(str(attrs.get("checkpoint", "")), int(attrs.get("img_encoder_img_size", 1024)))
        for f in sorted(path.glob("*.features.npy")):
            stem = f.name[:-len(".features.npy")]
            feats = torch.from_numpy(np.load(f))
            store.add(stem, feats.to(device) if device is not None else feats,
                      np.load(path / f"{stem}.original_size.npy"), np.load(path / f"{stem}.input_size.npy"))
        return store
    import h5py
    with h5py.File(path, "r") as h:
        store = EmbeddingStore(str(h.attrs["checkpoint"]), int(h.attrs["img_encoder_img_size"]))
        for name, g in h["img_embedding"].items():
            feats = torch.from_numpy(g["features"][:])
            store.add(name, feats.to(device) if device is not None else feats, g.attrs["original_size"], g.attrs["input_size"])
    return store
