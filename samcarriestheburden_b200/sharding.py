"""Image-level data parallelism for the two pipeline drivers.

Every image is independent (encoder, prompts, decode, upscale are all per image), so the path shards with NO
data-path collective: rank r of W processes images i with i % W == r (round-robin keeps native-size variation
balanced).  The only communication is one final gather of the per-rank results (NCCL over NVLink on GPUs,
gloo in the CPU tests)."""
from __future__ import annotations

from typing import List, Sequence

import torch
import torch.distributed as dist


def world() -> int:
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def rank() -> int:
    return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0


def shard_indices(n_items: int, rank_: int | None = None, world_: int | None = None) -> List[int]:
    """Indices owned by `rank_`: i % world == rank (static round-robin)."""
    w = world() if world_ is None else world_
    r = rank() if rank_ is None else rank_
    if not (0 <= r < w):
        raise ValueError(f"rank {r} outside world of size {w}")
    return list(range(r, n_items, w))


def gather_sharded(local: torch.Tensor, n_items: int) -> torch.Tensor:
    """All-gather per-rank result rows (rank r holds rows for shard_indices(n_items, r)) back into dataset order.
    local: [len(shard), ...] on the rank's device.  Returns [n_items, ...] on every rank."""
    w, r = world(), rank()
    if w == 1:
        return local
    per = (n_items + w - 1) // w  # ranks hold per or per-1 rows: pad to a common size for one fused all_gather
    pad = torch.zeros((per,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    out = torch.empty((w, per) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out.view(w * per, *local.shape[1:]), pad)
    # out[r, j] is item j * w + r  ->  interleave back to dataset order
    full = out.transpose(0, 1).reshape(per * w, *local.shape[1:])
    return full[:n_items].contiguous()
