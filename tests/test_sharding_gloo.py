"""World-size-2 CPU (gloo) test of the N>1 path: round-robin image sharding + the final gather."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from samcarriestheburden_b200 import sharding


def test_shard_indices_partition():
    for n in (0, 1, 7, 500):
        for w in (1, 2, 4, 8):
            parts = [sharding.shard_indices(n, r, w) for r in range(w)]
            assert sorted(i for p in parts for i in p) == list(range(n))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
    with pytest.raises(ValueError):
        sharding.shard_indices(5, 3, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_items, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        mine = sharding.shard_indices(n_items)
        # stand-in for the per-image result (an "embedding" whose value encodes the image index)
        local = torch.stack([torch.full((3, 2), float(i)) for i in mine]) if mine else torch.zeros((0, 3, 2))
        full = sharding.gather_sharded(local, n_items)
        expect = torch.arange(n_items, dtype=torch.float32)[:, None, None].expand(n_items, 3, 2)
        ret[rank] = bool(torch.equal(full, expect))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_items", [1, 5, 8])
def test_gather_sharded_world2(n_items):
    port = _free_port()
    ctx = mp.get_context("spawn")
    with ctx.Manager() as mgr:
        ret = mgr.dict()
        procs = [ctx.Process(target=_worker, args=(r, 2, port, n_items, ret)) for r in range(2)]
        for p in procs:
            p.start()
        for p in procs:
            p.join(timeout=120)
            assert p.exitcode == 0
        assert dict(ret) == {0: True, 1: True}
