"""world_size-2 gloo tests (CPU) of the N>1 host logic: batch sharding, barrier, max-over-ranks
timing and logits reassembly.  The forward has no data-path collective."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from kosmosx import dist as kd


def test_shard_range_partitions_exactly():
    for gb in (1, 7, 8, 64, 65):
        for world in (1, 2, 3, 8):
            spans = [kd.shard_range(gb, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == gb
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        kd.shard_range(8, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    r, _, w = kd.init_from_env("gloo")
    assert (r, w) == (rank, world)
    gb = 5
    text = torch.arange(gb * 3).view(gb, 3)
    images = torch.arange(gb * 2, dtype=torch.float32).view(gb, 2)
    t, im = kd.shard_batch(text, images, r, w)
    # stand-in for the per-rank forward: a function of the local shard only
    local = (t.float().sum(1, keepdim=True) + im.sum(1, keepdim=True)).view(-1, 1, 1)
    kd.barrier()
    full = kd.gather_logits(local, gb)
    want = (text.float().sum(1, keepdim=True) + images.sum(1, keepdim=True)).view(-1, 1, 1)
    ok = torch.equal(full, want)
    mx = kd.max_over_ranks(10.0 + rank)
    sm = kd.sum_over_ranks(t.shape[0])
    out[rank] = (ok, mx, sm)
    dist.destroy_process_group()


def test_two_rank_gloo_sharding_and_reductions():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    assert dict(out) == {0: (True, 11.0, 5.0), 1: (True, 11.0, 5.0)}
