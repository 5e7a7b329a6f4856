"""world_size-2 gloo test of the data-parallel exchange (flat gradient all-reduce, batch sharding)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from yolov7_d2_b200.dist import allreduce_gradients, shard_batch

    g = torch.Generator().manual_seed(123)
    per_rank = torch.randn(world, 1000, generator=g)      # every rank knows all "local gradients"
    flat = per_rank[rank].clone()
    allreduce_gradients(flat, average=True)
    ok_mean = torch.allclose(flat, per_rank.mean(0), atol=1e-6)
    flat = per_rank[rank].clone()
    allreduce_gradients(flat, average=False)
    ok_sum = torch.allclose(flat, per_rank.sum(0), atol=1e-6)
    lo, hi = shard_batch(8, rank, world)
    out[rank] = (ok_mean, ok_sum, lo, hi)
    dist.destroy_process_group()


def test_flat_gradient_allreduce_gloo():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    assert out[0] == (True, True, 0, 4) and out[1] == (True, True, 4, 8)


def test_shard_batch_rejects_uneven():
    from yolov7_d2_b200.dist import shard_batch

    with pytest.raises(ValueError):
        shard_batch(10, 0, 4)
