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


def test_bucket_slices_follow_the_plan_order():
    """head / neck / backbone weight slices are contiguous, BatchNorm parameters and biases ride with the last bucket, nothing is left out"""
    from yolov7_d2_b200.dist import bucket_slices

    layout = [("backbone.stem.conv.conv.weight", 0, 100), ("backbone.dark2.0.conv.weight", 100, 200), ("neck.a.conv.weight", 300, 50),
              ("head.stems.0.conv.weight", 352, 40), ("head.cls_preds.0.weight", 392, 10), ("head.reg_preds.0.weight", 404, 4),
              ("head.obj_preds.0.weight", 408, 1), ("backbone.stem.conv.bn.weight", 420, 4), ("backbone.stem.conv.bn.bias", 424, 4),
              ("head.cls_preds.0.bias", 428, 3)]
    parts = bucket_slices(layout, 432)
    assert parts == [[(352, 409)], [(300, 350)], [(0, 300), (420, 432)]]
    flat = sorted(r for p in parts for r in p)
    assert all(any(lo <= off and off + n <= hi for lo, hi in flat) for _, off, n in layout)


def _bucket_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import types

    from yolov7_d2_b200.dist import GradientBuckets

    layout = [("backbone.a.conv.weight", 0, 64), ("neck.b.conv.weight", 64, 32), ("head.c.conv.weight", 96, 16), ("head.c.bn.weight", 112, 8)]
    g = torch.Generator().manual_seed(7)
    per_rank = torch.randn(world, 120, generator=g)
    calls = []
    eng = types.SimpleNamespace(param_layout=layout, flat_grad=per_rank[rank].clone(), dev=torch.device("cpu"),
                                ranges={"head": (2, 3), "neck": (1, 2), "backbone": (0, 1)},
                                backward=lambda acc, rng, fresh: calls.append((rng, fresh)))
    gb = GradientBuckets(eng)
    gb.step_backward()
    gb.wait()
    out[rank] = (torch.allclose(eng.flat_grad, per_rank.sum(0), atol=1e-6), calls)
    dist.destroy_process_group()


def test_gradient_buckets_gloo():
    """host logic of the bucketed exchange on two gloo ranks: every slice is reduced exactly once, ranges run in backward order"""
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_bucket_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    for r in range(world):
        ok, calls = out[r]
        assert ok and calls == [((2, 3), True), ((1, 2), False), ((0, 1), False)]
