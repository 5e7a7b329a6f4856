"""Two-rank correctness of the data-parallel exchange: the bucketed, backward-overlapped all-reduce of yolov7_d2_b200.dist.GradientBuckets
leaves in every rank's flat gradient buffer the sum of the two ranks' single-GPU gradients.  With two or more CUDA devices: one rank per GPU over
NCCL (what `bench.py --gpus N` and the driver's scaling run use).  On a one-GPU box the two ranks share cuda:0 and exchange through gloo (NCCL
refuses two ranks on one device): the same engine kernels, range-by-range backward, communication stream and event ordering, without NVLink."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out, backend):
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dev = torch.device("cuda", rank if backend == "nccl" else 0)
    torch.cuda.set_device(dev)
    if backend == "nccl":
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    else:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    from yolov7_d2_b200 import synth
    from yolov7_d2_b200.dist import GradientBuckets
    from yolov7_d2_b200.engine import YoloxEngine

    eng = YoloxEngine(4, 128, 128, device=dev)
    eng.init_weights(0)  # same seed on both ranks: identical replicas
    shards = [synth.synthetic_batch(4, 128, seed=300 + r, max_gt=5) for r in range(world)]

    def load(r):
        eng.images_u8.copy_(shards[r][0].to(dev))
        eng.labels.copy_(shards[r][1].to(dev))

    gb = GradientBuckets(eng)
    covered = sorted(r for part in gb.slices for r in part)
    layout_ok = all(any(lo <= off and off + n <= hi for lo, hi in covered) for _, off, n in eng.param_layout)
    load(rank)
    eng.pack_weights()
    eng.preprocess()
    eng.forward_features(True)
    eng.assign_and_loss(True)
    gb.step_backward()
    gb.wait()
    torch.cuda.synchronize()
    reduced = eng.flat_grad.clone()
    ref = torch.zeros_like(reduced)
    for r in range(world):  # the same two single-GPU steps, whole-plan backward, no communication
        load(r)
        eng.train_step()
        torch.cuda.synchronize()
        ref += eng.flat_grad
    cos = float(torch.dot(reduced.double(), ref.double()) / (reduced.double().norm() * ref.double().norm()))
    rel = float((reduced - ref).abs().max() / ref.abs().max())
    out[rank] = (layout_ok, cos, rel)
    dist.destroy_process_group()


def test_bucketed_allreduce_equals_sum_of_single_gpu_gradients(cuda):
    import torch.multiprocessing as mp

    world = 2
    backend = "nccl" if torch.cuda.device_count() >= world else "gloo"
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out, backend), nprocs=world, join=True)
    for r in range(world):
        layout_ok, cos, rel = out[r]
        assert layout_ok, "a parameter lies outside every gradient bucket"
        # identical arithmetic up to the 16-bit storage noise of a repeated step (tests/test_engine_gpu.py::test_second_step_is_reproducible)
        assert cos >= 0.99999 and rel <= 4e-3, (r, cos, rel)
