#!/bin/bash
# two-GPU evidence with the final build: data-parallel correctness (NCCL) + N=2 bench lines, launched exactly as the driver does
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
t() { local name=$1; shift; timeout 900 "$@" > gpurun_out/r2c23_$name.txt 2>&1; echo "== $name rc=$?"; tail -${TAILN:-3} gpurun_out/r2c23_$name.txt | cut -c1-500; }
t dist python -m pytest tests/test_dist_gpu.py -x -q -m gpu
TAILN=1 t bench2 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 20 --warmup 3
TAILN=1 t ref2 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 bench.py --impl reference --gpus 2 --steps 1 --warmup 1
