#!/bin/bash
# two-GPU call: data-parallel correctness + the N=2 bench line (launched exactly as the driver does)
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
t() { local name=$1; shift; timeout 900 "$@" > gpurun_out/r2c9_$name.txt 2>&1; echo "== $name rc=$?"; tail -${TAILN:-6} gpurun_out/r2c9_$name.txt | cut -c1-400; }
nvidia-smi -L
t dist python -m pytest tests/test_dist_gpu.py tests/test_engine_gpu.py -x -q -m gpu
TAILN=3 t bench2 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3
TAILN=3 t bench1 python bench.py --steps 10 --warmup 3 --no-library-bar
TAILN=3 t ref2 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 1
