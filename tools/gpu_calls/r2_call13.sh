#!/bin/bash
# two-GPU call: data-parallel correctness + N=2 bench lines (YOLOX-s and YOLOX-ConvNeXt-T), launched exactly as the driver does; DETR timings on GPU 0
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
t() { local name=$1; shift; timeout 900 "$@" > gpurun_out/r2c13_$name.txt 2>&1; echo "== $name rc=$?"; tail -${TAILN:-4} gpurun_out/r2c13_$name.txt | cut -c1-600; }
t dist python -m pytest tests/test_dist_gpu.py -x -q -m gpu
TAILN=1 t bench2 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 10 --warmup 3
TAILN=1 t bench2_cnx python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 2 --steps 5 --warmup 3 --workload yolox_convnext --no-library-bar
TAILN=1 t attention python tools/bench_attention.py
