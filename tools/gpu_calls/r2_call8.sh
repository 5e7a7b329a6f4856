#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
t() { local name=$1; shift; eval "timeout 900 $*" > gpurun_out/r2c8_$name.txt 2>&1; echo "== $name rc=$?"; tail -${TAILN:-6} gpurun_out/r2c8_$name.txt | cut -c1-260; }
t cnx python -m pytest tests/test_yolox_convnext_gpu.py -x -q -m gpu -s
t modeling python -m pytest tests/test_modeling_gpu.py -x -q -m gpu
t conv_engine python -m pytest tests/test_conv_gpu.py tests/test_engine_gpu.py tests/test_optim_gpu.py -x -q -m gpu
TAILN=3 t bench python bench.py --steps 10 --warmup 3
TAILN=3 t bench_nographs YB200_API_GRAPHS=0 python bench.py --steps 10 --warmup 3 --no-library-bar
