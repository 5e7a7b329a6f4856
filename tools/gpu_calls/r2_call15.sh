#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
t() { local name=$1; shift; timeout 900 "$@" > gpurun_out/r2c15_$name.txt 2>&1; echo "== $name rc=$?"; tail -${TAILN:-4} gpurun_out/r2c15_$name.txt | cut -c1-400; }
b() { local name=$1; shift; timeout 600 env "$@" python bench.py --steps 10 --warmup 3 --no-library-bar --no-cpu-baseline --no-convnext > gpurun_out/r2c15_bench_$name.txt 2>&1; echo "== bench $name rc=$?"; tail -1 gpurun_out/r2c15_bench_$name.txt | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print(round(d['value'], 1), 'img/s', round(d['ms_per_step'], 3), 'ms  e2e', round(d['e2e']['value'], 1))
for c in d['kernel_classes'][:12]: print('   ', c['class'][:40], c['ms_per_step'])"; }
t tests python -m pytest tests/test_conv_gpu.py tests/test_simota_gpu.py tests/test_modeling_gpu.py tests/test_engine_gpu.py tests/test_dist_gpu.py -x -q -m gpu
b a YB200_PDL=1
b b YB200_PDL=1
timeout 600 python bench.py --steps 10 --warmup 3 --no-library-bar --no-cpu-baseline --no-convnext --no-prefetch | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('noprefetch e2e', round(d['e2e']['value'],1))"
