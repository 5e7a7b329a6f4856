#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
t() { local name=$1; shift; timeout 1200 "$@" > gpurun_out/r2c19_$name.txt 2>&1; echo "== $name rc=$?"; tail -${TAILN:-4} gpurun_out/r2c19_$name.txt | cut -c1-400; }
b() { local name=$1; shift; timeout 600 env "$@" python bench.py --steps 20 --warmup 3 --no-library-bar --no-cpu-baseline --no-convnext > gpurun_out/r2c19_bench_$name.txt 2>&1; echo "== bench $name rc=$?"; tail -1 gpurun_out/r2c19_bench_$name.txt | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print(round(d['value'], 1), 'img/s', round(d['ms_per_step'], 3), 'ms  e2e', round(d['e2e']['value'], 1))
for c in d['slowest_calls'][:5]: print('   ', c['call'][:70], c['ms'], c['floor_ms'])"; }
t tests python -m pytest tests/test_engine_gpu.py tests/test_conv_gpu.py tests/test_engine_headline_gpu.py -x -q -m gpu
b sparse YB200_STEM_SPARSE=1
b dense YB200_STEM_SPARSE=0
b sparse2 YB200_STEM_SPARSE=1
b dense2 YB200_STEM_SPARSE=0
