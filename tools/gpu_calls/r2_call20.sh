#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
t() { local name=$1; shift; timeout 1200 "$@" > gpurun_out/r2c20_$name.txt 2>&1; echo "== $name rc=$?"; tail -${TAILN:-4} gpurun_out/r2c20_$name.txt | cut -c1-400; }
b() { local name=$1; shift; timeout 600 env "$@" python bench.py --steps 20 --warmup 3 --no-library-bar --no-cpu-baseline --no-convnext > gpurun_out/r2c20_bench_$name.txt 2>&1; echo "== bench $name rc=$?"; tail -1 gpurun_out/r2c20_bench_$name.txt | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print(round(d['value'], 1), 'img/s', round(d['ms_per_step'], 3), 'ms  e2e', round(d['e2e']['value'], 1))
for c in d['kernel_classes'][:2]: print('   ', c['class'][:40], c['ms_per_step'])
for c in d['slowest_calls'][:3]: print('   ', c['call'][:70], c['ms'], c['floor_ms'])"; }
t tests python -m pytest tests/test_conv_gpu.py tests/test_engine_gpu.py tests/test_convnext_gpu.py tests/test_sparseinst_gpu.py -x -q -m gpu
b deep YB200_WGRAD_STAGES=6
b three YB200_WGRAD_STAGES=3
b deep2 YB200_WGRAD_STAGES=6
b three2 YB200_WGRAD_STAGES=3
