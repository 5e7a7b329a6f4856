#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
t() { local name=$1; shift; timeout 1200 "$@" > gpurun_out/r2c11_$name.txt 2>&1; echo "== $name rc=$?"; tail -${TAILN:-5} gpurun_out/r2c11_$name.txt | cut -c1-330; }
b() { local name=$1; shift; timeout 600 env "$@" python bench.py --steps 10 --warmup 3 --no-library-bar > gpurun_out/r2c11_bench_$name.txt 2>&1; echo "== bench $name rc=$?"; tail -1 gpurun_out/r2c11_bench_$name.txt | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print(round(d['value'], 1), 'img/s', round(d['ms_per_step'], 3), 'ms  e2e', round(d['e2e']['value'], 1), ' cnx', d['convnext'] and round(d['convnext']['images_per_s'], 1))
for c in d['kernel_classes'][:9]: print('   ', c['class'][:40], c['ms_per_step'])"; }
t tests python -m pytest tests -x -q -m gpu
b default YB200_PDL=1
b nopdl YB200_PDL=0
b phases4 YB200_DGRAD_PHASES=4
