#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
t() { local name=$1; shift; timeout 1200 "$@" > gpurun_out/r2c17_$name.txt 2>&1; echo "== $name rc=$?"; tail -${TAILN:-5} gpurun_out/r2c17_$name.txt | cut -c1-400; }
t tests python -m pytest tests -x -q -m gpu
timeout 600 python bench.py --steps 20 --warmup 3 --no-library-bar --no-cpu-baseline --no-convnext > gpurun_out/r2c17_bench.txt 2>&1; tail -1 gpurun_out/r2c17_bench.txt | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print(round(d['value'], 1), 'img/s', round(d['ms_per_step'], 3), 'ms  e2e', round(d['e2e']['value'], 1), d['e2e'].get('prefetch_host_ms_per_step'))"
