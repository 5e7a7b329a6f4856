#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
b() { local name=$1; shift; timeout 600 env "$@" python bench.py --steps 20 --warmup 3 --no-library-bar --no-cpu-baseline --no-convnext > gpurun_out/r2c16_bench_$name.txt 2>&1; echo "== bench $name rc=$?"; tail -1 gpurun_out/r2c16_bench_$name.txt | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print(round(d['value'], 1), 'img/s', round(d['ms_per_step'], 3), 'ms  e2e', round(d['e2e']['value'], 1), 'prefetch host ms', d['e2e'].get('prefetch_host_ms_per_step'))"; }
nproc; lscpu | grep -E "Model name|Socket|NUMA node\(s\)" | head -4
b stack YB200_GATHER=stack
b threads YB200_GATHER=threads
b direct YB200_GATHER=direct
b stack2 YB200_GATHER=stack
b threads2 YB200_GATHER=threads
b direct2 YB200_GATHER=direct
