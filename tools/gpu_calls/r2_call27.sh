#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python tools/bench_attention.py > gpurun_out/r2c27_attention.txt 2>&1; tail -1 gpurun_out/r2c27_attention.txt | cut -c1-300
timeout 600 python bench.py --steps 20 --warmup 3 --no-library-bar --no-cpu-baseline --no-convnext 2>/dev/null | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print(round(d['value'], 1), 'img/s', round(d['ms_per_step'], 3), 'ms  e2e', round(d['e2e']['value'], 1))"
