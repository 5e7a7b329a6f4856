#!/bin/bash
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 300 python bench.py --steps 20 --warmup 3 --no-library-bar --no-cpu-baseline --no-convnext 2>/dev/null | tail -1 > gpurun_out/r2c31_bench.json; python -c "
import json; d = json.load(open('gpurun_out/r2c31_bench.json')); print(round(d['value'], 1), 'img/s', round(d['ms_per_step'], 3), 'ms  e2e', round(d['e2e']['value'], 1))"
