#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
b() { local name=$1; shift; timeout 300 env "$@" python bench.py --steps 20 --warmup 3 --no-library-bar --no-cpu-baseline --no-convnext --no-e2e 2>/dev/null | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('$name', round(d['value'], 1), 'img/s', round(d['ms_per_step'], 3), 'ms', [(c['class'][:12], c['ms_per_step']) for c in d['kernel_classes'][:1]])"; }
b red230 YB200_BN_RED=2:3:0
b red220 YB200_BN_RED=2:2:0
b red430 YB200_BN_RED=4:3:0
b red420 YB200_BN_RED=4:2:0
b red130 YB200_BN_RED=1:3:0
b red230b YB200_BN_RED=2:3:0
