#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
t() { local name=$1; shift; timeout 1200 "$@" > gpurun_out/r2c14_$name.txt 2>&1; echo "== $name rc=$?"; tail -${TAILN:-4} gpurun_out/r2c14_$name.txt | cut -c1-400; }
b() { local name=$1; shift; timeout 600 env "$@" python bench.py --steps 10 --warmup 3 --no-library-bar > gpurun_out/r2c14_bench_$name.txt 2>&1; echo "== bench $name rc=$?"; tail -1 gpurun_out/r2c14_bench_$name.txt | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print(round(d['value'], 1), 'img/s', round(d['ms_per_step'], 3), 'ms  e2e', round(d['e2e']['value'], 1), ' cnx', d['convnext'] and round(d['convnext']['images_per_s'], 1))"; }
t smoke python -c "import __graft_entry__ as g; g.smoke()"
t dist1 python -m pytest tests/test_dist_gpu.py tests/test_sparseinst_gpu.py -x -q -m gpu
b pdl_wgrad_on YB200_PDL_WGRAD=1
b pdl_wgrad_off YB200_PDL_WGRAD=0
b pdl_wgrad_on2 YB200_PDL_WGRAD=1
b pdl_wgrad_off2 YB200_PDL_WGRAD=0
