#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
t() { eval timeout ${2:-600} python -m pytest $1 -m gpu -q -x --timeout=500 -p no:cacheprovider -s 2>&1 | tail -45 | cut -c1-400 > gpurun_out/r2_$3.log; echo "== $3: $(tail -1 gpurun_out/r2_$3.log)"; }
t tests/test_strict_gpu.py 600 strict
t "tests/test_conv_gpu.py -k 'fused_bn or apply_equals'" 300 conv_bnb
t tests/test_optim_gpu.py 300 optim
t tests/test_engine_gpu.py 600 engine
t tests/test_modeling_gpu.py 600 modeling
t tests/test_engine_headline_gpu.py 900 headline
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_a.json 2> gpurun_out/r2_bench_a.err; echo "bench rc=$?"; cut -c1-1500 gpurun_out/r2_bench_a.json
YB200_BN_FUSE=0 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-library-bar --no-convnext --no-e2e > gpurun_out/r2_bench_nofuse.json 2> gpurun_out/r2_bench_nofuse.err; echo "bench nofuse rc=$?"; cut -c1-300 gpurun_out/r2_bench_nofuse.json
bash tools/r2_profile_kernels.sh
