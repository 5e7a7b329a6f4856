#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
t() { eval timeout ${2:-600} python -m pytest $1 -m gpu -q -x --timeout=500 -p no:cacheprovider -s 2>&1 | tail -45 | cut -c1-400 > gpurun_out/r2_$3.log; echo "== $3: $(tail -1 gpurun_out/r2_$3.log)"; }
t tests/test_modeling_gpu.py 600 modeling
t tests/test_yolox_convnext_gpu.py 600 cnx_yolox
t "tests/test_detr_gpu.py -k stack" 300 detr_stack
t "tests/test_engine_gpu.py -k fused" 300 engine_fused
t "tests/test_engine_headline_gpu.py -k bs64" 600 headline64
for s in 2:3:0 2:4:0 1:4:0 4:3:0 4:4:0 1:6:0 2:6:0 2:4:8 2:4:16 4:4:16; do YB200_BN_RED=$s timeout 120 python tools/bench_bn.py 2>&1 | tail -9; done > gpurun_out/r2_bn_sweep.txt; grep -E "setting|total" gpurun_out/r2_bn_sweep.txt | paste - -
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-library-bar > gpurun_out/r2_bench_b.json 2> gpurun_out/r2_bench_b.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/r2_bench_b.json; tail -3 gpurun_out/r2_bench_b.err
bash tools/r2_profile_kernels.sh
