#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
t() { eval timeout ${2:-600} python -m pytest $1 -m gpu -q -x --timeout=500 -p no:cacheprovider -s 2>&1 | tail -60 | cut -c1-500 > gpurun_out/r2_$3.log; echo "== $3: $(tail -1 gpurun_out/r2_$3.log)"; }
t tests/test_yolox_convnext_gpu.py 600 cnx_yolox
t "tests/test_engine_headline_gpu.py -k bs64" 600 headline64
t "tests/test_simota_gpu.py tests/test_elementwise_gpu.py tests/test_engine_gpu.py" 900 core
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-library-bar > gpurun_out/r2_bench_c.json 2> gpurun_out/r2_bench_c.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/r2_bench_c.json; tail -3 gpurun_out/r2_bench_c.err
bash tools/r2_profile_kernels.sh
