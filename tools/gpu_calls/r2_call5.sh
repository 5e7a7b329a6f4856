#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
t() { eval timeout ${2:-600} python -m pytest $1 -m gpu -q -x --timeout=500 -p no:cacheprovider -s 2>&1 | tail -90 | cut -c1-700 > gpurun_out/r2_$3.log; echo "== $3: $(tail -1 gpurun_out/r2_$3.log)"; }
t "tests/test_conv_gpu.py -k 'fwd or dgrad'" 600 conv
t "tests/test_engine_gpu.py tests/test_elementwise_gpu.py" 600 engine
t tests/test_yolox_convnext_gpu.py 600 cnx_yolox
t "tests/test_sparseinst_gpu.py tests/test_engine_headline_gpu.py" 900 misc
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-library-bar > gpurun_out/r2_bench_d.json 2> gpurun_out/r2_bench_d.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/r2_bench_d.json; tail -3 gpurun_out/r2_bench_d.err
YB200_CONV_STAGED=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-library-bar --no-convnext --no-e2e > gpurun_out/r2_bench_d_classic.json 2> gpurun_out/r2_bench_d_classic.err; echo "bench classic rc=$?"; cut -c1-300 gpurun_out/r2_bench_d_classic.json
