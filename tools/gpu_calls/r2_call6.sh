#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
t() { eval timeout ${2:-600} python -m pytest $1 -m gpu -q -x --timeout=500 -p no:cacheprovider -s 2>&1 | tail -90 | cut -c1-700 > gpurun_out/r2_$3.log; echo "== $3: $(tail -1 gpurun_out/r2_$3.log)"; }
t "tests/test_engine_gpu.py tests/test_elementwise_gpu.py tests/test_simota_gpu.py" 900 core
t "tests/test_conv_gpu.py -k 'fwd or dgrad'" 600 conv
timeout 300 python tools/diag_width.py 0.75 > gpurun_out/r2_diag_width075.txt 2>&1; tail -16 gpurun_out/r2_diag_width075.txt
t tests/test_yolox_convnext_gpu.py 600 cnx_yolox
grep -n "largest cosine\|logits vs" gpurun_out/r2_cnx_yolox.log | cut -c1-700
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-library-bar > gpurun_out/r2_bench_e.json 2> gpurun_out/r2_bench_e.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/r2_bench_e.json; tail -3 gpurun_out/r2_bench_e.err
YB200_STEM_GROUP4=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-library-bar --no-convnext --no-e2e > gpurun_out/r2_bench_e_nogroup.json 2> gpurun_out/r2_bench_e_nogroup.err; echo "bench nogroup rc=$?"; cut -c1-300 gpurun_out/r2_bench_e_nogroup.json
t "tests/test_engine_headline_gpu.py tests/test_modeling_gpu.py tests/test_sparseinst_gpu.py tests/test_detr_gpu.py tests/test_strict_gpu.py" 1200 rest
