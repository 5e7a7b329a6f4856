#!/bin/bash
# round-2 GPU call 1: new parity tests first (cheap), then the baseline ncu captures
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
t() { timeout ${2:-600} python -m pytest $1 -m gpu -q -x --timeout=500 -p no:cacheprovider -s 2>&1 | tail -40 | cut -c1-300 > gpurun_out/r2_$3.log; echo "== $3: $(tail -1 gpurun_out/r2_$3.log)"; }
t tests/test_strict_gpu.py 600 strict
t tests/test_nms_gpu.py 300 nms
t tests/test_optim_gpu.py 300 optim
t tests/test_engine_headline_gpu.py 900 headline
t "tests/test_modeling_gpu.py tests/test_engine_gpu.py" 600 engine
YB200_DETR_TRAINING=1 timeout 600 python -m pytest tests/test_detr_gpu.py -m gpu -q --timeout=300 -s 2>&1 | tail -40 | cut -c1-300 > gpurun_out/r2_detr_training.log; echo "== detr: $(tail -1 gpurun_out/r2_detr_training.log)"
bash tools/r2_profile_baseline.sh
