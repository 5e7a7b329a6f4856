#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_detr_dropout_gpu.py tests/test_attention_gpu.py tests/test_attention_bwd_gpu.py tests/test_detr_gpu.py -q -m gpu 2>&1 | tail -3 | cut -c1-250
timeout 600 python tools/bench_attention.py > gpurun_out/r2c28_attention.txt 2>&1; tail -1 gpurun_out/r2c28_attention.txt | python -c "
import sys, json
d = json.loads(sys.stdin.read())
for k, v in d.items(): print(k, {a: round(b, 1) for a, b in v.items()})"
