#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_nms_gpu.py tests/test_modeling_gpu.py -x -q -m gpu 2>&1 | tail -3
timeout 300 python tools/bench_nms.py 2>&1 | tail -2
