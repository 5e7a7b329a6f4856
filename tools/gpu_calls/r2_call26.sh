#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_detr_dropout_gpu.py tests/test_attention_gpu.py tests/test_attention_bwd_gpu.py tests/test_detr_gpu.py -q -m gpu 2>&1 | tail -25 | cut -c1-250
