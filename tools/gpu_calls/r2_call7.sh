#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 300 python tools/diag_cnx.py > gpurun_out/r2_diag_cnx.txt 2>&1; tail -70 gpurun_out/r2_diag_cnx.txt | cut -c1-200
timeout 300 python tools/diag_width.py 0.75 > gpurun_out/r2_diag_width075.txt 2>&1; tail -6 gpurun_out/r2_diag_width075.txt
