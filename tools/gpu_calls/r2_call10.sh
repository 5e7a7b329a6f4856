#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
t() { local name=$1; shift; timeout 900 "$@" > gpurun_out/r2c10_$name.txt 2>&1; echo "== $name rc=$?"; tail -${TAILN:-6} gpurun_out/r2c10_$name.txt | cut -c1-300; }
t tests python -m pytest tests/test_conv_gpu.py tests/test_elementwise_gpu.py tests/test_engine_gpu.py tests/test_simota_gpu.py tests/test_optim_gpu.py -x -q -m gpu
export YB200_DUMP_CALLS=gpurun_out/r2c10_calls.jsonl
TAILN=2 t bench python bench.py --steps 10 --warmup 3 --no-library-bar
unset YB200_DUMP_CALLS
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2c10_launches.csv python tools/profile_step.py --batch 64 > gpurun_out/r2c10_ncu.txt 2>&1
python tools/summarize_launches.py gpurun_out/r2c10_launches.csv 45 gpurun_out/trace.json > gpurun_out/r2c10_launches.md 2>&1; head -60 gpurun_out/r2c10_launches.md | cut -c1-200
