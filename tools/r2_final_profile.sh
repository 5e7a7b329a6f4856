#!/bin/bash
# round-2 evidence run (one B200): launch list + DRAM traffic per kernel class, the default bench line, the reference arm, the per-call dump
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/final_launches.csv python tools/profile_step.py --batch 64 > gpurun_out/final_ncu.txt 2>&1
python tools/summarize_launches.py gpurun_out/final_launches.csv 40 gpurun_out/trace.json > gpurun_out/final_launches.md 2>&1
python tools/class_traffic.py gpurun_out/final_launches.csv gpurun_out/final_traffic.json > gpurun_out/final_traffic.md 2>&1; head -16 gpurun_out/final_traffic.md | cut -c1-160
gzip -f gpurun_out/final_launches.csv
export YB200_DUMP_CALLS=gpurun_out/final_calls.jsonl
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/final_bench.txt 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/final_bench.txt | cut -c1-400
unset YB200_DUMP_CALLS
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/final_bench_reference.txt 2>&1; echo "reference rc=$?"; tail -1 gpurun_out/final_bench_reference.txt | cut -c1-400
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
