#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
t() { local name=$1; shift; timeout 1200 "$@" > gpurun_out/r2c12_$name.txt 2>&1; echo "== $name rc=$?"; tail -${TAILN:-5} gpurun_out/r2c12_$name.txt | cut -c1-330; }
b() { local name=$1; shift; timeout 600 env "$@" python bench.py --steps 10 --warmup 3 --no-library-bar > gpurun_out/r2c12_bench_$name.txt 2>&1; echo "== bench $name rc=$?"; tail -1 gpurun_out/r2c12_bench_$name.txt | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print(round(d['value'], 1), 'img/s', round(d['ms_per_step'], 3), 'ms  e2e', round(d['e2e']['value'], 1), ' cnx', d['convnext'] and round(d['convnext']['images_per_s'], 1))
for c in d['slowest_calls'][:8]: print('   ', c['call'][:70], c['ms'], c['floor_ms'])"; }
t tests python -m pytest tests -x -q -m gpu
b default YB200_PDL=1
b densestem YB200_STEM_SPARSE=0
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2c12_launches.csv python tools/profile_step.py --batch 64 > gpurun_out/r2c12_ncu.txt 2>&1
python tools/summarize_launches.py gpurun_out/r2c12_launches.csv 40 gpurun_out/trace.json > gpurun_out/r2c12_launches.md 2>&1
python tools/class_traffic.py gpurun_out/r2c12_launches.csv gpurun_out/r2c12_traffic.json > gpurun_out/r2c12_traffic.md 2>&1; cat gpurun_out/r2c12_traffic.md | cut -c1-200
gzip -f gpurun_out/r2c12_launches.csv
