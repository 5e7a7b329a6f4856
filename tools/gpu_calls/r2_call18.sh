#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
t() { local name=$1; shift; timeout 1200 "$@" > gpurun_out/r2c18_$name.txt 2>&1; echo "== $name rc=$?"; tail -${TAILN:-5} gpurun_out/r2c18_$name.txt | cut -c1-400; }
t tests python -m pytest tests -q -m gpu
