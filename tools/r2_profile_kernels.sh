#!/bin/bash
# ncu --set full captures of the kernels VERDICT r1 named, taken from one real bs=64 training step (tools/profile_step.py)
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on --profile-from-start off --kernel-name-base demangled"
run() { # name regex skip count
  timeout 300 $NCU -k "regex:$2" -s $3 -c $4 -f -o gpurun_out/$1 python tools/profile_step.py --batch 64 > gpurun_out/$1.log 2>&1; echo "$1 rc=$? $(grep -c 'PROF.*%' gpurun_out/$1.log) lines"; }
I='.int.'
run r2_stem_conv "conv_gemm_persistent_kernel<${I}32, ${I}16, ${I}0>" 0 1
run r2_conv_64_64 "conv_gemm_persistent_kernel<${I}64, ${I}64, ${I}0>" 0 2
run r2_conv_128_64 "conv_gemm_persistent_kernel<${I}128, ${I}64, ${I}0>" 0 2
run r2_stem_wgrad "wgrad_gemm_kernel<${I}64>" 6 1
run r2_bn_bwd_reduce "bn_silu_bwd_reduce_kernel" 73 1
run r2_bn_bwd_apply "bn_silu_bwd_apply_kernel" 73 1
# keep gpurun_out small (64 MiB cap): export the metric tables as CSV and drop the reports
for f in gpurun_out/r2_*.ncu-rep; do
  ncu -i $f --page raw --csv > ${f%.ncu-rep}.raw.csv 2>/dev/null
  ncu -i $f --page details --csv > ${f%.ncu-rep}.details.csv 2>/dev/null
  rm -f $f
done
ls -la gpurun_out/ | head -40
