#!/bin/bash
# ncu --set full captures of the kernels VERDICT r1 named, taken from one real bs=64 training step (tools/profile_step.py)
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on --profile-from-start off --kernel-name-base demangled"
run() { # name regex skip count
  timeout 300 $NCU -k "regex:$2" -s $3 -c $4 -f -o gpurun_out/$1 python tools/profile_step.py --batch 64 > gpurun_out/$1.log 2>&1; echo "$1 rc=$? $(grep -c 'PROF.*Profiling' gpurun_out/$1.log) kernels"; }
run r2_stem_conv 'conv_gemm_persistent_kernel<32, 16, 0>' 0 1
run r2_conv_64_64 'conv_gemm_persistent_kernel<64, 64, 0>' 0 2
run r2_conv_128_64 'conv_gemm_persistent_kernel<128, 64, 0>' 0 2
run r2_dgrad_bnb 'conv_gemm_persistent_kernel<(32|64|128), (16|32|64), 2>' 30 3
run r2_stem_wgrad 'wgrad_gemm_kernel<64>' 6 1
run r2_bn_bwd_apply 'bn_silu_bwd_apply_kernel<true>' 60 1
run r2_wgrad_reduce 'wgrad_reduce_kernel' 60 2
ls -la gpurun_out/r2_*.ncu-rep
