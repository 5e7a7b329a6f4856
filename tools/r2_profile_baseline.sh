#!/bin/bash
# round-2 baseline evidence: ncu --set full captures of the memory-bound kernels VERDICT r1 named, taken from one real bs=64 step
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on --profile-from-start off"
run() { # name regex skip count
  timeout 300 $NCU -k "regex:$2" -s $3 -c $4 -f -o gpurun_out/$1 python tools/profile_step.py --batch 64 > gpurun_out/$1.log 2>&1; echo "$1 rc=$?"; }
run r2base_stem_conv 'conv_gemm_persistent_kernel<32, 16' 0 1
run r2base_bn_apply 'bn_apply_silu_kernel' 0 1
run r2base_simota_loss 'simota_prep_kernel|simota_match_kernel|yolox_loss_kernel' 0 3
run r2base_stem_wgrad 'wgrad_gemm_kernel<64>' 6 1
run r2base_bn_bwd_reduce 'bn_silu_bwd_reduce_kernel' 73 1
run r2base_bn_bwd_apply 'bn_silu_bwd_apply_kernel<1>' 71 1
run r2base_conv1x1 'conv_gemm_persistent_kernel<64, 64' 0 2
timeout 300 $NCU -k 'regex:nms_' -c 3 -f -o gpurun_out/r2base_nms python tools/profile_nms.py > gpurun_out/r2base_nms.log 2>&1; echo "nms rc=$?"
ls -la gpurun_out/*.ncu-rep
