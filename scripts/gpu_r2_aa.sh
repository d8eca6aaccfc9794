#!/bin/bash
# round 2, GPU call AA: ncu --set full of the FINAL kernels: convolution 1280 -> 1280 @32x32 (incremental producer, K-split tail),
# the four transformer GEMM flavours (224-register epilogue), Dh = 64 attention
mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'gemm_kernel' -s 1 -c 1 -o gpurun_out/aa_ncu_conv python scripts/ncu_target.py ksplit > gpurun_out/aa_ncu_conv.log 2>&1; tail -1 gpurun_out/aa_ncu_conv.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'gemm_kernel' -c 8 -o gpurun_out/aa_ncu_gemm python scripts/ncu_target.py unetgemm > gpurun_out/aa_ncu_gemm.log 2>&1; tail -1 gpurun_out/aa_ncu_gemm.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'attn64s_kernel' -s 1 -c 1 -o gpurun_out/aa_ncu_attn64s python scripts/ncu_target.py attn > gpurun_out/aa_ncu_attn.log 2>&1; tail -1 gpurun_out/aa_ncu_attn.log
ls -la gpurun_out/aa_*.ncu-rep
