#!/bin/bash
# round 2, GPU call E: small-CTA attention polynomial-exp2 variants; ncu full captures of the small-CTA attention kernel and
# the transformer GEMMs (producer with residual + statistics, LayerNorm-fold consumer, GEGLU)
mkdir -p gpurun_out
V=stable-diffusion-webui-forge_b200/variants
for v in s1 s5 s7; do
  echo "-- $v" | tee -a gpurun_out/e_attn.log
  B200_ATTN64_VER=2 B200FORGE_LIB=$V/lib_$v.so timeout 200 python scripts/kernel_perf.py attention 2>&1 | grep "attention B" | head -2 | tee -a gpurun_out/e_attn.log
done
echo "-- main lib VER=2, then VER=1" | tee -a gpurun_out/e_attn.log
B200_ATTN64_VER=2 timeout 200 python scripts/kernel_perf.py attention 2>&1 | grep "attention B" | head -2 | tee -a gpurun_out/e_attn.log
B200_ATTN64_VER=1 timeout 200 python scripts/kernel_perf.py attention 2>&1 | grep "attention B" | head -2 | tee -a gpurun_out/e_attn.log
echo "== same-process step-time A/B"
for v in 1 2; do B200_ATTN64_VER=$v timeout 300 python scripts/unet_step_time.py sdxl 40 2>&1 | grep -v Warn | tail -1 | sed "s/^/VER=$v /"; done | tee gpurun_out/e_step.log
for v in 1 2; do B200_ATTN64_VER=$v timeout 300 python scripts/unet_step_time.py sdxl 40 2>&1 | grep -v Warn | tail -1 | sed "s/^/VER=$v /"; done | tee -a gpurun_out/e_step.log
echo "== ncu"
B200_ATTN64_VER=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:'attn64s_kernel' -s 1 -c 1 -o gpurun_out/e_ncu_attn64s python scripts/ncu_target.py attn > gpurun_out/e_ncu_attn.log 2>&1; tail -2 gpurun_out/e_ncu_attn.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'gemm_kernel' -c 8 -o gpurun_out/e_ncu_gemm python scripts/ncu_target.py unetgemm > gpurun_out/e_ncu_gemm.log 2>&1; tail -2 gpurun_out/e_ncu_gemm.log
ls -la gpurun_out/*.ncu-rep
