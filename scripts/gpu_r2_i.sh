#!/bin/bash
# round 2, GPU call I: deferred row maximum in the small-CTA Dh=64 attention kernel (A/B against the variant without it)
mkdir -p gpurun_out
V=stable-diffusion-webui-forge_b200/variants
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "attention" -p no:cacheprovider 2>&1 | tail -4 | tee gpurun_out/i_pytest_attn.log
for rep in 1 2; do
echo "-- main (deferred max)" | tee -a gpurun_out/i_attn.log
timeout 200 python scripts/kernel_perf.py attention 2>&1 | grep "attention B" | head -4 | tee -a gpurun_out/i_attn.log
echo "-- sd0 (max first)" | tee -a gpurun_out/i_attn.log
B200FORGE_LIB=$V/lib_sd0.so timeout 200 python scripts/kernel_perf.py attention 2>&1 | grep "attention B" | head -4 | tee -a gpurun_out/i_attn.log
done
timeout 300 python scripts/unet_step_time.py 2>&1 | tail -3 | tee gpurun_out/i_step.log
B200FORGE_LIB=$V/lib_sd0.so timeout 300 python scripts/unet_step_time.py 2>&1 | tail -3 | tee -a gpurun_out/i_step.log
