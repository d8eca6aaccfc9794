#!/bin/bash
# round 2, GPU call L: P of the Dh=64 attention kernel in TMEM (tcgen05.st + tcgen05.mma with A from TMEM) against P through smem
mkdir -p gpurun_out
V=stable-diffusion-webui-forge_b200/variants
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "attention" -p no:cacheprovider 2>&1 | tail -8 | tee gpurun_out/l_pytest_attn.log
for v in main spt0 s0 s3 main spt0; do
  echo "-- $v" | tee -a gpurun_out/l_attn.log
  if [ $v = main ]; then L=""; else L=$V/lib_$v.so; fi
  B200FORGE_LIB=$L timeout 200 python scripts/kernel_perf.py attention 2>&1 | grep "attention B" | head -4 | tee -a gpurun_out/l_attn.log
done
timeout 300 python scripts/unet_step_time.py 2>&1 | tail -1 | tee gpurun_out/l_step.log
B200FORGE_LIB=$V/lib_spt0.so timeout 300 python scripts/unet_step_time.py 2>&1 | tail -1 | tee -a gpurun_out/l_step.log
timeout 300 python scripts/unet_step_time.py 2>&1 | tail -1 | tee -a gpurun_out/l_step.log
