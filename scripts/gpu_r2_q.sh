#!/bin/bash
# round 2, GPU call Q: upsample folded into the convolution (b200_conv3x3_up2x) — parity, A/B against upsample2x + conv3x3
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "conv3x3" -p no:cacheprovider 2>&1 | tail -15 | tee gpurun_out/q_pytest_conv.log
timeout 300 python scripts/upconv_perf.py 2>&1 | tee gpurun_out/q_upconv.log
for i in 1 2; do
  B200_UPCONV=1 timeout 300 python scripts/unet_step_time.py 2>&1 | tail -2 | tee -a gpurun_out/q_step.log
  B200_UPCONV=0 timeout 300 python scripts/unet_step_time.py 2>&1 | tail -1 | tee -a gpurun_out/q_step.log
done
timeout 1500 python -m pytest tests/test_unet_gpu.py tests/test_vae_gpu.py tests/test_bench_shapes_gpu.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -15 | tee gpurun_out/q_pytest_engines.log
