#!/bin/bash
# round 2, GPU call D: the small-CTA Dh=64 attention kernel (VER 2): parity tests, perf beside VER 1 and SDPA, cross-attention
# routing, fast-GELU GEGLU epilogue, UNet step time
mkdir -p gpurun_out
echo "== attention tests VER=2"
B200_ATTN64_VER=2 B200_ATTN64_CROSS=s timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "attention" -p no:cacheprovider 2>&1 | tail -6 | tee gpurun_out/d_pytest_attn2.log
echo "== perf"
for v in 1 2; do
  echo "-- VER=$v" | tee -a gpurun_out/d_attn.log
  B200_ATTN64_VER=$v timeout 200 python scripts/kernel_perf.py attention 2>&1 | grep "attention B" | tee -a gpurun_out/d_attn.log
done
echo "-- VER=2 cross=s" | tee -a gpurun_out/d_attn.log
B200_ATTN64_VER=2 B200_ATTN64_CROSS=s timeout 200 python scripts/kernel_perf.py attention 2>&1 | grep "attention B" | tee -a gpurun_out/d_attn.log
echo "-- VER=1 no turns (variant p0nt)" | tee -a gpurun_out/d_attn.log
B200FORGE_LIB=stable-diffusion-webui-forge_b200/variants/lib_p0nt.so timeout 200 python scripts/kernel_perf.py attention 2>&1 | grep "attention B" | head -2 | tee -a gpurun_out/d_attn.log
echo "== gemm tests + geglu perf"
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_unet_gpu.py -m gpu -q -p no:cacheprovider 2>&1 | tail -4 | tee gpurun_out/d_pytest.log
timeout 300 python scripts/gemm_bn_sweep.py 2>&1 | grep -E "ff1" | tee gpurun_out/d_ff1.log
echo "== step time"
for v in 1 2; do B200_ATTN64_VER=$v timeout 300 python scripts/unet_step_time.py sdxl 2>&1 | grep -v Warn | tail -1 | sed "s/^/VER=$v /"; done | tee gpurun_out/d_step.log
B200_ATTN64_VER=2 B200_ATTN64_CROSS=s timeout 300 python scripts/unet_step_time.py sdxl 2>&1 | grep -v Warn | tail -2 | sed "s/^/VER=2 cross=s /" | tee -a gpurun_out/d_step.log
B200_ATTN64_VER=2 timeout 300 python scripts/unet_step_time.py sd15 2>&1 | grep -v Warn | tail -1 | sed "s/^/VER=2 /" | tee -a gpurun_out/d_step.log
