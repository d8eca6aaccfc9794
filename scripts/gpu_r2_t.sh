#!/bin/bash
# round 2, GPU call T: K-split with the fast exchange (pre-pass, all loads in flight) + caller-owned workspace (works inside
# CUDA graphs); Dh = 128 attention with double-buffered P
mkdir -p gpurun_out
V=stable-diffusion-webui-forge_b200/variants
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "gemm or conv or attention_dh128 or attention" -p no:cacheprovider 2>&1 | tail -6 | tee gpurun_out/t_pytest.log
for v in 1 0 1 0; do
  echo "-- B200_GEMM_SPLITK=$v" | tee -a gpurun_out/t_gemm.log
  B200_GEMM_SPLITK=$v timeout 300 python scripts/kernel_perf.py gemm conv 2>&1 | grep "gemm M\|conv3x3" | cut -c1-150 | tee -a gpurun_out/t_gemm.log
done
echo "-- B200_GEMM_SPLITK_MIN=8" | tee -a gpurun_out/t_gemm.log
B200_GEMM_SPLITK_MIN=8 timeout 300 python scripts/kernel_perf.py gemm conv 2>&1 | grep "gemm M\|conv3x3" | cut -c1-150 | tee -a gpurun_out/t_gemm.log
for i in 1 2; do
  B200_GEMM_SPLITK=1 timeout 300 python scripts/unet_step_time.py 2>&1 | tail -2 | tee -a gpurun_out/t_step.log
  B200_GEMM_SPLITK=0 timeout 300 python scripts/unet_step_time.py 2>&1 | tail -1 | tee -a gpurun_out/t_step.log
done
B200_GEMM_SPLITK_MIN=8 timeout 300 python scripts/unet_step_time.py 2>&1 | tail -1 | tee -a gpurun_out/t_step.log
B200_GEMM_SPLITK=1 timeout 300 python scripts/unet_step_time.py sd15 2>&1 | tail -1 | tee -a gpurun_out/t_step.log
B200_GEMM_SPLITK=0 timeout 300 python scripts/unet_step_time.py sd15 2>&1 | tail -1 | tee -a gpurun_out/t_step.log
run() { echo "-- $1" | tee -a gpurun_out/t_attn.log; shift; env "$@" timeout 200 python scripts/kernel_perf.py attention 2>&1 | grep "Dh=128" | cut -c1-130 | tee -a gpurun_out/t_attn.log; }
run "Dh128 P double-buffered (poly 1 of 4)" X=1
run "Dh128 single P buffer" B200FORGE_LIB=$V/lib_d128p2off.so
run "Dh128 P2 poly 0" B200FORGE_LIB=$V/lib_d128p2poly0.so
run "Dh128 P2 poly 2 of 4" B200FORGE_LIB=$V/lib_d128p2poly3.so
run "Dh128 P double-buffered (poly 1 of 4)" X=1
timeout 1500 python -m pytest tests/test_unet_gpu.py tests/test_bench_shapes_gpu.py tests/test_flux_gpu.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -8 | tee gpurun_out/t_pytest_engines.log
