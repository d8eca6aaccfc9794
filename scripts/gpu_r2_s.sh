#!/bin/bash
# round 2, GPU call S: K-split of the last partly filled wave of GEMM / conv tiles (B200_GEMM_SPLITK) — parity, A/B
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "gemm or conv or attention_dh128" -p no:cacheprovider 2>&1 | tail -12 | tee gpurun_out/s_pytest.log
for v in 1 0 1 0; do
  echo "-- B200_GEMM_SPLITK=$v" | tee -a gpurun_out/s_gemm.log
  B200_GEMM_SPLITK=$v timeout 300 python scripts/kernel_perf.py gemm conv 2>&1 | grep "^gemm\|^conv\|gemm M\|conv3x3" | cut -c1-150 | tee -a gpurun_out/s_gemm.log
done
for i in 1 2; do
  B200_GEMM_SPLITK=1 timeout 300 python scripts/unet_step_time.py 2>&1 | tail -2 | tee -a gpurun_out/s_step.log
  B200_GEMM_SPLITK=0 timeout 300 python scripts/unet_step_time.py 2>&1 | tail -1 | tee -a gpurun_out/s_step.log
done
B200_GEMM_SPLITK=1 timeout 300 python scripts/unet_step_time.py sd15 2>&1 | tail -1 | tee -a gpurun_out/s_step.log
B200_GEMM_SPLITK=0 timeout 300 python scripts/unet_step_time.py sd15 2>&1 | tail -1 | tee -a gpurun_out/s_step.log
timeout 600 python scripts/flux_perf.py 2>&1 | tail -12 | tee gpurun_out/s_flux_perf.log
B200_ATTN128_VER=0 B200_GEMM_SPLITK=0 timeout 600 python scripts/flux_perf.py 2>&1 | tail -12 | tee -a gpurun_out/s_flux_perf.log
timeout 1500 python -m pytest tests/test_unet_gpu.py tests/test_vae_gpu.py tests/test_bench_shapes_gpu.py tests/test_flux_gpu.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -8 | tee gpurun_out/s_pytest_engines.log
