#!/bin/bash
# round 2, GPU call J: GEMM epilogue with the bias/LN rows staged before the accumulator wait and the residual prefetched one chunk ahead
mkdir -p gpurun_out
V=stable-diffusion-webui-forge_b200/variants
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_bench_shapes_gpu.py -m gpu -q -p no:cacheprovider 2>&1 | tail -4 | tee gpurun_out/j_pytest.log
for rep in 1 2; do
echo "-- new epilogue" | tee -a gpurun_out/j_gemm_sweep.log
timeout 300 python scripts/gemm_bn_sweep.py auto 2>&1 | grep -v Warn | tee -a gpurun_out/j_gemm_sweep.log
echo "-- old epilogue" | tee -a gpurun_out/j_gemm_sweep.log
B200FORGE_LIB=$V/lib_gemm_old.so timeout 300 python scripts/gemm_bn_sweep.py auto 2>&1 | grep -v Warn | tee -a gpurun_out/j_gemm_sweep.log
done
timeout 300 python scripts/unet_step_time.py 2>&1 | tail -1 | tee gpurun_out/j_step.log
B200FORGE_LIB=$V/lib_gemm_old.so timeout 300 python scripts/unet_step_time.py 2>&1 | tail -1 | tee -a gpurun_out/j_step.log
timeout 300 python scripts/unet_step_time.py 2>&1 | tail -1 | tee -a gpurun_out/j_step.log
B200FORGE_LIB=$V/lib_gemm_old.so timeout 300 python scripts/unet_step_time.py 2>&1 | tail -1 | tee -a gpurun_out/j_step.log
