#!/bin/bash
# round 2, GPU call K: GEMM epilogue (residual prefetch: one chunk ahead in registers + next tile's lines into L2), block_n = 256 for N >= 512
mkdir -p gpurun_out
V=stable-diffusion-webui-forge_b200/variants
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_bench_shapes_gpu.py tests/test_unet_gpu.py -m gpu -q -p no:cacheprovider 2>&1 | tail -4 | tee gpurun_out/k_pytest.log
echo "-- new epilogue" | tee gpurun_out/k_gemm_sweep.log
timeout 300 python scripts/gemm_bn_sweep.py 2>&1 | grep -E "auto| 256" | tee -a gpurun_out/k_gemm_sweep.log
echo "-- old epilogue" | tee -a gpurun_out/k_gemm_sweep.log
B200FORGE_LIB=$V/lib_gemm_old.so timeout 300 python scripts/gemm_bn_sweep.py 2>&1 | grep -E "auto| 256" | tee -a gpurun_out/k_gemm_sweep.log
for rep in 1 2; do
timeout 300 python scripts/unet_step_time.py 2>&1 | tail -1 | tee -a gpurun_out/k_step.log
B200FORGE_LIB=$V/lib_gemm_old.so timeout 300 python scripts/unet_step_time.py 2>&1 | tail -1 | tee -a gpurun_out/k_step.log
done
timeout 300 python scripts/shape_table.py sdxl 2>&1 | grep -v Warn > gpurun_out/k_shapes_sdxl.log; head -32 gpurun_out/k_shapes_sdxl.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'gemm_kernel' -c 4 -o gpurun_out/k_ncu_producer python scripts/ncu_target.py producer > gpurun_out/k_ncu.log 2>&1; tail -2 gpurun_out/k_ncu.log
