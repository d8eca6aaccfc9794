#!/bin/bash
# round 2, GPU call M: three epilogue warps per TMEM quadrant (setmaxnreg) against two
mkdir -p gpurun_out
V=stable-diffusion-webui-forge_b200/variants
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_bench_shapes_gpu.py tests/test_unet_gpu.py tests/test_flux_gpu.py -m gpu -q -p no:cacheprovider 2>&1 | tail -6 | tee gpurun_out/m_pytest.log
echo "-- 3 epilogue warps / quadrant" | tee gpurun_out/m_gemm_sweep.log
timeout 300 python scripts/gemm_bn_sweep.py 2>&1 | grep -E "auto" | tee -a gpurun_out/m_gemm_sweep.log
echo "-- 2 epilogue warps / quadrant" | tee -a gpurun_out/m_gemm_sweep.log
B200FORGE_LIB=$V/lib_gemm2w.so timeout 300 python scripts/gemm_bn_sweep.py 2>&1 | grep -E "auto" | tee -a gpurun_out/m_gemm_sweep.log
for rep in 1 2; do
timeout 300 python scripts/unet_step_time.py 2>&1 | tail -1 | tee -a gpurun_out/m_step.log
B200FORGE_LIB=$V/lib_gemm2w.so timeout 300 python scripts/unet_step_time.py 2>&1 | tail -1 | tee -a gpurun_out/m_step.log
done
timeout 300 python scripts/shape_table.py sdxl 2>&1 | grep -v Warn > gpurun_out/m_shapes_sdxl.log; head -24 gpurun_out/m_shapes_sdxl.log
