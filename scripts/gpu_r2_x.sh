#!/bin/bash
# round 2, GPU call X: GEMM epilogue with 224 registers (setmaxnreg by warpgroup) + accumulator chunk loaded one chunk ahead
mkdir -p gpurun_out
V=stable-diffusion-webui-forge_b200/variants
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "gemm or conv" -p no:cacheprovider 2>&1 | tail -4 | tee gpurun_out/x_pytest.log
run() { echo "-- $1" | tee -a gpurun_out/x_gemm.log; shift; env "$@" timeout 300 python scripts/kernel_perf.py gemm 2>&1 | grep "gemm M\|geglu\|ln" | cut -c1-150 | tee -a gpurun_out/x_gemm.log; env "$@" timeout 300 python scripts/ln_fold_perf.py 2>&1 | tail -12 | cut -c1-150 | tee -a gpurun_out/x_gemm.log; }
run "setmaxnreg 56/224 + prefetch" X=1
run "previous commit" B200FORGE_LIB=$V/lib_gprev.so
run "prefetch only (168 registers)" B200FORGE_LIB=$V/lib_gnosm.so
for i in 1 2; do
  timeout 300 python scripts/unet_step_time.py 2>&1 | tail -2 | tee -a gpurun_out/x_step.log
  B200FORGE_LIB=$V/lib_gprev.so timeout 300 python scripts/unet_step_time.py 2>&1 | tail -1 | tee -a gpurun_out/x_step.log
  B200FORGE_LIB=$V/lib_gnosm.so timeout 300 python scripts/unet_step_time.py 2>&1 | tail -1 | tee -a gpurun_out/x_step.log
done
timeout 300 python scripts/shape_table.py 2>&1 | grep -v Warn | head -14 > gpurun_out/x_shapes_sdxl.log; cat gpurun_out/x_shapes_sdxl.log
timeout 1500 python -m pytest tests/test_unet_gpu.py tests/test_bench_shapes_gpu.py tests/test_flux_gpu.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -4 | tee gpurun_out/x_pytest_engines.log
