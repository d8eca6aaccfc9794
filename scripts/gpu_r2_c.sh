#!/bin/bash
# round 2, GPU call C: exp-phase instruction-mix microbenchmark; the rewritten GEMM epilogue (tests, sweep, shape table, step time)
mkdir -p gpurun_out
echo "== exp mix microbench"; timeout 120 ./scripts/micro/exp_mix_bench > gpurun_out/c_expmix.log 2>&1; cat gpurun_out/c_expmix.log
echo "== tests"
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_unet_gpu.py tests/test_flux_gpu.py tests/test_vae_gpu.py tests/test_bench_shapes_gpu.py -m gpu -q -p no:cacheprovider -k "not flux_full_depth" > gpurun_out/c_pytest.log 2>&1
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/c_pytest.log | tail -8
echo "== gemm sweep"
timeout 400 python scripts/gemm_bn_sweep.py 2>&1 | grep -v Warn > gpurun_out/c_gemm_sweep.log
grep -E "auto|ff1" gpurun_out/c_gemm_sweep.log
echo "== shape table"
timeout 300 python scripts/shape_table.py sdxl 2>&1 | grep -v Warn > gpurun_out/c_shapes_sdxl.log
head -24 gpurun_out/c_shapes_sdxl.log
echo "== step time"
for wl in sdxl sd15; do timeout 300 python scripts/unet_step_time.py $wl 2>&1 | grep -v Warn | tail -2; done | tee gpurun_out/c_step.log
