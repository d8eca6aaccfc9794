#!/bin/bash
# round 2, GPU call R: Dh = 128 attention through the small-CTA kernel (attention64s.cu built for Dh = 128, 2 CTAs / SM)
mkdir -p gpurun_out
V=stable-diffusion-webui-forge_b200/variants
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "attention" -p no:cacheprovider 2>&1 | tail -12 | tee gpurun_out/r_pytest_attn.log
run() { echo "-- $1" | tee -a gpurun_out/r_attn.log; shift; env "$@" timeout 200 python scripts/kernel_perf.py attention 2>&1 | grep "attention B" | cut -c1-130 | tee -a gpurun_out/r_attn.log; }
run "small-CTA Dh128 (poly 1 of 4)" X=1
run "old attn128 kernel" B200_ATTN128_VER=0
run "small-CTA poly 0" B200FORGE_LIB=$V/lib_d128p0.so
run "small-CTA poly 2 of 4" B200FORGE_LIB=$V/lib_d128p3.so
run "small-CTA Dh128 (poly 1 of 4)" X=1
timeout 900 python -m pytest tests/test_flux_gpu.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -6 | tee gpurun_out/r_pytest_flux.log
timeout 600 python scripts/flux_perf.py 2>&1 | tail -12 | tee gpurun_out/r_flux_perf.log
B200_ATTN128_VER=0 timeout 600 python scripts/flux_perf.py 2>&1 | tail -12 | tee -a gpurun_out/r_flux_perf.log
