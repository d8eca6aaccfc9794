#!/bin/bash
# round 2, GPU call O: attention64s — P.V wait deferred to the first P store, cheap partial blocks, several query tiles per CTA for short key sequences
mkdir -p gpurun_out
V=stable-diffusion-webui-forge_b200/variants
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "attention" -p no:cacheprovider 2>&1 | tail -8 | tee gpurun_out/o_pytest_attn.log
B200_ATTN64S_TP=3 timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "attention" -p no:cacheprovider 2>&1 | tail -4 | tee -a gpurun_out/o_pytest_attn.log
run() { echo "-- $1" | tee -a gpurun_out/o_attn.log; shift; env "$@" timeout 200 python scripts/kernel_perf.py attention 2>&1 | grep "attention B" | head -4 | cut -c1-120 | tee -a gpurun_out/o_attn.log; }
run "main (tp auto)" X=1
run "prev" B200FORGE_LIB=$V/lib_aprev.so
run "main tp=1" B200_ATTN64S_TP=1
run "main tp=2" B200_ATTN64S_TP=2
run "main tp=4" B200_ATTN64S_TP=4
run "main (tp auto)" X=1
run "prev" B200FORGE_LIB=$V/lib_aprev.so
timeout 300 python scripts/unet_step_time.py 2>&1 | tail -1 | tee gpurun_out/o_step.log
B200FORGE_LIB=$V/lib_aprev.so timeout 300 python scripts/unet_step_time.py 2>&1 | tail -1 | tee -a gpurun_out/o_step.log
timeout 300 python scripts/unet_step_time.py 2>&1 | tail -1 | tee -a gpurun_out/o_step.log
