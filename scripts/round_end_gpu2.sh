#!/bin/bash
mkdir -p gpurun_out
REGEX='regex:gemm_kernel|attn|gn_|layernorm|sampler|im2col|upsample|nchw|nhwc|silu_kernel|timestep|unet_input|softmax_rows|vae_post|eps_to'
echo "== parity after kernel changes"; timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_unet_gpu.py -q -m gpu -k "groupnorm or attention or golden or full_width" --timeout 600 -p no:cacheprovider -s 2>&1 | grep -E "unet|passed|failed|rror" | cut -c1-160 | tail -12
echo "== launch list (one timed job: 2 sampler steps + VAE decode)"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$REGEX" -s 4250 -c 2400 --csv --log-file gpurun_out/launches_r1.csv \
   python bench.py --steps 1 --warmup 1 --sampler_steps 2 --no-cpu-baseline --no-gpu-reference > gpurun_out/bench_under_ncu.json 2>/dev/null; echo "exit $?"
echo "== ncu full"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:'gemm_kernel|attn64_kernel|gn_apply|gn_stats' -c 10 -o gpurun_out/ncu_r1_kernels python scripts/ncu_target.py all > gpurun_out/ncu_r1.log 2>&1; tail -1 gpurun_out/ncu_r1.log
