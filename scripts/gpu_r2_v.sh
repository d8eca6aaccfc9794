#!/bin/bash
# round 2, GPU call V: producer loop without run-time divisions (conv regression fix) — conv parity, conv perf, step time, launch list, bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_vae_gpu.py -m gpu -q -x -k "conv or vae or gemm" -p no:cacheprovider 2>&1 | tail -4 | tee gpurun_out/v_pytest.log
timeout 300 python scripts/kernel_perf.py conv 2>&1 | grep "conv3x3" | cut -c1-150 | tee gpurun_out/v_conv.log
timeout 300 python scripts/upconv_perf.py 2>&1 | tee -a gpurun_out/v_conv.log
timeout 300 python scripts/unet_step_time.py 2>&1 | tail -2 | tee gpurun_out/v_step.log
timeout 300 python scripts/unet_step_time.py 2>&1 | tail -1 | tee -a gpurun_out/v_step.log
timeout 300 python scripts/unet_step_time.py sd15 2>&1 | tail -1 | tee -a gpurun_out/v_step.log
echo "== launch list (timed region of one job: 2 sampler steps + VAE decode)"
B200_PROFILE_TIMED=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/v_launches.csv \
   python bench.py --steps 1 --warmup 1 --sampler_steps 2 --no-cpu-baseline --no-gpu-reference --no-parity > gpurun_out/v_bench_under_ncu.json 2>/dev/null; echo "exit $?"; wc -l gpurun_out/v_launches.csv
echo "== headline bench (driver's command)"
timeout 1500 python bench.py > gpurun_out/v_bench_sdxl.json 2> gpurun_out/v_bench_sdxl.err; echo "rc $?"; tail -c 400 gpurun_out/v_bench_sdxl.err; head -c 400 gpurun_out/v_bench_sdxl.json; echo
