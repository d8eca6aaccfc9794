#!/bin/bash
# round 2, GPU call U: narrow last N tile A/B, full GPU test suite, ncu captures of the new kernels, launch list of the timed region, headline bench
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -6 | tee gpurun_out/u_pytest.log
for v in 1 0 1 0; do
  echo "-- B200_GEMM_NARROW_LAST=$v" | tee -a gpurun_out/u_gemm.log
  B200_GEMM_NARROW_LAST=$v timeout 300 python scripts/kernel_perf.py gemm conv 2>&1 | grep "N=640 \|N=1920\|640+0->640\|640->640\|+640->640" | cut -c1-150 | tee -a gpurun_out/u_gemm.log
done
for i in 1 2; do
  B200_GEMM_NARROW_LAST=1 timeout 300 python scripts/unet_step_time.py 2>&1 | tail -2 | tee -a gpurun_out/u_step.log
  B200_GEMM_NARROW_LAST=0 timeout 300 python scripts/unet_step_time.py 2>&1 | tail -1 | tee -a gpurun_out/u_step.log
done
echo "== ncu"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'gemm_kernel' -c 4 -o gpurun_out/u_ncu_upconv python scripts/ncu_target.py upconv > gpurun_out/u_ncu_upconv.log 2>&1; tail -1 gpurun_out/u_ncu_upconv.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'attn64s_kernel' -s 1 -c 1 -o gpurun_out/u_ncu_attn128s python scripts/ncu_target.py attn128 > gpurun_out/u_ncu_attn128.log 2>&1; tail -1 gpurun_out/u_ncu_attn128.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'gemm_kernel' -s 1 -c 1 -o gpurun_out/u_ncu_ksplit python scripts/ncu_target.py ksplit > gpurun_out/u_ncu_ksplit.log 2>&1; tail -1 gpurun_out/u_ncu_ksplit.log
echo "== launch list (timed region of one job: 2 sampler steps + VAE decode)"
B200_PROFILE_TIMED=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/u_launches.csv \
   python bench.py --steps 1 --warmup 1 --sampler_steps 2 --no-cpu-baseline --no-gpu-reference --no-parity > gpurun_out/u_bench_under_ncu.json 2>/dev/null; echo "exit $?"; wc -l gpurun_out/u_launches.csv
echo "== headline bench (driver's command)"
timeout 1500 python bench.py > gpurun_out/u_bench_sdxl.json 2> gpurun_out/u_bench_sdxl.err; echo "rc $?"; tail -c 400 gpurun_out/u_bench_sdxl.err; head -c 600 gpurun_out/u_bench_sdxl.json; echo
