#!/bin/bash
# round 2, GPU call G: tiled-VAE fix, SD1.5 shape table + bench line, Flux bench line, ncu launch list of the timed region
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_vae_gpu.py tests/test_flux_gpu.py tests/test_plugin_gpu.py -m gpu -q -p no:cacheprovider 2>&1 | tail -4 | tee gpurun_out/g_pytest.log
timeout 300 python scripts/shape_table.py sd15 2>&1 | grep -v Warn > gpurun_out/g_shapes_sd15.log; head -30 gpurun_out/g_shapes_sd15.log
timeout 600 python bench.py --workload sd15 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/g_bench_sd15.json 2> gpurun_out/g_bench_sd15.err; echo "sd15 rc $?"; tail -c 300 gpurun_out/g_bench_sd15.err; head -c 400 gpurun_out/g_bench_sd15.json; echo
timeout 900 python bench.py --workload flux --steps 2 --warmup 3 > gpurun_out/g_bench_flux.json 2> gpurun_out/g_bench_flux.err; echo "flux rc $?"; tail -c 300 gpurun_out/g_bench_flux.err; head -c 400 gpurun_out/g_bench_flux.json; echo
echo "== launch list (timed region of one job: 2 sampler steps + VAE decode)"
B200_PROFILE_TIMED=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/g_launches.csv \
   python bench.py --steps 1 --warmup 1 --sampler_steps 2 --no-cpu-baseline --no-gpu-reference --no-parity > gpurun_out/g_bench_under_ncu.json 2>/dev/null; echo "exit $?"; wc -l gpurun_out/g_launches.csv
