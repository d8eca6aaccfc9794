#!/bin/bash
# Round-end measurement bundle (one gpurun call): GPU test suite, SDXL bench (+ reference arm), Flux bench, ncu launch
# list of exactly the timed region of one bench step, ncu full captures of the top kernels.
mkdir -p gpurun_out
echo "== pytest -m gpu (${B200_BUNDLE_TESTS:-tests})"; timeout 900 python -m pytest ${B200_BUNDLE_TESTS:-tests} -q -m gpu -p no:cacheprovider 2>&1 | tail -3
echo "== bench sdxl"; timeout 600 python bench.py > gpurun_out/bench_r1.json 2> gpurun_out/bench_r1.err; echo "exit $?"; tail -c 200 gpurun_out/bench_r1.err; cut -c1-400 gpurun_out/bench_r1.json
echo "== bench --impl reference"; timeout 400 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref_r1.json 2>/dev/null; echo "exit $?"; cut -c1-300 gpurun_out/bench_ref_r1.json
echo "== bench flux"; timeout 600 python bench.py --workload flux --steps 2 --warmup 1 > gpurun_out/bench_flux_r1.json 2> gpurun_out/bench_flux_r1.err; echo "exit $?"; tail -c 200 gpurun_out/bench_flux_r1.err; cut -c1-400 gpurun_out/bench_flux_r1.json
echo "== launch list (timed region of one job: 2 sampler steps + VAE decode)"
B200_PROFILE_TIMED=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_r1.csv \
   python bench.py --steps 1 --warmup 1 --sampler_steps 2 --no-cpu-baseline --no-gpu-reference > gpurun_out/bench_under_ncu.json 2>/dev/null; echo "exit $?"; wc -l gpurun_out/launches_r1.csv
echo "== ncu full"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:'gemm_kernel|attn64_kernel|attn128_kernel|gn_apply|gn_stats' -c 14 -o gpurun_out/ncu_r1_kernels python scripts/ncu_target.py all > gpurun_out/ncu_r1.log 2>&1; tail -1 gpurun_out/ncu_r1.log; ls -la gpurun_out | head -20
