#!/bin/bash
# Round-end measurement bundle (one gpurun call): full GPU test suite, bench, ncu launch list, ncu full captures.
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -q -m gpu --timeout 900 -p no:cacheprovider -x 2>&1 | tail -4
echo "== bench"; timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r1.json 2> gpurun_out/bench_r1.err; echo "exit $?"; tail -c 300 gpurun_out/bench_r1.err
echo "== launch list"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 2600 -c 2600 --csv --log-file gpurun_out/launches_r1.csv \
   python bench.py --steps 1 --warmup 1 --sampler_steps 2 --no-cpu-baseline --no-gpu-reference > gpurun_out/bench_under_ncu.json 2>/dev/null; echo "exit $?"
echo "== ncu full"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:'gemm_kernel|attn64_kernel|gn_apply|gn_stats' -s 8 -c 10 -o gpurun_out/ncu_r1_kernels python scripts/ncu_target.py all > gpurun_out/ncu_r1.log 2>&1; tail -2 gpurun_out/ncu_r1.log
