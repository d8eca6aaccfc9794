#!/bin/bash
# round 2, GPU call Y: final code — smoke(), whole GPU suite, launch list of the timed region, headline bench (driver's command)
mkdir -p gpurun_out
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2 | tee gpurun_out/y_smoke.log
timeout 1800 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -6 | tee gpurun_out/y_pytest.log
echo "== launch list (timed region of one job: 2 sampler steps + VAE decode)"
B200_PROFILE_TIMED=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/y_launches.csv \
   python bench.py --steps 1 --warmup 1 --sampler_steps 2 --no-cpu-baseline --no-gpu-reference --no-parity > gpurun_out/y_bench_under_ncu.json 2>/dev/null; echo "exit $?"; wc -l gpurun_out/y_launches.csv
echo "== headline bench (driver's command)"
timeout 1500 python bench.py > gpurun_out/y_bench_sdxl.json 2> gpurun_out/y_bench_sdxl.err; echo "rc $?"; tail -c 400 gpurun_out/y_bench_sdxl.err; head -c 300 gpurun_out/y_bench_sdxl.json; echo
timeout 600 python bench.py --impl reference > gpurun_out/y_bench_reference.json 2> gpurun_out/y_bench_reference.err; echo "ref rc $?"; head -c 600 gpurun_out/y_bench_reference.json; echo
