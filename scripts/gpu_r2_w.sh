#!/bin/bash
# round 2, GPU call W: final state — smoke(), whole GPU suite, bench lines of the other configs
mkdir -p gpurun_out
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2 | tee gpurun_out/w_smoke.log
timeout 1800 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -6 | tee gpurun_out/w_pytest.log
timeout 600 python bench.py --workload sd15 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/w_bench_sd15.json 2> gpurun_out/w_bench_sd15.err; echo "sd15 rc $?"; tail -c 300 gpurun_out/w_bench_sd15.err; head -c 300 gpurun_out/w_bench_sd15.json; echo
timeout 900 python bench.py --workload flux --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/w_bench_flux.json 2> gpurun_out/w_bench_flux.err; echo "flux rc $?"; tail -c 300 gpurun_out/w_bench_flux.err; head -c 300 gpurun_out/w_bench_flux.json; echo
timeout 900 python bench.py --sampler dpmpp_2m --steps 2 --warmup 3 --no-cpu-baseline --no-gpu-reference > gpurun_out/w_bench_dpmpp.json 2> gpurun_out/w_bench_dpmpp.err; echo "dpmpp rc $?"; tail -c 300 gpurun_out/w_bench_dpmpp.err; head -c 300 gpurun_out/w_bench_dpmpp.json; echo
timeout 300 python scripts/shape_table.py sd15 2>&1 | grep -v Warn > gpurun_out/w_shapes_sd15.log; head -5 gpurun_out/w_shapes_sd15.log
timeout 300 python scripts/shape_table.py 2>&1 | grep -v Warn > gpurun_out/w_shapes_sdxl.log; head -5 gpurun_out/w_shapes_sdxl.log
