#!/bin/bash
# round 2, GPU call F: everything with the new defaults: full GPU suite, smoke, shape table, SDXL bench line
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -rA -p no:cacheprovider > gpurun_out/f_pytest.log 2>&1
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/f_pytest.log | tail -8
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/f_smoke.log
timeout 300 python scripts/shape_table.py sdxl 2>&1 | grep -v Warn > gpurun_out/f_shapes_sdxl.log; head -16 gpurun_out/f_shapes_sdxl.log
timeout 900 python bench.py --steps 2 --warmup 3 > gpurun_out/f_bench.json 2> gpurun_out/f_bench.err; echo "bench rc $?"; tail -c 600 gpurun_out/f_bench.err; head -c 700 gpurun_out/f_bench.json
