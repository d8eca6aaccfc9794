#!/bin/bash
# round 2, GPU call A: baseline of the non-kernel-speed changes first (PDL off, round-1 attention issue order), then the new
# attention build, then the PDL A/B
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/a_gpu.txt 2>&1
export B200_PDL=0 B200_ATTN64_VER=0
timeout 1300 python -m pytest tests -m gpu -q -rA -p no:cacheprovider > gpurun_out/a_pytest.log 2>&1
echo "pytest rc $?" >> gpurun_out/a_pytest.log
grep -E "passed|failed|rc " gpurun_out/a_pytest.log | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/a_smoke.log 2>&1
echo "smoke rc $?" >> gpurun_out/a_smoke.log
tail -3 gpurun_out/a_smoke.log
timeout 900 python bench.py --steps 2 --warmup 3 > gpurun_out/a_bench.json 2> gpurun_out/a_bench.err
echo "bench rc $?"
tail -c 1200 gpurun_out/a_bench.err
head -c 1200 gpurun_out/a_bench.json
echo
echo "== attention VER 0 / 1"
B200_ATTN64_VER=0 timeout 300 python scripts/kernel_perf.py attention 2>&1 | grep -v Warn > gpurun_out/a_attn_v0.log
B200_ATTN64_VER=1 timeout 300 python scripts/kernel_perf.py attention 2>&1 | grep -v Warn > gpurun_out/a_attn_v1.log
cat gpurun_out/a_attn_v0.log gpurun_out/a_attn_v1.log
B200_ATTN64_VER=1 timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k attention -p no:cacheprovider 2>&1 | tail -8 > gpurun_out/a_pytest_attn1.log
cat gpurun_out/a_pytest_attn1.log
echo "== PDL A/B"
for v in 0 1; do for p in 0 1; do
  B200_ATTN64_VER=$v B200_PDL=$p timeout 300 python scripts/unet_step_time.py sdxl 2>&1 | grep -v Warn | tail -2 | sed "s/^/VER=$v PDL=$p /" >> gpurun_out/a_pdl.log
done; done
for p in 0 1; do B200_ATTN64_VER=1 B200_PDL=$p timeout 300 python scripts/unet_step_time.py sd15 2>&1 | grep -v Warn | tail -2 | sed "s/^/VER=1 PDL=$p /" >> gpurun_out/a_pdl.log; done
cat gpurun_out/a_pdl.log
B200_ATTN64_VER=1 B200_PDL=1 timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_unet_gpu.py tests/test_vae_gpu.py -m gpu -q -p no:cacheprovider 2>&1 | tail -8 > gpurun_out/a_pytest_pdl.log
cat gpurun_out/a_pytest_pdl.log
