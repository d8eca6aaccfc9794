#!/bin/bash
# Runs the kernel-level GPU parity tests, one pytest process per kernel family so that a device-side
# trap in one family does not poison the CUDA context of the others.  Logs land in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
for grp in "gemm" "conv" "attention" "groupnorm or layernorm" "layout or unet_input or softmax or sampler"; do
  tag=$(echo "$grp" | tr ' ' '_')
  echo "=== $grp ==="
  timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "$grp" -s --timeout 180 -p no:cacheprovider \
     > "gpurun_out/k_${tag}.log" 2>&1
  echo "exit $?"
  grep -E "passed|failed|error" "gpurun_out/k_${tag}.log" | tail -2
  grep -E "^\[parity\]|watchdog|FAILED|Error" "gpurun_out/k_${tag}.log" | head -60
done
