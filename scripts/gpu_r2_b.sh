#!/bin/bash
# round 2, GPU call B: the tests fixed after call A + the new ones, attention variants + phase profiles, per-shape table,
# GEMM block_n sweep
mkdir -p gpurun_out
V=stable-diffusion-webui-forge_b200/variants
timeout 900 python -m pytest tests -m gpu -q -rA -p no:cacheprovider -k "trajectory_psnr or chroma or p2_operations or p5_vae_encode or tiled or flux_vae or other_unet or lms_and or 16_channel" > gpurun_out/b_pytest.log 2>&1
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/b_pytest.log | tail -8
echo "== attention variants (VER 1)"
for v in p0 p4 p6 p0nt p4nt p6nt; do
  echo "-- $v" >> gpurun_out/b_attn_variants.log
  B200FORGE_LIB=$V/lib_$v.so timeout 200 python scripts/kernel_perf.py attention 2>&1 | grep "attention B" | head -2 >> gpurun_out/b_attn_variants.log
done
cat gpurun_out/b_attn_variants.log
echo "== phase profiles"
for v in p4 p0 p4nt; do
  echo "-- $v" >> gpurun_out/b_attn_phase.log
  B200FORGE_LIB=$V/libprof_$v.so timeout 200 python scripts/attn_phase_profile.py 2>&1 | grep -E "profile|^--" >> gpurun_out/b_attn_phase.log
done
cat gpurun_out/b_attn_phase.log
echo "== shape table"
timeout 300 python scripts/shape_table.py sdxl vae 2>&1 | grep -v Warn > gpurun_out/b_shapes_sdxl.log
head -45 gpurun_out/b_shapes_sdxl.log
echo "== gemm sweep"
timeout 400 python scripts/gemm_bn_sweep.py 2>&1 | grep -v Warn > gpurun_out/b_gemm_sweep.log
cat gpurun_out/b_gemm_sweep.log
