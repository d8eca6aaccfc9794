#!/bin/bash
# round 2, GPU call N: ncu captures of the Dh=64 attention kernel with P in TMEM (self L4096 and cross Lk=77)
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'attn64s_kernel' -s 1 -c 1 -o gpurun_out/n_ncu_attn64s python scripts/ncu_target.py attn > gpurun_out/n_ncu_attn.log 2>&1; tail -2 gpurun_out/n_ncu_attn.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'attn64s_kernel' -s 1 -c 1 -o gpurun_out/n_ncu_attn64s_cross python scripts/ncu_target.py attncross > gpurun_out/n_ncu_attnc.log 2>&1; tail -2 gpurun_out/n_ncu_attnc.log
