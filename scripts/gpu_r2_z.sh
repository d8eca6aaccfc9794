#!/bin/bash
# round 2, GPU call Z (2 GPUs): the multi-GPU bench path with the final code (replicas, uint8 NCCL gather)
mkdir -p gpurun_out
nvidia-smi -L | head -3
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 3 --no-cpu-baseline --no-gpu-reference > gpurun_out/z_bench_2gpu.json 2> gpurun_out/z_bench_2gpu.err; echo "rc $?"; tail -c 400 gpurun_out/z_bench_2gpu.err; head -c 500 gpurun_out/z_bench_2gpu.json; echo
