#!/bin/bash
# round 2, GPU call H (2 GPUs): the multi-GPU bench path (replicas, uint8 NCCL gather), headline sampler and configs[3] (DPM++ 2M)
mkdir -p gpurun_out
nvidia-smi -L | head -3
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 3 > gpurun_out/h_bench_2gpu.json 2> gpurun_out/h_bench_2gpu.err; echo "rc $?"; tail -c 400 gpurun_out/h_bench_2gpu.err; head -c 500 gpurun_out/h_bench_2gpu.json; echo
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 2 --warmup 3 --sampler dpmpp_2m > gpurun_out/h_bench_2gpu_dpmpp.json 2> gpurun_out/h_bench_2gpu_dpmpp.err; echo "rc $?"; tail -c 400 gpurun_out/h_bench_2gpu_dpmpp.err; head -c 500 gpurun_out/h_bench_2gpu_dpmpp.json; echo
timeout 600 python bench.py --steps 2 --warmup 3 --sampler dpmpp_2m --no-cpu-baseline > gpurun_out/h_bench_1gpu_dpmpp.json 2> gpurun_out/h_bench_1gpu_dpmpp.err; echo "rc $?"; tail -c 300 gpurun_out/h_bench_1gpu_dpmpp.err; head -c 400 gpurun_out/h_bench_1gpu_dpmpp.json; echo
timeout 300 python -m pytest tests/test_flux_gpu.py -m gpu -q -p no:cacheprovider -k "golden" 2>&1 | tail -3
