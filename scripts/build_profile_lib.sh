#!/bin/bash
# Instrumented builds of the Dh=64 attention kernel (clock64 phase profile, printed by the kernel) next to the product library.
# Usage: scripts/build_profile_lib.sh [name [extra nvcc flags...]]   ->  stable-diffusion-webui-forge_b200/variants/libprof_<name>.so
set -e
cd "$(dirname "$0")/../stable-diffusion-webui-forge_b200/csrc"
make -j8 > /dev/null
name=${1:-default}
shift || true
mkdir -p build/prof ../variants
nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC --expt-relaxed-constexpr \
     -DB200_ATTN_PROFILE "$@" -c attention64.cu -o build/prof/attention64.$name.o
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o ../variants/libprof_$name.so build/host_util.o build/gemm.o build/attention.o \
     build/prof/attention64.$name.o build/attention64s.o build/attention128.o build/elementwise.o build/sampler.o build/flux.o -cudart static
echo built ../variants/libprof_$name.so
