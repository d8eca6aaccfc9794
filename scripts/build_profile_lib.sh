#!/bin/bash
# Instrumented build of the Dh=64 attention kernel (clock64 phase profile, printed by the kernel) next to the product library.
set -e
cd "$(dirname "$0")/../stable-diffusion-webui-forge_b200/csrc"
make -j8 > /dev/null
mkdir -p build/prof
nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC --expt-relaxed-constexpr \
     -DB200_ATTN_PROFILE -c attention64.cu -o build/prof/attention64.o
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o ../libb200forge_prof.so build/host_util.o build/gemm.o build/attention.o \
     build/prof/attention64.o build/attention128.o build/elementwise.o build/sampler.o build/flux.o -cudart static
echo built ../libb200forge_prof.so
