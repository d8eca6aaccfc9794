#!/bin/bash
# A/B libraries for the Dh = 64 attention kernel's exp-phase options (share of polynomial exp2 as element PAIRS, MUFU
# turn-taking); measured in one GPU call with B200FORGE_LIB=<variant> (the issue structure is chosen at run time by
# B200_ATTN64_VER).  Usage: scripts/build_attn_variants.sh
set -e
cd "$(dirname "$0")/../stable-diffusion-webui-forge_b200/csrc"
make -j8 > /dev/null
mkdir -p build/var ../variants
build() {  # name, extra flags
  local name=$1; shift
  nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC --expt-relaxed-constexpr "$@" -c attention64.cu -o build/var/attention64.$name.o
  nvcc -shared -gencode arch=compute_100a,code=sm_100a -o ../variants/lib_$name.so build/host_util.o build/gemm.o build/attention.o \
       build/var/attention64.$name.o build/attention64s.o build/attention128.o build/elementwise.o build/sampler.o build/flux.o -cudart static
}
build p0 -DB200_ATTN_POLY_PAIRS=0x0 &
build p4 -DB200_ATTN_POLY_PAIRS=0x4 &
build p6 -DB200_ATTN_POLY_PAIRS=0x6 &
build p4nt -DB200_ATTN_POLY_PAIRS=0x4 -DB200_ATTN_NO_TURNS &
build p0nt -DB200_ATTN_POLY_PAIRS=0x0 -DB200_ATTN_NO_TURNS &
build p6nt -DB200_ATTN_POLY_PAIRS=0x6 -DB200_ATTN_NO_TURNS &
wait
ls -la ../variants
