#!/bin/bash
# A/B libraries for the attention kernels' exp-phase options (share of polynomial exp2, MUFU turn-taking); measured in
# one GPU call with B200FORGE_LIB=<variant>.  Usage: scripts/build_attn_variants.sh
set -e
cd "$(dirname "$0")/../stable-diffusion-webui-forge_b200/csrc"
make -j8 > /dev/null
mkdir -p build/var ../variants
build() {  # name, extra flags
  local name=$1; shift
  for f in attention64 attention128; do
    nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC --expt-relaxed-constexpr "$@" -c $f.cu -o build/var/$f.$name.o &
  done
  wait
  nvcc -shared -gencode arch=compute_100a,code=sm_100a -o ../variants/lib_$name.so build/host_util.o build/gemm.o build/attention.o \
       build/var/attention64.$name.o build/var/attention128.$name.o build/elementwise.o build/sampler.o build/flux.o -cudart static
}
build m00 -DB200_ATTN_POLY_MASK=0x00
build m10 -DB200_ATTN_POLY_MASK=0x10
build m12 -DB200_ATTN_POLY_MASK=0x12
build m52 -DB200_ATTN_POLY_MASK=0x52
build m00nt -DB200_ATTN_POLY_MASK=0x00 -DB200_ATTN_NO_TURNS
build m12nt -DB200_ATTN_POLY_MASK=0x12 -DB200_ATTN_NO_TURNS
ls -la ../variants
