#!/bin/bash
# A/B libraries for the Dh = 64 attention kernels' exp-phase options, measured in one GPU call with B200FORGE_LIB=<variant>
# (the kernel organisation is chosen at run time by B200_ATTN64_VER).  Usage: scripts/build_attn_variants.sh
#   p*   attention64.cu  (two tiles / CTA): polynomial-exp2 pair mask, nt = no MUFU turn-taking
#   s*   attention64s.cu (small CTA, 3 / SM): polynomial-exp2 pair mask; spt0 = P through shared memory (round-2 first form)
set -e
cd "$(dirname "$0")/../stable-diffusion-webui-forge_b200/csrc"
make -j8 > /dev/null
mkdir -p build/var ../variants
NV="nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC --expt-relaxed-constexpr"
build() {  # name, source stem, extra flags
  local name=$1 stem=$2; shift 2
  $NV "$@" -c $stem.cu -o build/var/$stem.$name.o
  local a64=build/attention64.o a64s=build/attention64s.o
  [ $stem = attention64 ] && a64=build/var/$stem.$name.o || a64s=build/var/$stem.$name.o
  nvcc -shared -gencode arch=compute_100a,code=sm_100a -o ../variants/lib_$name.so build/host_util.o build/gemm.o build/attention.o \
       $a64 $a64s build/attention128.o build/elementwise.o build/sampler.o build/flux.o -cudart static
}
build s1 attention64s -DB200_ATTN64S_POLY_PAIRS=0x1 &
build s0 attention64s -DB200_ATTN64S_POLY_PAIRS=0x0 &
build s3 attention64s -DB200_ATTN64S_POLY_PAIRS=0x3 &
build spt0 attention64s -DB200_ATTN64S_P_TMEM=0 &
wait
ls -la ../variants
