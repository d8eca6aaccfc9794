// Microbenchmark: TMEM -> register bandwidth (tcgen05.ld 32x32b.x32) per SM for 4 and 8 reader warps.
#include <cstdio>
#include <cuda_runtime.h>
#include "../../stable-diffusion-webui-forge_b200/csrc/common.cuh"
using namespace b200;

__global__ void ldtm_kernel(int iters, int warps_active, long long* cycles, unsigned* sink) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) { tmem_alloc(smem_u32(&slot), 512); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t base = slot;
  unsigned acc = 0;
  __syncthreads();
  long long t0 = clock64();
  if (warp < warps_active) {
    const uint32_t addr = base + ((uint32_t)((warp & 3) * 32) << 16);
    for (int it = 0; it < iters; ++it) {
      uint32_t v[32];
#pragma unroll
      for (int c = 0; c < 128; c += 32) {
        tmem_ld_32x32(addr + c + ((warp >> 2) * 128), v);
        tmem_ld_wait();
        acc += v[0] + v[31];
      }
    }
  }
  __syncthreads();
  long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  __syncthreads();
  if (warp == 0) tmem_dealloc(base, 512);
}

int main() {
  long long* cyc; unsigned* sink;
  cudaMalloc(&cyc, 148 * 8); cudaMalloc(&sink, 148 * 256 * 4);
  for (int w : {1, 4, 8}) {
    const int iters = 2000;
    ldtm_kernel<<<148, 256>>>(iters, w, cyc, sink);
    cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double bytes = (double)iters * 4 * 32 * 32 * 4 * w;   // per SM: iters * 4 chunks * (32 lanes x 32 cols x 4 B) * warps
    printf("warps=%d  cycles=%lld  bytes/clk/SM=%.1f  err=%s\n", w, h[0], bytes / (double)h[0], cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
