// Microbenchmark of the softmax exp-phase instruction mix on one SM sub-partition: how many clocks per exponential does ONE
// warp (and 2 / 3 / 4 warps on the same scheduler) need for
//   mode 0: MUFU.EX2 only                         mode 1: + scale-and-subtract FFMA            mode 2: + row-sum FADD
//   mode 3: + F2FP pack                            mode 4: packed FFMA2 / FADD2 variant of 3    mode 5: mode 4 + st.shared.v4 per 8
// 128 independent exponentials per iteration per thread, operands in registers (like the attention kernel's S row).
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o scripts/micro/exp_mix_bench scripts/micro/exp_mix_bench.cu
#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ float ex2a(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ u64 pk2(float a, float b) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void upk2(u64 v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 r; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }

template <int MODE>
__global__ void __launch_bounds__(512, 1) k(float* out, long long* clk, int iters, float sl2, float nm) {
  __shared__ uint4 stg[512 * 2];
  float v[128];
#pragma unroll
  for (int i = 0; i < 128; ++i) v[i] = (float)(threadIdx.x + i) * 1e-3f;
  float acc = 0.f, acc1 = 0.f;
  u64 a2 = pk2(0.f, 0.f), b2 = pk2(0.f, 0.f);
  unsigned hs = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int c = 0; c < 128; c += 8) {
      float pe[8];
      if (MODE <= 3) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float x = v[c + i];
          if (MODE >= 1) x = fmaf(x, sl2, nm);
          pe[i] = ex2a(x);
        }
        if (MODE >= 2) {
          acc += (pe[0] + pe[2]) + (pe[4] + pe[6]);
          acc1 += (pe[1] + pe[3]) + (pe[5] + pe[7]);
        }
      } else {
        const u64 s2 = pk2(sl2, sl2), n2 = pk2(nm, nm);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float xa, xb;
          upk2(fma2(pk2(v[c + 2 * q], v[c + 2 * q + 1]), s2, n2), xa, xb);
          pe[2 * q] = ex2a(xa);
          pe[2 * q + 1] = ex2a(xb);
        }
        a2 = add2(a2, pk2(pe[0], pe[1]));
        b2 = add2(b2, pk2(pe[2], pe[3]));
        a2 = add2(a2, pk2(pe[4], pe[5]));
        b2 = add2(b2, pk2(pe[6], pe[7]));
      }
      if (MODE >= 3) {
        __half2 h0 = __floats2half2_rn(pe[0], pe[1]), h1 = __floats2half2_rn(pe[2], pe[3]);
        __half2 h2 = __floats2half2_rn(pe[4], pe[5]), h3 = __floats2half2_rn(pe[6], pe[7]);
        uint4 o = make_uint4(*(unsigned*)&h0, *(unsigned*)&h1, *(unsigned*)&h2, *(unsigned*)&h3);
        if (MODE >= 5) stg[threadIdx.x * 2 + ((c >> 3) & 1)] = o;
        else hs ^= o.x ^ o.y ^ o.z ^ o.w;
      } else if (MODE < 2) {
#pragma unroll
        for (int i = 0; i < 8; ++i) acc += pe[i] * 1e-30f;  // keep the results alive (cheap, but not free: see mode 2 vs 0)
      }
      v[c] += 1e-6f * (float)it;  // loop-carried change so the iterations are not hoisted
    }
  }
  const long long t1 = clock64();
  float a, b, cc, d;
  upk2(a2, a, b);
  upk2(b2, cc, d);
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc + acc1 + a + b + cc + d + (float)hs + (MODE >= 5 ? (float)stg[threadIdx.x].x : 0.f);
  if (threadIdx.x % 32 == 0) clk[blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32] = t1 - t0;
}

template <int MODE>
static void run(const char* name) {
  float* out;
  long long* clk;
  cudaMalloc(&out, 512 * 4);
  cudaMalloc(&clk, 16 * 8);
  const int iters = 200;
  for (int warps_per_smsp = 1; warps_per_smsp <= 4; ++warps_per_smsp) {
    const int threads = warps_per_smsp * 4 * 32;  // warp w runs on scheduler w % 4
    k<MODE><<<1, threads>>>(out, clk, iters, 0.125f, -3.f);
    k<MODE><<<1, threads>>>(out, clk, iters, 0.125f, -3.f);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[16];
    cudaMemcpy(h, clk, sizeof(long long) * warps_per_smsp * 4, cudaMemcpyDeviceToHost);
    long long mx = 0;
    for (int i = 0; i < warps_per_smsp * 4; ++i) mx = h[i] > mx ? h[i] : mx;
    const double per = (double)mx / ((double)iters * 128.0);
    printf("%-34s %d warp(s)/scheduler: %6.2f clk per exponential per warp, %5.2f exp/clk/SM  %s\n", name, warps_per_smsp, per,
           warps_per_smsp * 4 * 32.0 / per, e == cudaSuccess ? "" : cudaGetErrorString(e));
  }
}

int main() {
  run<0>("MUFU.EX2 only");
  run<1>("FFMA + MUFU");
  run<2>("FFMA + MUFU + FADD");
  run<3>("FFMA + MUFU + FADD + F2FP");
  run<4>("FFMA2 + MUFU + FADD2 + F2FP");
  run<5>("FFMA2 + MUFU + FADD2 + F2FP + STS");
  return 0;
}
