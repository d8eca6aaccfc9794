// b200forge — shared device helpers for the sm_100a kernels (raw PTX, no CUTLASS dependency).
//
// Everything here is a thin wrapper over one PTX instruction family:
//   mbarrier.*                         producer/consumer pipelines
//   cp.async.bulk.tensor.*             TMA tile loads (2D/3D/4D, zero-fill out of bounds)
//   tcgen05.{alloc,mma,commit,ld,...}  5th-gen tensor cores with TMEM accumulators
// plus the two descriptor encoders (shared-memory matrix descriptor, instruction descriptor).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>

#ifndef B200_WATCHDOG
#define B200_WATCHDOG 1   // spin-wait watchdog: a broken pipeline traps instead of hanging the GPU
#endif

namespace b200 {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---------------------------------------------------------------- programmatic dependent launch
// launch_dependents: the next kernel on the stream may start scheduling its CTAs; wait: block until every prerequisite grid
// has completed and its memory is visible.  Both are no-ops for a kernel launched without the PDL attribute.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// ---------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
#if B200_WATCHDOG
  long long t0 = clock64();
#endif
  while (!mbar_try_wait(bar, parity)) {
#if B200_WATCHDOG
    if (clock64() - t0 > 4000000000LL) {  // ~2 s at 2 GHz: the pipeline is dead
      printf("b200forge watchdog: mbarrier wait timed out (block %d thread %d bar 0x%x parity %u)\n",
             (int)blockIdx.x, (int)threadIdx.x, bar, parity);
      __trap();
    }
#endif
  }
}

// same wait without the diagnostic printf (whose argument marshalling costs registers): for warps that run on a 24-register
// budget after setmaxnreg.dec
__device__ __forceinline__ void mbar_wait_quiet(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
#if B200_WATCHDOG
  const unsigned t0 = (unsigned)clock();
#endif
  while (!mbar_try_wait(bar, parity)) {
#if B200_WATCHDOG
    if ((unsigned)clock() - t0 > 3000000000u) __trap();  // ~1.5 s: the pipeline is dead
#endif
  }
}

// generic-proxy writes (st.shared) -> visible to the async proxy (TMA / tcgen05.mma operand reads)
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ---------------------------------------------------------------- TMA
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

// TMA store: smem tile -> global (coalesced, clipped at the tensor edge), tracked by bulk async-groups
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(src), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void bulk_commit_group() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_group_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void named_bar_arrive(uint32_t id, uint32_t nthreads) {
  asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void named_bar_sync(uint32_t id, uint32_t nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ---------------------------------------------------------------- tcgen05 / TMEM
__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t ncols) {  // whole warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {  // whole warp
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {  // whole warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem desc] * B[smem desc]; one thread issues for the whole CTA.
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// One lane of a converged warp.  Issuing tcgen05 / TMA instructions from `if (elect_one())` inside warp-uniform control
// flow (instead of an `if (lane == 0)` branch) lets the compiler keep descriptors in uniform registers: measured
// ~75 clk -> a few clk per tcgen05.mma.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}

// mbarrier arrives once every tcgen05 op issued so far by this thread has completed
// (implies tcgen05.fence::before_thread_sync).
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// TMEM -> registers: this warp's 32 lanes x 32 consecutive fp32 columns (lane = accumulator row).
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---------------------------------------------------------------- 2-CTA (cta_group::2) variants
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the barrier at the same smem offset in CTA `cta` of the cluster
__device__ __forceinline__ void mbar_arrive_remote(uint32_t bar, uint32_t cta) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}"
      ::"r"(bar), "r"(cta)
      : "memory");
}
// TMA loads issued by either CTA of a pair; the transaction bytes are credited to CTA 0's barrier
// (peer bit 24 of the shared::cluster address cleared).
__device__ __forceinline__ void tma_load_2d_cg2(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar & 0xFEFFFFFFu), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d_cg2(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1,
                                                int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar & 0xFEFFFFFFu), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_cg2(uint32_t smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish_cg2() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_cg2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// M = 256 MMA across the CTA pair, issued by one thread of CTA 0
__device__ __forceinline__ void umma_f16_cg2(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// commit -> arrive on the barrier at this smem offset in every CTA of `mask`
__device__ __forceinline__ void umma_commit_cg2(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"(mask)
               : "memory");
}

// ---------------------------------------------------------------- descriptors
// Shared-memory matrix descriptor (64-bit), PTX ISA "tcgen05 matrix descriptor":
//   [0,14)  start address >> 4        [16,30) leading-dim byte offset >> 4
//   [32,46) stride-dim byte offset >> 4   [46,48) version = 1 on sm_100
//   [61,64) layout: 0 none, 2 = 128B swizzle, 4 = 64B, 6 = 32B
// K-major, 128B swizzle (rows of 64 half elements = 128 B, 8-row swizzle atom = 1024 B):
//   SBO = 1024 (next 8-row group), LBO unused. Advancing K by 16 elements = +32 B on the start address.
// MN-major, 128B swizzle (tile stored [k][64 mn-elements]): SBO = 1024 (next 8 k-rows),
//   LBO = byte distance between 64-wide mn atoms. Advancing K by 16 = +2048 B.
__device__ __forceinline__ uint64_t make_smem_desc_sw128(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// Instruction descriptor for kind::f16 (fp16/bf16 operands, fp32 accumulate):
//   [4,6) D format (1 = f32)  [7,10) A format (0 f16, 1 bf16)  [10,13) B format
//   [15] A major (0 = K)  [16] B major (0 = K, 1 = MN)  [17,23) N>>3  [24,29) M>>4
__host__ __device__ __forceinline__ uint32_t make_idesc_f16(int M, int N, bool bf16, bool a_mn_major,
                                                            bool b_mn_major) {
  uint32_t d = 0;
  d |= 1u << 4;
  d |= (bf16 ? 1u : 0u) << 7;
  d |= (bf16 ? 1u : 0u) << 10;
  d |= (a_mn_major ? 1u : 0u) << 15;
  d |= (b_mn_major ? 1u : 0u) << 16;
  d |= (uint32_t)(N >> 3) << 17;
  d |= (uint32_t)(M >> 4) << 24;
  return d;
}

// ---------------------------------------------------------------- packed fp32 pairs
// sm_100 FFMA2 / FADD2 / FMUL2 process the two floats of an even/odd register pair per instruction: half the FP32-pipe
// issue slots per element for the epilogue's and the softmax's elementwise arithmetic.
typedef unsigned long long f32x2_t;
__device__ __forceinline__ f32x2_t pk2(float a, float b) {
  f32x2_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void upk2(f32x2_t v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ f32x2_t fma2(f32x2_t a, f32x2_t b, f32x2_t c) {
  f32x2_t r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ f32x2_t add2(f32x2_t a, f32x2_t b) {
  f32x2_t r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2_t mul2(f32x2_t a, f32x2_t b) {
  f32x2_t r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}

// ---------------------------------------------------------------- small math / pack helpers
template <bool BF16>
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  if constexpr (BF16) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
  } else {
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
  }
}
template <bool BF16>
__device__ __forceinline__ float2 unpack2(uint32_t u) {
  if constexpr (BF16) {
    return __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&u));
  } else {
    return __half22float2(*reinterpret_cast<__half2*>(&u));
  }
}
template <bool BF16>
__device__ __forceinline__ float ld1(const void* p, size_t i) {
  if constexpr (BF16) return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p)[i]);
  else return __half2float(reinterpret_cast<const __half*>(p)[i]);
}
template <bool BF16>
__device__ __forceinline__ void st1(void* p, size_t i, float v) {
  if constexpr (BF16) reinterpret_cast<__nv_bfloat16*>(p)[i] = __float2bfloat16_rn(v);
  else reinterpret_cast<__half*>(p)[i] = __float2half_rn(v);
}
__device__ __forceinline__ float silu_f(float x) { return x / (1.0f + __expf(-x)); }
// one MUFU op instead of two (ex2 + rcp): sigmoid(x) = 0.5 + 0.5*tanh(x/2).  Used where the SiLU is applied to
// every element of a large activation and the kernel would otherwise be MUFU- rather than HBM-bound (ncu: XU 59 %).
__device__ __forceinline__ float silu_fast_f(float x) {
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.5f * x));
  return x * fmaf(0.5f, t, 0.5f);
}
__device__ __forceinline__ float gelu_erf_f(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }
// F.gelu (exact / erf form) on TWO values: 0.5 x (1 + erf(x / sqrt 2)) with erf from Abramowitz & Stegun 7.1.26
//   erf(z) = 1 - (a1 t + a2 t^2 + a3 t^3 + a4 t^4 + a5 t^5) exp(-z^2),  t = 1 / (1 + 0.3275911 z),  z >= 0,  |error| <= 1.5e-7
// (far below the fp16 / bf16 output rounding): per element one MUFU.RCP, one MUFU.EX2 and ~5 packed FMAs instead of erff's
// ~25-instruction branchy evaluation — the GEGLU epilogue of the feed-forward GEMM (the largest GEMM of the UNet) ran 11-18 %
// under the same GEMM without an epilogue because of it.
__device__ __forceinline__ f32x2_t gelu_erf_x2(f32x2_t x2) {
  float xa, xb;
  upk2(x2, xa, xb);
  const float za = fabsf(xa) * 0.70710678118654752f, zb = fabsf(xb) * 0.70710678118654752f;
  float ta, tb, ea, eb;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(ta) : "f"(fmaf(0.3275911f, za, 1.0f)));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(tb) : "f"(fmaf(0.3275911f, zb, 1.0f)));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ea) : "f"(za * za * -1.4426950408889634f));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(eb) : "f"(zb * zb * -1.4426950408889634f));
  const f32x2_t t = pk2(ta, tb);
  f32x2_t p = fma2(pk2(1.061405429f, 1.061405429f), t, pk2(-1.453152027f, -1.453152027f));
  p = fma2(p, t, pk2(1.421413741f, 1.421413741f));
  p = fma2(p, t, pk2(-0.284496736f, -0.284496736f));
  p = fma2(p, t, pk2(0.254829592f, 0.254829592f));
  p = mul2(mul2(p, t), pk2(ea, eb));                       // 1 - erf(z)
  float qa, qb;
  upk2(p, qa, qb);
  const float erfa = copysignf(1.0f - qa, xa), erfb = copysignf(1.0f - qb, xb);
  const f32x2_t h = mul2(x2, pk2(0.5f, 0.5f));
  return fma2(h, pk2(erfa, erfb), h);
}

// 2^x on the FMA pipe (Cody-Waite split + degree-3 minimax polynomial on [-0.5, 0.5], max relative error 7.5e-5 — a
// fraction of an fp16 ulp).  Measured with the clock64 phase profile (scripts/attn_phase_profile.py): one warp's exp
// phase of 128 exponentials takes 2040 clk = 16 clk per MUFU.EX2, twice the pipe's 8 clk/instruction peak, so the
// attention kernels move a share of the exponentials here.  A/B on B200 (profiles/experiments/README.md): 1 of 8 is the
// best split (+4.5 % at Dh = 64); larger shares lengthen the other tile's phases by as much as they shorten this one.
__device__ __forceinline__ float exp2_poly3(float x) {
  x = fmaxf(x, -125.0f);
  const float t = x + 12582912.0f;         // 1.5 * 2^23: the integer part of x lands in the low mantissa bits
  const float f = x - (t - 12582912.0f);   // fractional part in [-0.5, 0.5]
  float p = fmaf(0.0551716685f, f, 0.2426111251f);
  p = fmaf(p, f, 0.6932609677f);
  p = fmaf(p, f, 0.9999280572f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));  // p * 2^n through the exponent field
}

// two 2^x on the FP32 pipe: the packed form of exp2_poly3
__device__ __forceinline__ void exp2_poly3_x2(f32x2_t x, float& ea, float& eb) {
  float xa, xb;
  upk2(x, xa, xb);
  const f32x2_t xc = pk2(fmaxf(xa, -125.0f), fmaxf(xb, -125.0f));
  const f32x2_t t = add2(xc, pk2(12582912.0f, 12582912.0f));
  const f32x2_t f = add2(xc, fma2(t, pk2(-1.0f, -1.0f), pk2(12582912.0f, 12582912.0f)));  // x - (t - magic)
  f32x2_t p = fma2(pk2(0.0551716685f, 0.0551716685f), f, pk2(0.2426111251f, 0.2426111251f));
  p = fma2(p, f, pk2(0.6932609677f, 0.6932609677f));
  p = fma2(p, f, pk2(0.9999280572f, 0.9999280572f));
  float pa, pb, ta, tb;
  upk2(p, pa, pb);
  upk2(t, ta, tb);
  ea = __int_as_float(__float_as_int(pa) + (__float_as_int(ta) << 23));
  eb = __int_as_float(__float_as_int(pb) + (__float_as_int(tb) << 23));
}

// nn.GELU(approximate="tanh"): 0.5 x (1 + tanh(sqrt(2/pi) (x + 0.044715 x^3))), MUFU.TANH (rel. error ~2^-11)
__device__ __forceinline__ float gelu_tanh_f(float x) {
  const float u = 0.7978845608028654f * fmaf(0.044715f * x, x * x, x);
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(u));
  const float hx = 0.5f * x;
  return fmaf(hx, t, hx);
}

}  // namespace b200
