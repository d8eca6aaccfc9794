// b200forge — head-dim-64 attention forward, the SDXL / SD2.x shape (every SDXL attention has Dh = 64).
//
// At Dh = 64 the kernel is bound by the exponentials, not the MMAs (128x128 exps per tile-block on a 16/clk MUFU
// = 1024 cycles vs 512 MMA cycles), so the design goal is to keep the MUFU pipe busy all the time:
//   * one CTA per SM works on TWO 128-query tiles; two softmax warpgroups (4 warps each, thread = query row)
//     ping-pong, so one group's exponentials overlap the other group's QK^T / PV MMAs;
//   * O stays in TMEM and accumulates across key blocks (tcgen05.mma accumulate) — no per-block read-back.
//     The softmax reference max is only advanced when a row's block max exceeds it by more than 2^8
//     ("lazy rescale"): then, and only then, the warp rescales its O rows in TMEM (tcgen05.ld/st);
//   * a softmax thread pulls its whole S row (128 fp32) into registers with one TMEM round trip and releases S
//     immediately, so QK_{j+1} runs underneath the exponentials of block j; PV_j follows when P_j is staged.
//     Barriers per tile: S-full, S-consumed, P-full, PV-done (+ the K/V ring).  The softmax warpgroups take
//     216 registers each via setmaxnreg, the control warps drop to 56.
// TMEM columns: S0 [0,128)  S1 [128,256)  O0 [256,320)  O1 [320,384).
// smem: Q 2x16 KB | K/V ring 4x16 KB | P 2x32 KB (reused as the output staging tile at the end).
#include "common.cuh"
#include "host_util.h"
#include <stdlib.h>

namespace b200 {

struct Attn64Params {
  int B, H, Lq, Lk;
  int BKV, n_kv, q_tiles;  // q_tiles: 256-query CTA tiles
  float scale_log2;
  void* O;
  long long o_stride_b, o_stride_l;
  uint32_t idesc_qk, idesc_pv;
};

__device__ __forceinline__ float ex2a(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ void tmem_st_32x32(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
        "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]),
        "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]),
        "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// Phase profile of the softmax loop (compile with -DB200_ATTN_PROFILE): lane 0 of the first softmax warp of each tile in
// CTA 0 accumulates clock64 deltas per phase and prints the per-key-block averages at the end.
#ifdef B200_ATTN_PROFILE
#define PROF_DECL long long pf_t = clock64(), pf_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0}
#define PROF_MARK(i) do { const long long pf_n = clock64(); pf_acc[i] += pf_n - pf_t; pf_t = pf_n; } while (0)
#else
#define PROF_DECL
#define PROF_MARK(i)
#endif

// ---- packed fp32 pairs (sm_100: FFMA2 / FADD2 process two floats of an even/odd register pair per instruction).  The exp
// phase is bound by the FP32 pipe as much as by the MUFU (clock64 profile: 16 clk per exponential with one scalar FFMA + one
// scalar FADD per element; a 32-lane FP32 instruction occupies the pipe for 2 clk), so the scale-and-subtract and the row sum
// run packed: half the FP32-pipe instructions per element.
static constexpr int kTile = 128 * 128;  // bytes of one 128-row x 64-half tile
static constexpr int kRingSlots = 4;
static constexpr float kRescaleThreshold = 8.0f;  // log2(256)
#ifndef B200_ATTN_POLY_MASK
#define B200_ATTN_POLY_MASK 0x10
#endif
static constexpr unsigned kPolyMask = B200_ATTN_POLY_MASK;  // VER 0: elements (i mod 8) whose exp2 runs on the FMA pipe
#ifndef B200_ATTN_POLY_PAIRS
#define B200_ATTN_POLY_PAIRS 0x0
#endif
static constexpr unsigned kPolyPairs = B200_ATTN_POLY_PAIRS;  // VER 1: element PAIRS (of the 4 per 8 elements) on the FMA pipe

// VER 0: the round-1 kernel — ONE MMA-issue warp walks [QK t0, QK t1, PV t0, PV t1] in order; scalar FFMA / FADD softmax.
// VER 1: one MMA-issue warp PER TILE (warps 1 and 3): the two tiles' softmax warpgroups run half a period apart, and a single
//        in-order issuer made each tile's QK wait for the other tile's S to be consumed (clock64 profile: 440 clk per key
//        block waiting for S); K/V ring slots are released by both issuers (barrier count 2).  Softmax arithmetic packed
//        (FFMA2 / FADD2), row max over four independent chains, polynomial exp2 on whole pairs.
// Both issue from a converged warp with one elected lane (descriptors stay in uniform registers).
template <bool BF16, int VER>
__global__ void __launch_bounds__(384, 1)
attn64_kernel(const __grid_constant__ CUtensorMap mapQ, const __grid_constant__ CUtensorMap mapK,
              const __grid_constant__ CUtensorMap mapV, const Attn64Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t q_smem = base;                       // 2 tiles
  const uint32_t ring_smem = base + 2 * kTile;        // 4 tiles
  const uint32_t p_smem = ring_smem + kRingSlots * kTile;  // 2 x (2 atoms)
  const uint32_t bar_base = p_smem + 4 * kTile;
  const uint32_t q_full = bar_base;
  auto ring_full = [&](int i) { return bar_base + 8u * (1 + i); };
  auto ring_empty = [&](int i) { return bar_base + 8u * (1 + kRingSlots + i); };
  auto s_full = [&](int t) { return bar_base + 8u * (1 + 2 * kRingSlots + t); };
  auto p_full = [&](int t) { return bar_base + 8u * (3 + 2 * kRingSlots + t); };
  auto pv_done = [&](int t) { return bar_base + 8u * (5 + 2 * kRingSlots + t); };
  auto s_cons = [&](int t) { return bar_base + 8u * (7 + 2 * kRingSlots + t); };
  const uint32_t tmem_slot = bar_base + 8u * (9 + 2 * kRingSlots);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int qt = blockIdx.x % p.q_tiles;
  const int h = (blockIdx.x / p.q_tiles) % p.H;
  const int b = blockIdx.x / (p.q_tiles * p.H);
  const int q0 = qt * 256;
  const int BKV = p.BKV;
  const int n_kv = p.n_kv;

  pdl_launch_dependents();
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&mapQ);
    tma_prefetch_desc(&mapK);
    tma_prefetch_desc(&mapV);
    mbar_init(q_full, 1);
    for (int i = 0; i < kRingSlots; ++i) {
      mbar_init(ring_full(i), 1);
      mbar_init(ring_empty(i), VER == 1 ? 2 : 1);  // VER 1: one release per issuing warp
    }
    for (int t = 0; t < 2; ++t) {
      mbar_init(s_full(t), 1);
      mbar_init(p_full(t), 128);
      mbar_init(pv_done(t), 1);
      mbar_init(s_cons(t), 128);
    }
    mbar_fence_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  pdl_wait();  // set-up done; q / k / v are the predecessor's output

  if (warp < 4) {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 72;");  // 4 x (168 - 72) released = 8 x (216 - 168) taken by the softmax warps
  if (warp == 0 && lane == 0) {
    // ------------------------------------------------------------------ TMA producer
    mbar_expect_tx(q_full, 2u * kTile);
    tma_load_3d(q_smem, &mapQ, q_full, h * 64, q0, b);
    tma_load_3d(q_smem + kTile, &mapQ, q_full, h * 64, q0 + 128, b);
    const uint32_t kv_bytes = (uint32_t)BKV * 128u;
    for (int idx = 0; idx < 2 * n_kv; ++idx) {  // even: K_{idx/2}, odd: V_{idx/2}
      const int slot = idx % kRingSlots;
      const uint32_t phase = (uint32_t)(idx / kRingSlots) & 1u;
      mbar_wait(ring_empty(slot), phase ^ 1u);
      mbar_expect_tx(ring_full(slot), kv_bytes);
      tma_load_3d(ring_smem + slot * kTile, (idx & 1) ? &mapV : &mapK, ring_full(slot), h * 64, (idx >> 1) * BKV, b);
    }
  } else if (VER == 0 && warp == 1) {
    // ------------------------------------------------------------------ MMA issuer: the warp stays converged and one
    // elected lane issues, so every descriptor below lives in uniform registers (a lane-0 branch costs ~75 clk per
    // tcgen05.mma in register-to-uniform moves; at N = 64 an MMA is only 32 clk of tensor time)
    const uint64_t qdesc0 = make_smem_desc_sw128(q_smem, 0, 1024);
    const uint64_t kdesc0 = make_smem_desc_sw128(ring_smem, 0, 1024);
    const uint64_t vdesc0 = make_smem_desc_sw128(ring_smem, kTile, 1024);  // MN-major V
    const uint64_t pdesc0 = make_smem_desc_sw128(p_smem, 0, 1024);
    const uint32_t idesc_qk = p.idesc_qk, idesc_pv = p.idesc_pv;
    auto wait_full = [&](int idx) {
      mbar_wait(ring_full(idx % kRingSlots), (uint32_t)(idx / kRingSlots) & 1u);
      tc_fence_after();
    };
    auto issue_qk = [&](int idx, int t) {  // S_t = Q_t K^T
      const uint64_t kd = kdesc0 + (uint64_t)((idx % kRingSlots) * (kTile >> 4));
      const uint64_t qd = qdesc0 + (uint64_t)(t * (kTile >> 4));
      const uint32_t s_tmem = tmem_base + (uint32_t)t * 128u;
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_f16(s_tmem, qd + (uint64_t)(k * 2), kd + (uint64_t)(k * 2), idesc_qk, k != 0 ? 1u : 0u);
        umma_commit(s_full(t));
      }
      __syncwarp();
    };
    mbar_wait(q_full, 0);
    wait_full(0);
    issue_qk(0, 0);
    issue_qk(0, 1);
    if (elect_one()) umma_commit(ring_empty(0));
    __syncwarp();
    const int ksteps = BKV >> 4;
#pragma unroll 1
    for (int j = 0; j < n_kv; ++j) {
      const int vidx = 2 * j + 1, kidx = 2 * j + 2;
      // QK_{j+1} as soon as the softmax threads have pulled S_j into registers (runs under their exps)
      if (j + 1 < n_kv) {
        wait_full(kidx);
        for (int t = 0; t < 2; ++t) {
          mbar_wait(s_cons(t), (uint32_t)j & 1u);
          tc_fence_after();
          issue_qk(kidx, t);
        }
        if (elect_one()) umma_commit(ring_empty(kidx % kRingSlots));
        __syncwarp();
      }
      wait_full(vidx);
      const uint64_t vd = vdesc0 + (uint64_t)((vidx % kRingSlots) * (kTile >> 4));
      for (int t = 0; t < 2; ++t) {
        mbar_wait(p_full(t), (uint32_t)j & 1u);
        tc_fence_after();
        const uint64_t pd = pdesc0 + (uint64_t)(t * 2 * (kTile >> 4));
        const uint32_t o_tmem = tmem_base + 256u + (uint32_t)t * 64u;
        const uint32_t acc0 = j != 0 ? 1u : 0u;
        if (elect_one()) {
          if (ksteps == 8) {
#pragma unroll
            for (int kk = 0; kk < 8; ++kk)
              umma_f16(o_tmem, pd + (uint64_t)((kk >> 2) * (kTile >> 4) + (kk & 3) * 2), vd + (uint64_t)(kk * (2048 >> 4)),
                       idesc_pv, kk ? 1u : acc0);
          } else {
            for (int kk = 0; kk < ksteps; ++kk)
              umma_f16(o_tmem, pd + (uint64_t)((kk >> 2) * (kTile >> 4) + (kk & 3) * 2), vd + (uint64_t)(kk * (2048 >> 4)),
                       idesc_pv, kk ? 1u : acc0);
          }
          umma_commit(pv_done(t));
        }
        __syncwarp();
      }
      if (elect_one()) umma_commit(ring_empty(vidx % kRingSlots));
      __syncwarp();
    }
  } else if (VER == 1 && (warp == 1 || warp == 3)) {
    // ------------------------------------------------------------------ MMA issuer of tile t (warp 1: t = 0, warp 3: t = 1)
    const int t = warp >> 1;
    const uint64_t qdesc = make_smem_desc_sw128(q_smem, 0, 1024) + (uint64_t)(t * (kTile >> 4));
    const uint64_t kdesc0 = make_smem_desc_sw128(ring_smem, 0, 1024);
    const uint64_t vdesc0 = make_smem_desc_sw128(ring_smem, kTile, 1024);  // MN-major V
    const uint64_t pdesc = make_smem_desc_sw128(p_smem, 0, 1024) + (uint64_t)(t * 2 * (kTile >> 4));
    const uint32_t idesc_qk = p.idesc_qk, idesc_pv = p.idesc_pv;
    const uint32_t s_tmem = tmem_base + (uint32_t)t * 128u;
    const uint32_t o_tmem = tmem_base + 256u + (uint32_t)t * 64u;
    auto wait_full = [&](int idx) {
      mbar_wait(ring_full(idx % kRingSlots), (uint32_t)(idx / kRingSlots) & 1u);
      tc_fence_after();
    };
    auto issue_qk = [&](int idx) {  // S_t = Q_t K^T, then this warp's release of the K slot
      const uint64_t kd = kdesc0 + (uint64_t)((idx % kRingSlots) * (kTile >> 4));
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_f16(s_tmem, qdesc + (uint64_t)(k * 2), kd + (uint64_t)(k * 2), idesc_qk, k != 0 ? 1u : 0u);
        umma_commit(s_full(t));
        umma_commit(ring_empty(idx % kRingSlots));
      }
      __syncwarp();
    };
    mbar_wait(q_full, 0);
    wait_full(0);
    issue_qk(0);
    const int ksteps = BKV >> 4;
#pragma unroll 1
    for (int j = 0; j < n_kv; ++j) {
      const int vidx = 2 * j + 1, kidx = 2 * j + 2;
      // QK_{j+1} as soon as this tile's softmax threads have pulled S_j into registers (runs under their exps)
      if (j + 1 < n_kv) {
        wait_full(kidx);
        mbar_wait(s_cons(t), (uint32_t)j & 1u);
        tc_fence_after();
        issue_qk(kidx);
      }
      wait_full(vidx);
      const uint64_t vd = vdesc0 + (uint64_t)((vidx % kRingSlots) * (kTile >> 4));
      mbar_wait(p_full(t), (uint32_t)j & 1u);
      tc_fence_after();
      const uint32_t acc0 = j != 0 ? 1u : 0u;
      if (elect_one()) {
        if (ksteps == 8) {
#pragma unroll
          for (int kk = 0; kk < 8; ++kk)
            umma_f16(o_tmem, pdesc + (uint64_t)((kk >> 2) * (kTile >> 4) + (kk & 3) * 2), vd + (uint64_t)(kk * (2048 >> 4)),
                     idesc_pv, kk ? 1u : acc0);
        } else {
          for (int kk = 0; kk < ksteps; ++kk)
            umma_f16(o_tmem, pdesc + (uint64_t)((kk >> 2) * (kTile >> 4) + (kk & 3) * 2), vd + (uint64_t)(kk * (2048 >> 4)),
                     idesc_pv, kk ? 1u : acc0);
        }
        umma_commit(pv_done(t));
        umma_commit(ring_empty(vidx % kRingSlots));
      }
      __syncwarp();
    }
  }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 216;");
    // ------------------------------------------------------------------ softmax warpgroups (t = tile)
    const int t = (warp - 4) >> 2;
    const int quad = warp & 3;
    const int r = quad * 32 + lane;  // row inside the tile
    const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
    const uint32_t s_addr = tmem_base + (uint32_t)t * 128u + lane_addr;
    const uint32_t o_addr = tmem_base + 256u + (uint32_t)t * 64u + lane_addr;
    const uint32_t p_tile = p_smem + (uint32_t)t * 2 * kTile;
    const uint32_t p_row = p_tile + (uint32_t)r * 128u;
    const uint32_t sw = (uint32_t)(r & 7);
    const float sl2 = p.scale_log2;
    float m_ref = -INFINITY, l_run = 0.f;
#ifndef B200_ATTN_NO_TURNS
    if (t == 1) named_bar_arrive(2, 256);  // warpgroup 0 takes the first turn on the MUFU
#endif
    PROF_DECL;

    for (int j = 0; j < n_kv; ++j) {
      int nvalid = p.Lk - j * BKV;
      if (nvalid > BKV) nvalid = BKV;
      mbar_wait(s_full(t), (uint32_t)j & 1u);
      tc_fence_after();
      PROF_MARK(0);  // waiting for S_j
      // the whole S row -> registers in one TMEM round trip, then hand S back to the tensor core
      uint32_t v[128];
      tmem_ld_32x32(s_addr + 0u, *reinterpret_cast<uint32_t(*)[32]>(&v[0]));
      tmem_ld_32x32(s_addr + 32u, *reinterpret_cast<uint32_t(*)[32]>(&v[32]));
      tmem_ld_32x32(s_addr + 64u, *reinterpret_cast<uint32_t(*)[32]>(&v[64]));
      tmem_ld_32x32(s_addr + 96u, *reinterpret_cast<uint32_t(*)[32]>(&v[96]));
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(s_cons(t));
      PROF_MARK(1);  // TMEM -> registers
      const bool full_blk = nvalid == 128;
      float mx = -INFINITY;
      if (full_blk) {
        if constexpr (VER == 1) {
          // four independent chains: one chain of 64 dependent 3-input maxima cost ~375 clk of pure latency per key block
          float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
#pragma unroll
          for (int i = 0; i < 128; i += 8) {
            m0 = fmaxf(m0, fmaxf(__uint_as_float(v[i + 0]), __uint_as_float(v[i + 1])));
            m1 = fmaxf(m1, fmaxf(__uint_as_float(v[i + 2]), __uint_as_float(v[i + 3])));
            m2 = fmaxf(m2, fmaxf(__uint_as_float(v[i + 4]), __uint_as_float(v[i + 5])));
            m3 = fmaxf(m3, fmaxf(__uint_as_float(v[i + 6]), __uint_as_float(v[i + 7])));
          }
          mx = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
        } else {
#pragma unroll
          for (int i = 0; i < 128; i += 2) mx = fmaxf(mx, fmaxf(__uint_as_float(v[i]), __uint_as_float(v[i + 1])));
        }
      } else {
#pragma unroll
        for (int i = 0; i < 128; ++i)
          if (i < nvalid) mx = fmaxf(mx, __uint_as_float(v[i]));
      }
      const float m_blk = mx * sl2;
      PROF_MARK(2);  // row max
      // PV_{j-1} must have retired before O is rescaled and before P is overwritten
      if (j > 0) {
        mbar_wait(pv_done(t), (uint32_t)(j - 1) & 1u);
        tc_fence_after();
      }
      PROF_MARK(3);  // waiting for PV_{j-1}
      if (j == 0) {
        m_ref = m_blk;
      } else {
        const bool need = m_blk > m_ref + kRescaleThreshold;
        if (__any_sync(0xffffffffu, need)) {
          const float alpha = need ? ex2a(m_ref - m_blk) : 1.0f;
#pragma unroll
          for (int c = 0; c < 64; c += 32) {
            uint32_t w[32];
            tmem_ld_32x32(o_addr + (uint32_t)c, w);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) w[i] = __float_as_uint(__uint_as_float(w[i]) * alpha);
            tmem_st_32x32(o_addr + (uint32_t)c, w);
          }
          tmem_st_wait();
          l_run *= alpha;
          if (need) m_ref = m_blk;
        }
      }
      // p = exp2(s*scale - m_ref); row sum; stage P (128B-swizzled K-major A operand).
      // The two warpgroups take turns on this MUFU-bound phase (named-barrier hand-off) so that one group's
      // exponentials overlap the other group's TMEM loads / max / stores instead of both fighting for the MUFU.
      PROF_MARK(4);  // rescale
#ifndef B200_ATTN_NO_TURNS
      named_bar_sync(2 + t, 256);
#endif
      PROF_MARK(5);  // waiting for the turn
      float rs0 = 0.f, rs1 = 0.f;
      const float nm = -m_ref;
      if (VER == 1 && full_blk) {
        // packed arithmetic: x = s * scale_log2 - m_ref and the row sum handle two elements per FP32-pipe instruction
        const f32x2_t sl2p = pk2(sl2, sl2), nmp = pk2(nm, nm);
        f32x2_t acc0 = pk2(0.f, 0.f), acc1 = pk2(0.f, 0.f);
#pragma unroll
        for (int c = 0; c < 128; c += 8) {
          float pe[8];
#pragma unroll
          for (int q2 = 0; q2 < 4; ++q2) {
            const f32x2_t x = fma2(pk2(__uint_as_float(v[c + 2 * q2]), __uint_as_float(v[c + 2 * q2 + 1])), sl2p, nmp);
            if ((kPolyPairs >> q2) & 1) {
              exp2_poly3_x2(x, pe[2 * q2], pe[2 * q2 + 1]);
            } else {
              float xa, xb;
              upk2(x, xa, xb);
              pe[2 * q2] = ex2a(xa);
              pe[2 * q2 + 1] = ex2a(xb);
            }
          }
          acc0 = add2(acc0, pk2(pe[0], pe[1]));
          acc1 = add2(acc1, pk2(pe[2], pe[3]));
          acc0 = add2(acc0, pk2(pe[4], pe[5]));
          acc1 = add2(acc1, pk2(pe[6], pe[7]));
          const uint32_t addr = p_row + (uint32_t)(c >> 6) * kTile + (((((uint32_t)c & 63u) >> 3) ^ sw) << 4);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(pack2<BF16>(pe[0], pe[1])),
                       "r"(pack2<BF16>(pe[2], pe[3])), "r"(pack2<BF16>(pe[4], pe[5])), "r"(pack2<BF16>(pe[6], pe[7])));
        }
        float a0, a1, b0, b1;
        upk2(acc0, a0, a1);
        upk2(acc1, b0, b1);
        rs0 = a0 + b0;
        rs1 = a1 + b1;
      } else if (full_blk) {
#pragma unroll
        for (int c = 0; c < 128; c += 8) {
          float pe[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float xs = fmaf(__uint_as_float(v[c + i]), sl2, nm);
            pe[i] = ((kPolyMask >> i) & 1) ? exp2_poly3(xs) : ex2a(xs);  // kPolyMask: which of every 8 exponentials run on the FMA pipe
          }
          rs0 += (pe[0] + pe[2]) + (pe[4] + pe[6]);
          rs1 += (pe[1] + pe[3]) + (pe[5] + pe[7]);
          const uint32_t addr = p_row + (uint32_t)(c >> 6) * kTile + (((((uint32_t)c & 63u) >> 3) ^ sw) << 4);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(pack2<BF16>(pe[0], pe[1])),
                       "r"(pack2<BF16>(pe[2], pe[3])), "r"(pack2<BF16>(pe[4], pe[5])), "r"(pack2<BF16>(pe[6], pe[7])));
        }
      } else {
#pragma unroll
        for (int c = 0; c < 128; c += 8) {
          if (c < BKV) {
            float pe[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) pe[i] = (c + i < nvalid) ? ex2a(fmaf(__uint_as_float(v[c + i]), sl2, nm)) : 0.f;
            rs0 += (pe[0] + pe[2]) + (pe[4] + pe[6]);
            rs1 += (pe[1] + pe[3]) + (pe[5] + pe[7]);
            const uint32_t addr = p_row + (uint32_t)(c >> 6) * kTile + (((((uint32_t)c & 63u) >> 3) ^ sw) << 4);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(pack2<BF16>(pe[0], pe[1])),
                         "r"(pack2<BF16>(pe[2], pe[3])), "r"(pack2<BF16>(pe[4], pe[5])), "r"(pack2<BF16>(pe[6], pe[7])));
          }
        }
      }
      PROF_MARK(6);  // exp phase
#ifndef B200_ATTN_NO_TURNS
      if (!(t == 1 && j == n_kv - 1)) named_bar_arrive(3 - t, 256);
#endif
      l_run += rs0 + rs1;
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(p_full(t));
      PROF_MARK(7);  // hand-off
    }
#ifdef B200_ATTN_PROFILE
    if (blockIdx.x == 0 && quad == 0 && lane == 0)
      printf("attn64 profile tile %d: per key block clk  waitS %lld  ldtm %lld  max %lld  waitPV %lld  rescale %lld  waitTurn %lld  exp %lld  handoff %lld  (n_kv %d)\n",
             t, pf_acc[0] / n_kv, pf_acc[1] / n_kv, pf_acc[2] / n_kv, pf_acc[3] / n_kv, pf_acc[4] / n_kv, pf_acc[5] / n_kv,
             pf_acc[6] / n_kv, pf_acc[7] / n_kv, n_kv);
#endif

    // ---- output: O_t / l -> fp16 -> warp-private staging (the P tile is free now) -> coalesced stores
    mbar_wait(pv_done(t), (uint32_t)(n_kv - 1) & 1u);
    tc_fence_after();
    const float inv = 1.0f / l_run;
    const uint32_t stg = p_tile + (uint32_t)quad * 4096u;  // 32 rows x 128 B per warp
    __syncwarp();
#pragma unroll
    for (int c = 0; c < 64; c += 32) {
      uint32_t v[32];
      tmem_ld_32x32(o_addr + (uint32_t)c, v);
      tmem_ld_wait();
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const uint32_t o0 = pack2<BF16>(__uint_as_float(v[g * 8 + 0]) * inv, __uint_as_float(v[g * 8 + 1]) * inv);
        const uint32_t o1 = pack2<BF16>(__uint_as_float(v[g * 8 + 2]) * inv, __uint_as_float(v[g * 8 + 3]) * inv);
        const uint32_t o2 = pack2<BF16>(__uint_as_float(v[g * 8 + 4]) * inv, __uint_as_float(v[g * 8 + 5]) * inv);
        const uint32_t o3 = pack2<BF16>(__uint_as_float(v[g * 8 + 6]) * inv, __uint_as_float(v[g * 8 + 7]) * inv);
        const uint32_t chunk = (uint32_t)(c >> 3) + (uint32_t)g;  // 16B chunk index inside the 128B row
        const uint32_t addr = stg + (uint32_t)lane * 128u + ((chunk ^ (uint32_t)(lane & 7)) << 4);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(o0), "r"(o1), "r"(o2), "r"(o3) : "memory");
      }
    }
    __syncwarp();
    {
      const int piece = lane & 7;
      char* obase = reinterpret_cast<char*>(p.O) + ((size_t)b * p.o_stride_b + (size_t)h * 64) * 2;
#pragma unroll
      for (int jj = 0; jj < 8; ++jj) {
        const int rl = jj * 4 + (lane >> 3);
        const uint32_t addr = stg + (uint32_t)rl * 128u + ((((uint32_t)piece) ^ (uint32_t)(rl & 7)) << 4);
        uint32_t o0, o1, o2, o3;
        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(o0), "=r"(o1), "=r"(o2), "=r"(o3) : "r"(addr));
        const int q = q0 + t * 128 + quad * 32 + rl;
        if (q < p.Lq)
          *reinterpret_cast<uint4*>(obase + ((size_t)q * p.o_stride_l + piece * 8) * 2) = make_uint4(o0, o1, o2, o3);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

template <bool BF16, int VER>
static int launch_attn64(const CUtensorMap& mQ, const CUtensorMap& mK, const CUtensorMap& mV, const Attn64Params& p,
                         cudaStream_t stream) {
  const size_t smem = (size_t)kTile * (2 + kRingSlots + 4) + 1024 + 256;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(attn64_kernel<BF16, VER>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      set_error("attention64: smem attr: %s", cudaGetErrorString(e));
      return B200_ECUDA;
    }
    attr_done = true;
  }
  const int grid = p.q_tiles * p.H * p.B;
  {
    cudaError_t e = launch_pdl(attn64_kernel<BF16, VER>, dim3(grid), dim3(384), smem, stream, 1, mQ, mK, mV, p);
    if (e != cudaSuccess) {
      set_error("attention64: launch failed: %s", cudaGetErrorString(e));
      return B200_ECUDA;
    }
  }
  B200_CHECK_LAUNCH("attention64");
  return B200_OK;
}

int attention64s_dispatch(const void* q, const void* k, const void* v, void* o, const b200_attn_desc* d, cudaStream_t st);

// called from b200_attention (attention.cu) for Dh == 64
int attention64_dispatch(const void* q, const void* k, const void* v, void* o, const b200_attn_desc* d, cudaStream_t st) {
  Attn64Params p;
  memset(&p, 0, sizeof(p));
  p.B = d->B;
  p.H = d->H;
  p.Lq = d->Lq;
  p.Lk = d->Lk;
  p.BKV = d->Lk >= 128 ? 128 : ((d->Lk + 15) / 16) * 16;
  p.n_kv = (d->Lk + p.BKV - 1) / p.BKV;
  p.q_tiles = (d->Lq + 255) / 256;
  p.scale_log2 = d->scale * 1.4426950408889634f;
  p.O = o;
  p.o_stride_b = d->o_stride_b;
  p.o_stride_l = d->o_stride_l;
  const bool bf = d->dtype == B200_BF16;
  p.idesc_qk = make_idesc_f16(128, p.BKV, bf, false, false);
  p.idesc_pv = make_idesc_f16(128, 64, bf, false, true);
  const uint64_t cols = (uint64_t)d->H * 64;
  CUtensorMap mQ, mK, mV;
  auto make3 = [&](CUtensorMap* m, const void* base, int L, long long sl, long long sb, int rows) {
    uint64_t dims[3] = {cols, (uint64_t)L, (uint64_t)d->B};
    uint64_t str[2] = {(uint64_t)sl * 2, (uint64_t)sb * 2};
    uint32_t box[3] = {64, (uint32_t)rows, 1};
    return make_tmap(m, d->dtype, base, 3, dims, str, box);
  };
  int rc = make3(&mQ, q, d->Lq, d->q_stride_l, d->q_stride_b, 128);
  if (rc) return rc;
  rc = make3(&mK, k, d->Lk, d->k_stride_l, d->k_stride_b, p.BKV);
  if (rc) return rc;
  rc = make3(&mV, v, d->Lk, d->v_stride_l, d->v_stride_b, p.BKV);
  if (rc) return rc;
  // B200_ATTN64_VER = 0 | 1 | 2 forces one build (A/B measurements).  Default 2, the small-CTA kernel (attention64s.cu):
  // measured on B200 at B16 H10 L4096 / B16 H20 L1024: VER 0 640 / 510 TF/s, VER 1 692 / 544, VER 2 790 / 652 (cuDNN SDPA 945 / 830)
  static int ver = -1;
  if (ver < 0) {
    const char* e = getenv("B200_ATTN64_VER");
    ver = !e ? 2 : (e[0] == '0' ? 0 : (e[0] == '1' ? 1 : 2));
  }
  if (ver == 2) return attention64s_dispatch(q, k, v, o, d, st);
  if (ver == 1) return bf ? launch_attn64<true, 1>(mQ, mK, mV, p, st) : launch_attn64<false, 1>(mQ, mK, mV, p, st);
  return bf ? launch_attn64<true, 0>(mQ, mK, mV, p, st) : launch_attn64<false, 0>(mQ, mK, mV, p, st);
}

}  // namespace b200
