// b200forge — attention forward, "small CTA" build: ONE 128-query tile per CTA, key blocks of 64, several CTAs per SM.
// Written for head dim 64 (SD1.5 / SDXL: three CTAs per SM, described first); the same pipeline is also built for head dim 128
// (Flux / SD3: two CTAs per SM — see the note above the kernel).
//
// Why a second organisation (measured on B200, profiles/experiments/README.md):
//   * the exponential pipe (MUFU, 16 / clk / SM) bounds Dh = 64 attention; one softmax warp alone reaches ~10 clk per
//     exponential with packed FP32 arithmetic (8 is the pipe's limit), two warps on one scheduler together 96 % of the pipe
//     (scripts/micro/exp_mix_bench.cu) — so the pipe is saturated only while (at least) two warps per scheduler are inside
//     their exp phase;
//   * in the two-tiles-per-CTA kernel (attention64.cu) each softmax warp spends ~1400 clk per key block OUTSIDE the exp phase
//     (waiting for S, TMEM -> registers, row max, waiting for P.V, lazy-rescale check, hand-off) against ~1900 inside, and
//     there are exactly two softmax warps per scheduler: the pipe idles ~45 % of the time; CTA start-up (barrier init, TMEM
//     allocation, first Q / K loads: ~5000 clk) and the output tail are not overlapped with anything either (one CTA per SM):
//     19 % of the run at L = 1024.
// Three resident CTAs give every scheduler three softmax warps from independent pipelines — the non-exp phases, the
// start-up and the tail of one CTA run under the exp phases of the other two — without any cross-tile hand-shaking.
// Per CTA: 64 KB smem (Q 2 x 16 K | K/V ring 4 x 8 K), 128 + 32 TMEM columns (S [0,64) | O [64,128) | P: 64 keys as fp16 pairs),
// 256 threads in two warpgroups (setmaxnreg is a warpgroup-wide instruction: the register-poor and the register-rich roles must not share one):
//   warp 0  TMA producer          warp 1  TMEM owner + MMA issuer (converged warp, elected lane)      warps 2-3  idle
//   warps 4-7  softmax, thread = query row: S row (64 fp32) -> registers in one TMEM round trip, S released at once
//              (QK_{j+1} runs under the exponentials of block j), packed FFMA2 / FADD2 arithmetic, O accumulates in TMEM with
//              lazy rescale, P written straight into TMEM (tcgen05.st) as the A operand of P.V (tcgen05.mma with A in TMEM):
//              no shared-memory round trip for P (it was half of the MMA's operand reads and 8 STS.128 per thread and block
//              on the MIO queue that MUFU shares).  B200_ATTN64S_P_TMEM=0 builds the round-2 first form (P through smem).
#include "common.cuh"
#include "host_util.h"
#include <stdlib.h>

namespace b200 {

struct Attn64sParams {
  int B, H, Lq, Lk;
  int BKV, n_kv, q_tiles;  // q_tiles: 128-query tiles
  int tiles_per_cta;       // consecutive query tiles of one (batch, head) per CTA (> 1 for short key sequences)
  float scale_log2;
  void* O;
  long long o_stride_b, o_stride_l;
  uint32_t idesc_qk, idesc_pv;
};

namespace {

__device__ __forceinline__ float ex2s(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void tmem_st_32x32s(uint32_t taddr, const uint32_t (&v)[32]) {
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
__device__ __forceinline__ void tmem_st_32x32_x16(uint32_t taddr, const uint32_t* v) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
        "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
// D[tmem] (+)= A[tmem] . B[smem]: A is M x K with one row per TMEM lane and two 16-bit K elements per 32-bit column
__device__ __forceinline__ void umma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_st_waits() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

constexpr int kQAtom = 128 * 128;  // bytes: 128 query rows x 64 halfs (one 128B-swizzle atom; a Dh = 128 tile is two of them)
constexpr int kKVAtom = 64 * 128;  // bytes: 64 key rows x 64 halfs
constexpr int kSlots = 4;
constexpr float kRescale = 8.0f;   // log2(256): O is rescaled only when a row's block maximum exceeds the reference by more
constexpr uint32_t kTmemCols = 128;
constexpr uint32_t kTmemColsP = 32;
#ifndef B200_ATTN64S_P_TMEM
#define B200_ATTN64S_P_TMEM 1
#endif
constexpr bool kPT = B200_ATTN64S_P_TMEM != 0;
#ifndef B200_ATTN64S_POLY_PAIRS
#define B200_ATTN64S_POLY_PAIRS 0x1
#endif
#ifndef B200_ATTN128S_P2
#define B200_ATTN128S_P2 0  // Dh = 128: 1 = P double-buffered in TMEM (measured: 1133 vs 1144-1151 TF/s single-buffered: no gain)
#endif
#ifndef B200_ATTN128S_POLY_PAIRS
#define B200_ATTN128S_POLY_PAIRS 0x0  // measured at B4 H24 L4352: none 1155, 1 of 4 1151, 2 of 4 1105 TF/s
#endif
constexpr unsigned kPolyPairs128S = B200_ATTN128S_POLY_PAIRS;  // the same choice for the Dh = 128 build
constexpr unsigned kPolyPairsS = B200_ATTN64S_POLY_PAIRS;  // element PAIRS (of the 4 per 8 elements) whose exp2 runs on the FMA pipe:
// measured (B16 H10 L4096 / B16 H20 L1024, TF/s): none 772 / 643, 1 of 4 790 / 652, 2 of 4 727 / 611, 3 of 4 653 / 560

}  // namespace

// __maxnreg__(80): 3 CTAs x 256 threads x 80 registers = 60 K of the SM's 64 K; with setmaxnreg in the kernel ptxas takes the
// cap as the launch-time count, which the dec / inc below redistribute (4 x (80 - 24) released = 4 x (136 - 80) taken).
//
// DH = 128 (Flux / SD3; round 2, late): the same pipeline with TWO CTAs per SM — per 64-key block the tensor core now has as
// much work as the MUFU (Q.K^T + P.V = 2 x 256 clk against 512 clk of exponentials), so the second CTA's MMAs run under the
// first one's softmax and vice versa.  Per CTA: Q 32 KB (one buffer: one query tile per CTA) + K/V ring 4 x 16 KB = 96 KB smem;
// TMEM 256 columns: S [0,64) | O [64,192) | P [192,224) [224,256) (double-buffered); Q / K / V tiles are two 64-wide swizzle atoms side by side; 128
// registers per thread at launch (48 control / 208 softmax after setmaxnreg).
template <bool BF16, int DH>
__global__ void __maxnreg__(DH == 64 ? 80 : 128)
attn64s_kernel(const __grid_constant__ CUtensorMap mapQ, const __grid_constant__ CUtensorMap mapK,
               const __grid_constant__ CUtensorMap mapV, const Attn64sParams p) {
  static_assert(DH == 64 || DH == 128, "head dim");
  static_assert(kPT || DH == 64, "P through shared memory is only kept for Dh = 64");
  constexpr int kAtoms = DH / 64;            // 64-wide swizzle atoms per row
  constexpr int kQTile = kAtoms * kQAtom;    // bytes of one Q tile
  constexpr int kKVTile = kAtoms * kKVAtom;  // bytes of one ring slot (64 keys)
  constexpr int kQBufs = DH == 64 ? 2 : 1;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t q_smem0 = base;                     // Q buffers: tile t lives in buffer t % kQBufs
  const uint32_t ring_smem = base + kQBufs * kQTile;
  const uint32_t p_smem = ring_smem + kSlots * kKVTile;        // P tile (only without kPT)
  const uint32_t bar_base = p_smem + (kPT ? 0u : (uint32_t)kQTile);
  auto q_full = [&](int i) { return bar_base + 8u * i; };
  auto q_empty = [&](int i) { return bar_base + 8u * (2 + i); };
  auto ring_full = [&](int i) { return bar_base + 8u * (4 + i); };
  auto ring_empty = [&](int i) { return bar_base + 8u * (4 + kSlots + i); };
  const uint32_t s_full = bar_base + 8u * (4 + 2 * kSlots);
  const uint32_t s_cons = s_full + 8u;
  // Dh = 128: P is double-buffered in TMEM (block g in buffer g & 1, 2 x 32 columns fit the 256-column allocation), so the
  // softmax of block g + 1 does not wait for P.V_g — it only needs P.V_{g-1} (its buffer's previous reader), and P.V_g itself
  // only on the rare lazy rescale of O.  Each buffer has its own pair of barriers (a single barrier could run two phases
  // ahead of a waiter, which a parity wait cannot tell from "not yet").
  constexpr bool kP2 = DH == 128 && B200_ATTN128S_P2 != 0;
  auto p_full = [&](int g) { return s_full + 16u + (kP2 ? 8u * (uint32_t)(g & 1) : 0u); };
  auto pv_done = [&](int g) { return s_full + 32u + (kP2 ? 8u * (uint32_t)(g & 1) : 0u); };
  auto pbuf_phase = [&](int g) { return (uint32_t)(kP2 ? (g >> 1) : g) & 1u; };
  const uint32_t tmem_slot = s_full + 48u;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int groups = (p.q_tiles + p.tiles_per_cta - 1) / p.tiles_per_cta;
  const int grp = blockIdx.x % groups;
  const int h = (blockIdx.x / groups) % p.H;
  const int b = blockIdx.x / (groups * p.H);
  const int qt0 = grp * p.tiles_per_cta;
  const int nt = min(p.tiles_per_cta, p.q_tiles - qt0);  // query tiles of this CTA
  const int BKV = p.BKV;
  const int n_kv = p.n_kv;
  const int n_blk = nt * n_kv;  // key blocks over all tiles: the pipeline below runs over this flat sequence

  pdl_launch_dependents();
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&mapQ);
    tma_prefetch_desc(&mapK);
    tma_prefetch_desc(&mapV);
    for (int i = 0; i < 2; ++i) {
      mbar_init(q_full(i), 1);
      mbar_init(q_empty(i), 1 + 128);  // the tile's last Q.K^T retired + the softmax threads are done staging O in the buffer
    }
    for (int i = 0; i < kSlots; ++i) {
      mbar_init(ring_full(i), 1);
      mbar_init(ring_empty(i), 1);
    }
    mbar_init(s_full, 1);
    mbar_init(s_cons, 128);
    for (int i = 0; i < 2; ++i) {
      mbar_init(p_full(i), 128);
      mbar_init(pv_done(i), 1);
    }
    mbar_fence_init();
  }
  if (warp == 1) {
    if constexpr (DH == 64) {
      tmem_alloc(tmem_slot, kTmemCols);
      if (kPT) tmem_alloc(tmem_slot + 4u, kTmemColsP);  // 3 CTAs x (128 + 32) columns fit the SM's 512
    } else {
      tmem_alloc(tmem_slot, 256);  // 2 CTAs x 256 columns
    }
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  uint32_t tmem_p = 0;
  if constexpr (DH == 64) {
    if (kPT) asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_p) : "r"(tmem_slot + 4u));
  } else {
    tmem_p = tmem_base + 192u;
  }
  pdl_wait();  // set-up done; q / k / v are the predecessor's output

  if (warp < 4) {
    if constexpr (DH == 64) asm volatile("setmaxnreg.dec.sync.aligned.u32 24;");
    else asm volatile("setmaxnreg.dec.sync.aligned.u32 48;");  // eight unrolled MMAs per Q.K^T: 24 registers spill in the issue loop
    if (warp >= 2) {
      // idle warps of the control warpgroup
    } else if (warp == 0) {
      if (lane == 0) {
        // ------------------------------------------------------------------ TMA producer
        const uint32_t kv_atom = (uint32_t)BKV * 128u;  // bytes of one 64-wide atom of a key block
        const uint32_t kv_bytes = kv_atom * kAtoms;
        for (int t = 0; t < nt; ++t) {
          const int buf = t % kQBufs;
          if (t >= kQBufs) mbar_wait_quiet(q_empty(buf), (uint32_t)((t / kQBufs) - 1) & 1u);
          mbar_expect_tx(q_full(buf), (uint32_t)kQTile);
#pragma unroll
          for (int a = 0; a < kAtoms; ++a)
            tma_load_3d(q_smem0 + buf * kQTile + a * kQAtom, &mapQ, q_full(buf), h * DH + a * 64, (qt0 + t) * 128, b);
          for (int i = 0; i < 2 * n_kv; ++i) {  // even: K_{i/2}, odd: V_{i/2}; the ring index runs on across tiles
            const int idx = t * 2 * n_kv + i;
            const int slot = idx % kSlots;
            const uint32_t phase = (uint32_t)(idx / kSlots) & 1u;
            mbar_wait_quiet(ring_empty(slot), phase ^ 1u);
            mbar_expect_tx(ring_full(slot), kv_bytes);
#pragma unroll
            for (int a = 0; a < kAtoms; ++a)
              tma_load_3d(ring_smem + slot * kKVTile + a * kv_atom, (i & 1) ? &mapV : &mapK, ring_full(slot), h * DH + a * 64,
                          (i >> 1) * BKV, b);
          }
        }
      }
    } else {
      // ------------------------------------------------------------------ MMA issuer (converged warp, elected lane)
      const uint64_t qdesc0 = make_smem_desc_sw128(q_smem0, 0, 1024);
      const uint64_t kdesc0 = make_smem_desc_sw128(ring_smem, 0, 1024);
      const uint32_t kv_atom = (uint32_t)BKV * 128u;
      const uint64_t vdesc0 = make_smem_desc_sw128(ring_smem, kv_atom, 1024);  // MN-major V: 64-dim atoms kv_atom apart (LBO)
      const uint64_t pdesc = make_smem_desc_sw128(p_smem, 0, 1024);
      const uint32_t idesc_qk = p.idesc_qk, idesc_pv = p.idesc_pv;
      const uint32_t s_tmem = tmem_base, o_tmem = tmem_base + 64u;
      const int ksteps = BKV >> 4;
      const uint64_t k_atom_enc = (uint64_t)(kv_atom >> 4);
      auto wait_full = [&](int idx) {
        mbar_wait_quiet(ring_full(idx % kSlots), (uint32_t)(idx / kSlots) & 1u);
        tc_fence_after();
      };
      // S = Q_t K_j^T for flat block g = t * n_kv + j (M 128, N BKV, K 64); releases the K slot, and the Q buffer after a
      // tile's last block
      auto issue_qk = [&](int g, int t, int j) {
        if (j == 0) {
          mbar_wait_quiet(q_full(t % kQBufs), (uint32_t)(t / kQBufs) & 1u);
          tc_fence_after();
        }
        const int idx = 2 * g;
        wait_full(idx);
        const uint64_t qd = qdesc0 + (uint64_t)((t % kQBufs) * (kQTile >> 4));
        const uint64_t kd = kdesc0 + (uint64_t)((idx % kSlots) * (kKVTile >> 4));
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < DH / 16; ++k)  // 16 dims per step: +32 B inside a swizzle atom, the next atom after four steps
            umma_f16(s_tmem, qd + (uint64_t)((k >> 2) * (kQAtom >> 4) + (k & 3) * 2), kd + (uint64_t)(k >> 2) * k_atom_enc + (uint64_t)((k & 3) * 2),
                     idesc_qk, k != 0 ? 1u : 0u);
          umma_commit(s_full);
          umma_commit(ring_empty(idx % kSlots));
          if (j == n_kv - 1) umma_commit(q_empty(t % kQBufs));
        }
        __syncwarp();
      };
      issue_qk(0, 0, 0);
      int t = 0, j = 0;  // tile and block of g
#pragma unroll 1
      for (int g = 0; g < n_blk; ++g) {
        const int vidx = 2 * g + 1;
        const int jn = (j + 1 == n_kv) ? 0 : j + 1, tn = (j + 1 == n_kv) ? t + 1 : t;
        if (g + 1 < n_blk) {  // Q.K^T of the next block (or of the next tile's first) as soon as S_g sits in registers
          mbar_wait_quiet(s_cons, (uint32_t)g & 1u);
          tc_fence_after();
          issue_qk(g + 1, tn, jn);
        }
        wait_full(vidx);
        const uint64_t vd = vdesc0 + (uint64_t)((vidx % kSlots) * (kKVTile >> 4));
        mbar_wait_quiet(p_full(g), pbuf_phase(g));
        tc_fence_after();
        const uint32_t pa = tmem_p + (kP2 ? (uint32_t)(g & 1) * 32u : 0u);
        const uint32_t acc0 = j != 0 ? 1u : 0u;  // a tile's first P.V overwrites O (its predecessor's O was read before P_g was published)
        if (elect_one()) {
          if (kPT) {  // A = P from TMEM: 16 keys = 8 columns per k step
            if (ksteps == 4) {
#pragma unroll
              for (int kk = 0; kk < 4; ++kk)
                umma_f16_ts(o_tmem, pa + (uint32_t)(kk * 8), vd + (uint64_t)(kk * (2048 >> 4)), idesc_pv, kk ? 1u : acc0);
            } else {
              for (int kk = 0; kk < ksteps; ++kk)
                umma_f16_ts(o_tmem, pa + (uint32_t)(kk * 8), vd + (uint64_t)(kk * (2048 >> 4)), idesc_pv, kk ? 1u : acc0);
            }
          } else if (ksteps == 4) {
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
              umma_f16(o_tmem, pdesc + (uint64_t)(kk * 2), vd + (uint64_t)(kk * (2048 >> 4)), idesc_pv, kk ? 1u : acc0);
          } else {
            for (int kk = 0; kk < ksteps; ++kk)
              umma_f16(o_tmem, pdesc + (uint64_t)(kk * 2), vd + (uint64_t)(kk * (2048 >> 4)), idesc_pv, kk ? 1u : acc0);
          }
          umma_commit(pv_done(g));
          umma_commit(ring_empty(vidx % kSlots));
        }
        __syncwarp();
        t = tn;
        j = jn;
      }
    }
  } else {
    if constexpr (DH == 64) asm volatile("setmaxnreg.inc.sync.aligned.u32 136;");  // 4 x 136 + 4 x 24 = 8 x 80 (the launch-time allocation)
    else asm volatile("setmaxnreg.inc.sync.aligned.u32 208;");                      // 4 x 208 + 4 x 48 = 8 x 128
    // ------------------------------------------------------------------ softmax warpgroup (warps 4-7)
    const int quad = warp & 3;       // TMEM lane quadrant of this warp
    const int r = quad * 32 + lane;  // row inside the tile
    const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
    const uint32_t s_addr = tmem_base + lane_addr;
    const uint32_t o_addr = tmem_base + 64u + lane_addr;
    const uint32_t p_addr0 = tmem_p + lane_addr;
    const uint32_t p_row = p_smem + (uint32_t)r * 128u;
    const uint32_t sw = (uint32_t)(r & 7);
    const float sl2 = p.scale_log2;

    for (int t = 0; t < nt; ++t) {
    float m_ref = -INFINITY, l_run = 0.f;
    for (int j = 0; j < n_kv; ++j) {
      const int g = t * n_kv + j;  // flat block index: every barrier's phase runs on across the CTA's tiles
      int nvalid = p.Lk - j * BKV;
      if (nvalid > BKV) nvalid = BKV;
      mbar_wait(s_full, (uint32_t)g & 1u);
      tc_fence_after();
      uint32_t v[64];
      tmem_ld_32x32(s_addr + 0u, *reinterpret_cast<uint32_t(*)[32]>(&v[0]));
      tmem_ld_32x32(s_addr + 32u, *reinterpret_cast<uint32_t(*)[32]>(&v[32]));
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(s_cons);  // S is in registers: the tensor core may overwrite it with QK_{j+1}
      const bool full_blk = nvalid == 64;
      float mx;
      if (full_blk) {
        float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
        for (int i = 0; i < 64; i += 4) {
          m0 = fmaxf(m0, fmaxf(__uint_as_float(v[i + 0]), __uint_as_float(v[i + 1])));
          m1 = fmaxf(m1, fmaxf(__uint_as_float(v[i + 2]), __uint_as_float(v[i + 3])));
        }
        mx = fmaxf(m0, m1);
      } else {
        mx = -INFINITY;
#pragma unroll
        for (int i = 0; i < 64; ++i)
          if (i < nvalid) mx = fmaxf(mx, __uint_as_float(v[i]));
      }
      const float m_blk = mx * sl2;
      // PV_{g-1} must have retired before O is rescaled and before P is overwritten (a tile's first block: the previous
      // tile's output phase has waited for it).  Moving this wait behind the first 32 exponentials was measured: slower
      // (825 -> 785 TF/s at L4096) — the warps of a CTA then reach their MUFU phases together instead of staggered.
      if (kP2) {
        if (j > 1) {  // this P buffer's previous reader
          mbar_wait(pv_done(g - 2), pbuf_phase(g - 2));
          tc_fence_after();
        }
      } else if (j > 0) {
        mbar_wait(pv_done(g - 1), pbuf_phase(g - 1));
        tc_fence_after();
      }
      if (j == 0) {
        m_ref = m_blk;
      } else {
        const bool need = m_blk > m_ref + kRescale;
        if (__any_sync(0xffffffffu, need)) {  // lazy rescale: rare after the first few key blocks
          if (kP2) {  // O must hold every earlier block
            mbar_wait(pv_done(g - 1), pbuf_phase(g - 1));
            tc_fence_after();
          }
          const float alpha = need ? ex2s(m_ref - m_blk) : 1.0f;
#pragma unroll
          for (int c = 0; c < DH; c += 32) {
            uint32_t w[32];
            tmem_ld_32x32(o_addr + (uint32_t)c, w);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) w[i] = __float_as_uint(__uint_as_float(w[i]) * alpha);
            tmem_st_32x32s(o_addr + (uint32_t)c, w);
          }
          tmem_st_waits();
          l_run *= alpha;
          if (need) m_ref = m_blk;
        }
      }
      // p = exp2(s * scale - m_ref), row sum, P -> TMEM (two 16-column stores, the first under the second half's
      // exponentials) or -> smem (128B-swizzled K-major A operand: one 64-key atom per row)
      const float nm = -m_ref;
      const uint32_t p_addr = p_addr0 + (kP2 ? (uint32_t)(g & 1) * 32u : 0u);
      float rs;
      uint32_t pq[16];  // kPT: 32 keys of this row as fp16 / bf16 pairs
      auto put8 = [&](int c, const float (&pe)[8]) {
        if (kPT) {
          const int o = (c >> 1) & 15;
          pq[o + 0] = pack2<BF16>(pe[0], pe[1]);
          pq[o + 1] = pack2<BF16>(pe[2], pe[3]);
          pq[o + 2] = pack2<BF16>(pe[4], pe[5]);
          pq[o + 3] = pack2<BF16>(pe[6], pe[7]);
          if ((c & 31) == 24) tmem_st_32x32_x16(p_addr + (uint32_t)(c >> 5) * 16u, pq);
        } else {
          const uint32_t addr = p_row + ((((uint32_t)c >> 3) ^ sw) << 4);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(pack2<BF16>(pe[0], pe[1])),
                       "r"(pack2<BF16>(pe[2], pe[3])), "r"(pack2<BF16>(pe[4], pe[5])), "r"(pack2<BF16>(pe[6], pe[7])));
        }
      };
      if (full_blk) {
        const f32x2_t sl2p = pk2(sl2, sl2), nmp = pk2(nm, nm);
        f32x2_t acc0 = pk2(0.f, 0.f), acc1 = pk2(0.f, 0.f);
#pragma unroll
        for (int c = 0; c < 64; c += 8) {
          float pe[8];
#pragma unroll
          for (int q2 = 0; q2 < 4; ++q2) {
            const f32x2_t x = fma2(pk2(__uint_as_float(v[c + 2 * q2]), __uint_as_float(v[c + 2 * q2 + 1])), sl2p, nmp);
            if (((DH == 64 ? kPolyPairsS : kPolyPairs128S) >> q2) & 1) {
              exp2_poly3_x2(x, pe[2 * q2], pe[2 * q2 + 1]);
            } else {
              float xa, xb;
              upk2(x, xa, xb);
              pe[2 * q2] = ex2s(xa);
              pe[2 * q2 + 1] = ex2s(xb);
            }
          }
          acc0 = add2(acc0, pk2(pe[0], pe[1]));
          acc1 = add2(acc1, pk2(pe[2], pe[3]));
          acc0 = add2(acc0, pk2(pe[4], pe[5]));
          acc1 = add2(acc1, pk2(pe[6], pe[7]));
          put8(c, pe);
        }
        float a0, a1, b0, b1;
        upk2(acc0, a0, a1);
        upk2(acc1, b0, b1);
        rs = (a0 + b0) + (a1 + b1);
      } else {
        float rs0 = 0.f, rs1 = 0.f;
#pragma unroll
        for (int c = 0; c < 64; c += 8) {
          if (kPT || c < BKV) {  // kPT: the whole 32-column P region is written (zeros past the block), the MMA reads BKV / 2
            float pe[8];
            if (c < nvalid) {  // warp-uniform: groups past the block's last key cost no MUFU slots (cross-attention: 13 of 64)
#pragma unroll
              for (int i = 0; i < 8; ++i) pe[i] = (c + i < nvalid) ? ex2s(fmaf(__uint_as_float(v[c + i]), sl2, nm)) : 0.f;
            } else {
#pragma unroll
              for (int i = 0; i < 8; ++i) pe[i] = 0.f;
            }
            rs0 += (pe[0] + pe[2]) + (pe[4] + pe[6]);
            rs1 += (pe[1] + pe[3]) + (pe[5] + pe[7]);
            put8(c, pe);
          }
        }
        rs = rs0 + rs1;
      }
      l_run += rs;
      if (kPT) tmem_st_waits();
      else fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(p_full(g));
    }

    // ---- output: O / l -> fp16 / bf16 -> warp-private staging (the Q tile, or the smem P tile, is free now) -> coalesced stores
    mbar_wait(pv_done(t * n_kv + n_kv - 1), pbuf_phase(t * n_kv + n_kv - 1));
    tc_fence_after();
    const float inv = 1.0f / l_run;
    const int q0 = (qt0 + t) * 128;
    constexpr uint32_t kRowB = DH * 2;  // bytes of one output row
    const uint32_t stg = (kPT ? q_smem0 + (uint32_t)(t % kQBufs) * kQTile : p_smem) + (uint32_t)quad * (32u * kRowB);  // 32 rows per warp
    __syncwarp();
#pragma unroll
    for (int c = 0; c < DH; c += 32) {
      uint32_t v[32];
      tmem_ld_32x32(o_addr + (uint32_t)c, v);
      tmem_ld_wait();
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const uint32_t o0 = pack2<BF16>(__uint_as_float(v[g * 8 + 0]) * inv, __uint_as_float(v[g * 8 + 1]) * inv);
        const uint32_t o1 = pack2<BF16>(__uint_as_float(v[g * 8 + 2]) * inv, __uint_as_float(v[g * 8 + 3]) * inv);
        const uint32_t o2 = pack2<BF16>(__uint_as_float(v[g * 8 + 4]) * inv, __uint_as_float(v[g * 8 + 5]) * inv);
        const uint32_t o3 = pack2<BF16>(__uint_as_float(v[g * 8 + 6]) * inv, __uint_as_float(v[g * 8 + 7]) * inv);
        const uint32_t chunk = (uint32_t)(c >> 3) + (uint32_t)g;  // 16B chunk index inside the row
        const uint32_t addr = stg + (uint32_t)lane * kRowB + ((chunk ^ (uint32_t)(lane & 7)) << 4);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(o0), "r"(o1), "r"(o2), "r"(o3) : "memory");
      }
    }
    __syncwarp();
    {
      constexpr int kPieces = DH / 8;       // 16-byte pieces per row: a warp's store covers 32 / kPieces whole rows
      constexpr int kRowsPer = 32 / kPieces;
      const int piece = lane & (kPieces - 1);
      char* obase = reinterpret_cast<char*>(p.O) + ((size_t)b * p.o_stride_b + (size_t)h * DH) * 2;
#pragma unroll
      for (int jj = 0; jj < 32 / kRowsPer; ++jj) {
        const int rl = jj * kRowsPer + lane / kPieces;
        const uint32_t addr = stg + (uint32_t)rl * kRowB + ((((uint32_t)piece) ^ (uint32_t)(rl & 7)) << 4);
        uint32_t o0, o1, o2, o3;
        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(o0), "=r"(o1), "=r"(o2), "=r"(o3) : "r"(addr));
        const int q = q0 + quad * 32 + rl;
        if (q < p.Lq)
          *reinterpret_cast<uint4*>(obase + ((size_t)q * p.o_stride_l + piece * 8) * 2) = make_uint4(o0, o1, o2, o3);
      }
    }
    tc_fence_before();          // O has been read out of TMEM: the next tile's first P.V may overwrite it (ordered by p_full)
    mbar_arrive(q_empty(t % kQBufs));  // ... and this tile's Q buffer (the staging tile) may take the tile after next
    }  // tiles
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    if constexpr (DH == 64) {
      tmem_dealloc(tmem_base, kTmemCols);
      if (kPT) tmem_dealloc(tmem_p, kTmemColsP);
    } else {
      tmem_dealloc(tmem_base, 256);
    }
  }
}

template <bool BF16, int DH>
static int launch_attn64s(const CUtensorMap& mQ, const CUtensorMap& mK, const CUtensorMap& mV, const Attn64sParams& p,
                          cudaStream_t stream) {
  constexpr int kQTile = (DH / 64) * kQAtom, kKVTile = (DH / 64) * kKVAtom;
  const size_t smem = DH == 64 ? (size_t)kQTile * (kPT ? 2 : 3) + (size_t)kSlots * kKVTile + 1024 + 256
                               : (size_t)kQTile + (size_t)kSlots * kKVTile + 1024 + 256;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(attn64s_kernel<BF16, DH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      set_error("attention%ds: smem attr: %s", DH, cudaGetErrorString(e));
      return B200_ECUDA;
    }
    attr_done = true;
  }
  const int grid = ((p.q_tiles + p.tiles_per_cta - 1) / p.tiles_per_cta) * p.H * p.B;
  cudaError_t e = launch_pdl(attn64s_kernel<BF16, DH>, dim3(grid), dim3(256), smem, stream, 1, mQ, mK, mV, p);
  if (e != cudaSuccess) {
    set_error("attention%ds: launch failed: %s", DH, cudaGetErrorString(e));
    return B200_ECUDA;
  }
  B200_CHECK_LAUNCH("attention64s");
  return B200_OK;
}

template <int DH>
static int attention_s_dispatch(const void* q, const void* k, const void* v, void* o, const b200_attn_desc* d, cudaStream_t st) {
  Attn64sParams p;
  memset(&p, 0, sizeof(p));
  p.B = d->B;
  p.H = d->H;
  p.Lq = d->Lq;
  p.Lk = d->Lk;
  p.BKV = d->Lk >= 64 ? 64 : ((d->Lk + 15) / 16) * 16;
  p.n_kv = (d->Lk + p.BKV - 1) / p.BKV;
  p.q_tiles = (d->Lq + 127) / 128;
  // Short key sequences (cross-attention: 77 keys = 2 blocks): a CTA's life is a latency chain (set-up, Q from DRAM, two
  // MMA -> softmax round trips, output) with ~1 us of work in it; several consecutive query tiles per CTA put the next
  // tile's Q load and first Q.K^T under the current tile's softmax and output phases.  B200_ATTN64S_TP forces a count.
  // (Dh = 128 has one Q buffer: always one tile per CTA.)
  if (DH == 64) {
    static int forced = -1;
    if (forced < 0) {
      const char* e = getenv("B200_ATTN64S_TP");
      forced = e ? atoi(e) : 0;
    }
    int tp = 1;
    if (forced > 0) {
      tp = forced;
    } else if (p.n_kv <= 2) {
      const long long tiles = (long long)p.q_tiles * d->H * d->B;
      const long long resident = 3LL * num_sms();
      tp = (int)(tiles / (2 * resident));  // keep at least ~2 waves of CTAs
      if (tp > 4) tp = 4;
    }
    if (tp < 1) tp = 1;
    if (tp > p.q_tiles) tp = p.q_tiles;
    p.tiles_per_cta = tp;
  } else {
    p.tiles_per_cta = 1;
  }
  p.scale_log2 = d->scale * 1.4426950408889634f;
  p.O = o;
  p.o_stride_b = d->o_stride_b;
  p.o_stride_l = d->o_stride_l;
  const bool bf = d->dtype == B200_BF16;
  p.idesc_qk = make_idesc_f16(128, p.BKV, bf, false, false);
  p.idesc_pv = make_idesc_f16(128, DH, bf, false, true);
  const uint64_t cols = (uint64_t)d->H * DH;
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
  return bf ? launch_attn64s<true, DH>(mQ, mK, mV, p, st) : launch_attn64s<false, DH>(mQ, mK, mV, p, st);
}

// called from attention64_dispatch (attention64.cu) when B200_ATTN64_VER = 2
int attention64s_dispatch(const void* q, const void* k, const void* v, void* o, const b200_attn_desc* d, cudaStream_t st) {
  return attention_s_dispatch<64>(q, k, v, o, d, st);
}
// called from attention128_dispatch (attention128.cu) unless B200_ATTN128_VER = 0
int attention128s_dispatch(const void* q, const void* k, const void* v, void* o, const b200_attn_desc* d, cudaStream_t st) {
  return attention_s_dispatch<128>(q, k, v, o, d, st);
}

}  // namespace b200
