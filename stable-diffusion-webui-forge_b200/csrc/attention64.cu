// b200forge — head-dim-64 attention forward, the SDXL / SD2.x shape (every SDXL attention has Dh = 64).
//
// One CTA per SM works on TWO 128-query tiles; per tile one MMA-issuing thread and one softmax warpgroup
// (4 warps, thread = query row):
//   QK_j (S_t in TMEM) -> softmax_j: whole S row to registers in one TMEM round trip, lazy running max,
//   P = exp2(.) computed two-per-MUFU-op in the operand precision and written back OVER S in TMEM
//   -> PV_j as a TS-mode MMA (A = P from tensor memory, B = [V_j | ones] straight from the [key, Dh] tile as an
//   MN-major operand) accumulating O_t and the row sums in TMEM across key blocks -> QK_{j+1} ...
// The other tile's MMAs fill the tensor pipe while this tile is in its softmax.  The tensor pipe executes in issue
// order, so "S_{j+1} full" implies "PV_j retired": O is rescaled in place (tcgen05.ld/st) only when a row's block
// max exceeds the reference max by 2^8.  No P round trip through shared memory, no row-sum FADDs, no O read-back.
// TMEM columns: S0/P0 [0,128)  S1/P1 [128,256)  O0|l0 [256,336)  O1|l1 [336,416).
// smem: Q 2x16 KB | K/V ring 4x16 KB | ones 16 KB.
#include "common.cuh"
#include "host_util.h"

namespace b200 {

struct Attn64Params {
  int B, H, Lq, Lk;
  int BKV, n_kv, q_tiles;  // q_tiles: 256-query CTA tiles
  float scale_log2;
  void* O;
  long long o_stride_b, o_stride_l;
  uint32_t idesc_qk, idesc_pv, idesc_l;
};

__device__ __forceinline__ float ex2a(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// two exponentials per MUFU op on packed 16-bit lanes: P is consumed as fp16/bf16 by the PV MMA anyway, and
// the row sum is taken by the tensor core from the same rounded values (ones-column MMA), so the softmax stays
// self-consistent.  Halves the MUFU work that bounds this kernel.
template <bool BF16>
__device__ __forceinline__ uint32_t ex2_pack(float a, float b) {
  uint32_t r;
  if constexpr (BF16) {
    asm("{\n\t.reg .b32 t;\n\tcvt.rn.bf16x2.f32 t, %2, %1;\n\tex2.approx.ftz.bf16x2 %0, t;\n\t}" : "=r"(r) : "f"(a), "f"(b));
  } else {
    asm("{\n\t.reg .b32 t;\n\tcvt.rn.f16x2.f32 t, %2, %1;\n\tex2.approx.f16x2 %0, t;\n\t}" : "=r"(r) : "f"(a), "f"(b));
  }
  return r;
}
__device__ __forceinline__ void tmem_ld_32x32_x1(uint32_t taddr, uint32_t& v) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_st_32x32_x1(uint32_t taddr, uint32_t v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(taddr), "r"(v) : "memory");
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

static constexpr int kTile = 128 * 128;  // bytes of one 128-row x 64-half tile
static constexpr int kRingSlots = 4;
static constexpr float kRescaleThreshold = 8.0f;  // log2(256)

// D[tmem] (+)= A[tmem] * B[smem desc]: the P operand of P.V is read straight from tensor memory
__device__ __forceinline__ void umma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

template <bool BF16>
__global__ void __launch_bounds__(384, 1)
attn64_kernel(const __grid_constant__ CUtensorMap mapQ, const __grid_constant__ CUtensorMap mapK,
              const __grid_constant__ CUtensorMap mapV, const Attn64Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t q_smem = base;                              // 2 tiles (reused as output staging at the end)
  const uint32_t ring_smem = base + 2 * kTile;               // 4 K/V tiles
  const uint32_t ones_smem = ring_smem + kRingSlots * kTile; // [128 keys][64] of 1.0: second MN atom of the PV B operand
  const uint32_t bar_base = ones_smem + kTile;
  const uint32_t q_full = bar_base;
  auto ring_full = [&](int i) { return bar_base + 8u * (1 + i); };
  auto ring_empty = [&](int i) { return bar_base + 8u * (1 + kRingSlots + i); };
  auto s_full = [&](int t) { return bar_base + 8u * (1 + 2 * kRingSlots + t); };
  auto p_full = [&](int t) { return bar_base + 8u * (3 + 2 * kRingSlots + t); };
  auto o_final = [&](int t) { return bar_base + 8u * (5 + 2 * kRingSlots + t); };
  const uint32_t tmem_slot = bar_base + 8u * (7 + 2 * kRingSlots);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int qt = blockIdx.x % p.q_tiles;
  const int h = (blockIdx.x / p.q_tiles) % p.H;
  const int b = blockIdx.x / (p.q_tiles * p.H);
  const int q0 = qt * 256;
  const int BKV = p.BKV;
  const int n_kv = p.n_kv;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&mapQ);
    tma_prefetch_desc(&mapK);
    tma_prefetch_desc(&mapV);
    mbar_init(q_full, 1);
    for (int i = 0; i < kRingSlots; ++i) {
      mbar_init(ring_full(i), 1);
      mbar_init(ring_empty(i), 2);  // both tiles' issuers release a K/V slot
    }
    for (int t = 0; t < 2; ++t) {
      mbar_init(s_full(t), 1);
      mbar_init(p_full(t), 128);
      mbar_init(o_final(t), 1);
    }
    mbar_fence_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  if (warp >= 4) {  // fill the ones tile (every element 1.0, so the swizzle is irrelevant)
    const uint32_t one2 = BF16 ? 0x3F803F80u : 0x3C003C00u;
    for (int i = (int)threadIdx.x - 128; i < kTile / 16; i += 256)
      asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(ones_smem + 16u * i), "r"(one2) : "memory");
    fence_proxy_async_smem();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp < 4) {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 72;");
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
  } else if ((warp == 1 || warp == 3) && lane == 0) {
    // ------------------------------------------------------------------ MMA issuers: warp 1 -> tile 0, warp 3 -> tile 1
    // Per tile the chain is  QK_j -> softmax_j (P written over S in TMEM) -> PV_j -> QK_{j+1} ...; the tensor pipe
    // runs in issue order, so QK_{j+1} cannot overwrite P_j before PV_j has consumed it, and "S_{j+1} full"
    // implies "O holds blocks <= j".  The other tile's MMAs fill the pipe while this tile is in its softmax.
    const int t = warp >> 1;
    const uint32_t s_tmem = tmem_base + (uint32_t)t * 128u;   // S_t; P_t aliases its first 64 columns
    const uint32_t o_tmem = tmem_base + 256u + (uint32_t)t * 80u;  // O_t [64] | row sums [16]
    const uint64_t qdesc = make_smem_desc_sw128(q_smem + t * kTile, 0, 1024);
    const uint64_t kdesc0 = make_smem_desc_sw128(ring_smem, 0, 1024);
    const uint32_t idesc_qk = p.idesc_qk, idesc_pv = p.idesc_pv;
    auto wait_full = [&](int idx) {
      mbar_wait(ring_full(idx % kRingSlots), (uint32_t)(idx / kRingSlots) & 1u);
      tc_fence_after();
    };
    auto issue_qk = [&](int idx) {
      const uint64_t kd = kdesc0 + (uint64_t)((idx % kRingSlots) * (kTile >> 4));
      umma_f16(s_tmem, qdesc, kd, idesc_qk, 0u);
      umma_f16(s_tmem, qdesc + 2, kd + 2, idesc_qk, 1u);
      umma_f16(s_tmem, qdesc + 4, kd + 4, idesc_qk, 1u);
      umma_f16(s_tmem, qdesc + 6, kd + 6, idesc_qk, 1u);
      umma_commit(s_full(t));
      umma_commit(ring_empty(idx % kRingSlots));
    };
    mbar_wait(q_full, 0);
    wait_full(0);
    issue_qk(0);
    const int ksteps = BKV >> 4;
    for (int j = 0; j < n_kv; ++j) {
      const int vidx = 2 * j + 1, kidx = 2 * j + 2;
      wait_full(vidx);
      mbar_wait(p_full(t), (uint32_t)j & 1u);
      tc_fence_after();
      // B = [V_j | ones]: 64 value columns from the ring slot (MN-major), 16 more from the ones tile via LBO
      const uint32_t v_addr = ring_smem + (uint32_t)(vidx % kRingSlots) * kTile;
      const uint64_t vd = make_smem_desc_sw128(v_addr, ones_smem - v_addr, 1024);
      const uint32_t acc0 = j != 0 ? 1u : 0u;
      if (ksteps == 8) {
#pragma unroll
        for (int kk = 0; kk < 8; ++kk)
          umma_f16_ts(o_tmem, s_tmem + (uint32_t)kk * 8u, vd + (uint64_t)(kk * (2048 >> 4)), idesc_pv, kk ? 1u : acc0);
      } else {
        for (int kk = 0; kk < ksteps; ++kk)
          umma_f16_ts(o_tmem, s_tmem + (uint32_t)kk * 8u, vd + (uint64_t)(kk * (2048 >> 4)), idesc_pv, kk ? 1u : acc0);
      }
      umma_commit(ring_empty(vidx % kRingSlots));
      if (j + 1 < n_kv) {
        wait_full(kidx);
        issue_qk(kidx);
      } else {
        umma_commit(o_final(t));
      }
    }
  }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 216;");
    // ------------------------------------------------------------------ softmax warpgroups (t = tile, thread = row)
    const int t = (warp - 4) >> 2;
    const int quad = warp & 3;
    const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
    const uint32_t s_addr = tmem_base + (uint32_t)t * 128u + lane_addr;
    const uint32_t o_addr = tmem_base + 256u + (uint32_t)t * 80u + lane_addr;
    const float sl2 = p.scale_log2;
    float m_ref = -INFINITY;

    for (int j = 0; j < n_kv; ++j) {
      int nvalid = p.Lk - j * BKV;
      if (nvalid > BKV) nvalid = BKV;
      mbar_wait(s_full(t), (uint32_t)j & 1u);
      tc_fence_after();
      uint32_t v[128];
      tmem_ld_32x32(s_addr + 0u, *reinterpret_cast<uint32_t(*)[32]>(&v[0]));
      tmem_ld_32x32(s_addr + 32u, *reinterpret_cast<uint32_t(*)[32]>(&v[32]));
      tmem_ld_32x32(s_addr + 64u, *reinterpret_cast<uint32_t(*)[32]>(&v[64]));
      tmem_ld_32x32(s_addr + 96u, *reinterpret_cast<uint32_t(*)[32]>(&v[96]));
      tmem_ld_wait();
      const bool full_blk = nvalid == 128;
      float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
      if (full_blk) {
#pragma unroll
        for (int i = 0; i < 128; i += 8) {
          mx0 = fmaxf(mx0, fmaxf(__uint_as_float(v[i]), __uint_as_float(v[i + 1])));
          mx1 = fmaxf(mx1, fmaxf(__uint_as_float(v[i + 2]), __uint_as_float(v[i + 3])));
          mx2 = fmaxf(mx2, fmaxf(__uint_as_float(v[i + 4]), __uint_as_float(v[i + 5])));
          mx3 = fmaxf(mx3, fmaxf(__uint_as_float(v[i + 6]), __uint_as_float(v[i + 7])));
        }
      } else {
#pragma unroll
        for (int i = 0; i < 128; ++i)
          if (i < nvalid) mx0 = fmaxf(mx0, __uint_as_float(v[i]));
      }
      const float m_blk = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)) * sl2;
      if (j == 0) {
        m_ref = m_blk;
      } else {
        const bool need = m_blk > m_ref + kRescaleThreshold;
        if (__any_sync(0xffffffffu, need)) {
          // S_j full => PV_{j-1} retired (in-order tensor pipe): rescale O and the row sums in place
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
          {
            uint32_t lv;
            tmem_ld_32x32_x1(o_addr + 64u, lv);
            tmem_ld_wait();
            tmem_st_32x32_x1(o_addr + 64u, __float_as_uint(__uint_as_float(lv) * alpha));
          }
          if (need) m_ref = m_blk;
        }
      }
      // P = exp2(s*scale*log2e - m_ref) in the operand precision (2 per MUFU op), written over S in TMEM:
      // column c of P_t holds keys (2c, 2c+1) -> the K-major A operand of the TS-mode P.V MMA
      const float nm = -m_ref;
      uint32_t pk[64];
#pragma unroll
      for (int i = 0; i < 64; ++i) {
        float t0 = fmaf(__uint_as_float(v[2 * i]), sl2, nm);
        float t1 = fmaf(__uint_as_float(v[2 * i + 1]), sl2, nm);
        if (!full_blk) {
          if (2 * i >= nvalid) t0 = -INFINITY;
          if (2 * i + 1 >= nvalid) t1 = -INFINITY;
        }
        pk[i] = ex2_pack<BF16>(t0, t1);
      }
      tmem_st_32x32(s_addr + 0u, *reinterpret_cast<uint32_t(*)[32]>(&pk[0]));
      tmem_st_32x32(s_addr + 32u, *reinterpret_cast<uint32_t(*)[32]>(&pk[32]));
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(p_full(t));
    }

    // ---- output: O_t / l -> fp16 -> warp-private staging (the Q tile is free now) -> coalesced stores
    mbar_wait(o_final(t), 0);
    tc_fence_after();
    uint32_t lsum;
    tmem_ld_32x32_x1(o_addr + 64u, lsum);
    tmem_ld_wait();
    const float inv = 1.0f / __uint_as_float(lsum);
    const uint32_t stg = q_smem + (uint32_t)t * kTile + (uint32_t)quad * 4096u;  // 32 rows x 128 B per warp
    __syncwarp();
#pragma unroll
    for (int c = 0; c < 64; c += 32) {
      uint32_t w[32];
      tmem_ld_32x32(o_addr + (uint32_t)c, w);
      tmem_ld_wait();
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const uint32_t o0 = pack2<BF16>(__uint_as_float(w[g * 8 + 0]) * inv, __uint_as_float(w[g * 8 + 1]) * inv);
        const uint32_t o1 = pack2<BF16>(__uint_as_float(w[g * 8 + 2]) * inv, __uint_as_float(w[g * 8 + 3]) * inv);
        const uint32_t o2 = pack2<BF16>(__uint_as_float(w[g * 8 + 4]) * inv, __uint_as_float(w[g * 8 + 5]) * inv);
        const uint32_t o3 = pack2<BF16>(__uint_as_float(w[g * 8 + 6]) * inv, __uint_as_float(w[g * 8 + 7]) * inv);
        const uint32_t chunk = (uint32_t)(c >> 3) + (uint32_t)g;
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

template <bool BF16>
static int launch_attn64(const CUtensorMap& mQ, const CUtensorMap& mK, const CUtensorMap& mV, const Attn64Params& p,
                         cudaStream_t stream) {
  const size_t smem = (size_t)kTile * (2 + kRingSlots + 1) + 1024 + 256;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(attn64_kernel<BF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      set_error("attention64: smem attr: %s", cudaGetErrorString(e));
      return B200_ECUDA;
    }
    attr_done = true;
  }
  const int grid = p.q_tiles * p.H * p.B;
  attn64_kernel<BF16><<<grid, 384, smem, stream>>>(mQ, mK, mV, p);
  B200_CHECK_LAUNCH("attention64");
  return B200_OK;
}

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
  p.idesc_pv = make_idesc_f16(128, 80, bf, false, true);  // N = 64 value columns + 16 row-sum columns
  p.idesc_l = 0;
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
  return bf ? launch_attn64<true>(mQ, mK, mV, p, st) : launch_attn64<false>(mQ, mK, mV, p, st);
}

}  // namespace b200
