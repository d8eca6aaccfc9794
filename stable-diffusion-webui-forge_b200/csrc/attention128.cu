// b200forge — head-dim-128 attention forward: the Flux / SD3 shape (24 heads x 128, 4096 + 256 tokens).
//
// Same organisation as attention64.cu (two 128-query tiles per CTA, one CTA per SM, O accumulated in TMEM with lazy
// rescaling, the whole S row pulled into registers, QK_{j+1} issued underneath the exponentials of block j, the two
// softmax warpgroups taking turns on the MUFU-bound phase), re-sized for Dh = 128:
//   * per tile-block the tensor core now has as much work as the MUFU (QK^T + PV = 2 x 128x128x128 = 1024 clk of
//     tcgen05.mma vs 128x128 exp2 = 1024 clk), so MMA issue cost matters: the issuing warp runs its loop converged and
//     only the tcgen05 instructions sit under elect.sync, which lets descriptors live in uniform registers;
//   * TMEM is full: S0 [0,128) S1 [128,256) O0 [256,384) O1 [384,512);
//   * smem: Q 2 x 32 KB | K/V ring 3 x 32 KB | P 2 x 32 KB (reused as output staging) = 224 KB.  A tile of 128 rows x
//     128 halfs is two 128B-swizzled atoms (dims 0-63 | 64-127) 16 KB apart.
#include "common.cuh"
#include "host_util.h"
#include <stdlib.h>

namespace b200 {

struct Attn128Params {
  int B, H, Lq, Lk;
  int BKV, n_kv, q_tiles;  // q_tiles: 256-query CTA tiles
  float scale_log2;
  void* O;
  long long o_stride_b, o_stride_l;
  uint32_t idesc_qk, idesc_pv;
};

namespace a128 {

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

}  // namespace a128
using namespace a128;

static constexpr int kAtom = 128 * 128;       // bytes of one 128-row x 64-half swizzle atom
static constexpr int kTile128 = 2 * kAtom;    // 128 rows x 128 halfs
static constexpr int kRing128 = 3;
static constexpr float kRescale128 = 8.0f;    // log2(256)
#ifndef B200_ATTN_POLY_MASK
#define B200_ATTN_POLY_MASK 0x10
#endif
static constexpr unsigned kPolyMask = B200_ATTN_POLY_MASK;  // elements (i mod 8) whose exp2 runs on the FMA pipe

template <bool BF16>
__global__ void __launch_bounds__(384, 1)
attn128_kernel(const __grid_constant__ CUtensorMap mapQ, const __grid_constant__ CUtensorMap mapK,
               const __grid_constant__ CUtensorMap mapV, const Attn128Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t q_smem = base;                                // 2 tiles
  const uint32_t ring_smem = base + 2 * kTile128;              // 3 K/V tiles
  const uint32_t p_smem = ring_smem + kRing128 * kTile128;     // 2 P tiles (atoms by 64 keys)
  const uint32_t bar_base = p_smem + 2 * kTile128;
  const uint32_t q_full = bar_base;
  auto ring_full = [&](int i) { return bar_base + 8u * (1 + i); };
  auto ring_empty = [&](int i) { return bar_base + 8u * (1 + kRing128 + i); };
  auto s_full = [&](int t) { return bar_base + 8u * (1 + 2 * kRing128 + t); };
  auto p_full = [&](int t) { return bar_base + 8u * (3 + 2 * kRing128 + t); };
  auto pv_done = [&](int t) { return bar_base + 8u * (5 + 2 * kRing128 + t); };
  auto s_cons = [&](int t) { return bar_base + 8u * (7 + 2 * kRing128 + t); };
  const uint32_t tmem_slot = bar_base + 8u * (9 + 2 * kRing128);

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
    for (int i = 0; i < kRing128; ++i) {
      mbar_init(ring_full(i), 1);
      mbar_init(ring_empty(i), 1);
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
  asm volatile("setmaxnreg.dec.sync.aligned.u32 72;");  // 4 x (168 - 72) released = 8 x (216 - 168) taken below
  if (warp == 0 && lane == 0) {
    // ------------------------------------------------------------------ TMA producer
    mbar_expect_tx(q_full, 2u * kTile128);
#pragma unroll
    for (int t = 0; t < 2; ++t) {
      tma_load_3d(q_smem + t * kTile128, &mapQ, q_full, h * 128, q0 + t * 128, b);
      tma_load_3d(q_smem + t * kTile128 + kAtom, &mapQ, q_full, h * 128 + 64, q0 + t * 128, b);
    }
    const uint32_t kv_bytes = (uint32_t)BKV * 256u;
    for (int idx = 0; idx < 2 * n_kv; ++idx) {  // even: K_{idx/2}, odd: V_{idx/2}
      const int slot = idx % kRing128;
      const uint32_t phase = (uint32_t)(idx / kRing128) & 1u;
      mbar_wait(ring_empty(slot), phase ^ 1u);
      mbar_expect_tx(ring_full(slot), kv_bytes);
      const CUtensorMap* m = (idx & 1) ? &mapV : &mapK;
      tma_load_3d(ring_smem + slot * kTile128, m, ring_full(slot), h * 128, (idx >> 1) * BKV, b);
      tma_load_3d(ring_smem + slot * kTile128 + kAtom, m, ring_full(slot), h * 128 + 64, (idx >> 1) * BKV, b);
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer: the warp stays converged, one
    // elected lane issues; every operand below is warp-uniform
    const uint64_t qdesc0 = make_smem_desc_sw128(q_smem, 0, 1024);
    const uint64_t kdesc0 = make_smem_desc_sw128(ring_smem, 0, 1024);
    const uint64_t vdesc0 = make_smem_desc_sw128(ring_smem, kAtom, 1024);  // MN-major V: 64-dim atoms kAtom apart
    const uint64_t pdesc0 = make_smem_desc_sw128(p_smem, 0, 1024);
    const uint32_t idesc_qk = p.idesc_qk, idesc_pv = p.idesc_pv;
    auto wait_full = [&](int idx) {
      mbar_wait(ring_full(idx % kRing128), (uint32_t)(idx / kRing128) & 1u);
      tc_fence_after();
    };
    auto issue_qk = [&](int idx, int t) {  // S_t = Q_t K^T, 8 k-steps of 16 over the 128 dims
      const uint64_t kd = kdesc0 + (uint64_t)((idx % kRing128) * (kTile128 >> 4));
      const uint64_t qd = qdesc0 + (uint64_t)(t * (kTile128 >> 4));
      const uint32_t s_tmem = tmem_base + (uint32_t)t * 128u;
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const uint64_t off = (uint64_t)((k >> 2) * (kAtom >> 4) + (k & 3) * 2);
          umma_f16(s_tmem, qd + off, kd + off, idesc_qk, k != 0 ? 1u : 0u);
        }
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
      if (j + 1 < n_kv) {  // QK_{j+1} as soon as the softmax threads hold S_j in registers
        wait_full(kidx);
        for (int t = 0; t < 2; ++t) {
          mbar_wait(s_cons(t), (uint32_t)j & 1u);
          tc_fence_after();
          issue_qk(kidx, t);
        }
        if (elect_one()) umma_commit(ring_empty(kidx % kRing128));
        __syncwarp();
      }
      wait_full(vidx);
      const uint64_t vd = vdesc0 + (uint64_t)((vidx % kRing128) * (kTile128 >> 4));
      for (int t = 0; t < 2; ++t) {
        mbar_wait(p_full(t), (uint32_t)j & 1u);
        tc_fence_after();
        const uint64_t pd = pdesc0 + (uint64_t)(t * (kTile128 >> 4));
        const uint32_t o_tmem = tmem_base + 256u + (uint32_t)t * 128u;
        const uint32_t acc0 = j != 0 ? 1u : 0u;
        if (elect_one()) {
          if (ksteps == 8) {
#pragma unroll
            for (int kk = 0; kk < 8; ++kk)
              umma_f16(o_tmem, pd + (uint64_t)((kk >> 2) * (kAtom >> 4) + (kk & 3) * 2), vd + (uint64_t)(kk * (2048 >> 4)),
                       idesc_pv, kk ? 1u : acc0);
          } else {
            for (int kk = 0; kk < ksteps; ++kk)
              umma_f16(o_tmem, pd + (uint64_t)((kk >> 2) * (kAtom >> 4) + (kk & 3) * 2), vd + (uint64_t)(kk * (2048 >> 4)),
                       idesc_pv, kk ? 1u : acc0);
          }
          umma_commit(pv_done(t));
        }
        __syncwarp();
      }
      if (elect_one()) umma_commit(ring_empty(vidx % kRing128));
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
    const uint32_t o_addr = tmem_base + 256u + (uint32_t)t * 128u + lane_addr;
    const uint32_t p_tile = p_smem + (uint32_t)t * kTile128;
    const uint32_t p_row = p_tile + (uint32_t)r * 128u;
    const uint32_t sw = (uint32_t)(r & 7);
    const float sl2 = p.scale_log2;
    float m_ref = -INFINITY, l_run = 0.f;
#ifndef B200_ATTN_NO_TURNS
    if (t == 1) named_bar_arrive(2, 256);  // warpgroup 0 takes the first turn on the MUFU
#endif

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
      tc_fence_before();
      mbar_arrive(s_cons(t));
      const bool full_blk = nvalid == 128;
      float mx = -INFINITY;
      if (full_blk) {
#pragma unroll
        for (int i = 0; i < 128; i += 2) mx = fmaxf(mx, fmaxf(__uint_as_float(v[i]), __uint_as_float(v[i + 1])));
      } else {
#pragma unroll
        for (int i = 0; i < 128; ++i)
          if (i < nvalid) mx = fmaxf(mx, __uint_as_float(v[i]));
      }
      const float m_blk = mx * sl2;
      // PV_{j-1} must have retired before O is rescaled and before P is overwritten
      if (j > 0) {
        mbar_wait(pv_done(t), (uint32_t)(j - 1) & 1u);
        tc_fence_after();
      }
      if (j == 0) {
        m_ref = m_blk;
      } else {
        const bool need = m_blk > m_ref + kRescale128;
        if (__any_sync(0xffffffffu, need)) {
          const float alpha = need ? ex2a(m_ref - m_blk) : 1.0f;
#pragma unroll 1
          for (int c = 0; c < 128; c += 32) {
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
#ifndef B200_ATTN_NO_TURNS
      named_bar_sync(2 + t, 256);
#endif
      float rs0 = 0.f, rs1 = 0.f;
      const float nm = -m_ref;
      if (full_blk) {
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
          const uint32_t addr = p_row + (uint32_t)(c >> 6) * kAtom + (((((uint32_t)c & 63u) >> 3) ^ sw) << 4);
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
            const uint32_t addr = p_row + (uint32_t)(c >> 6) * kAtom + (((((uint32_t)c & 63u) >> 3) ^ sw) << 4);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(pack2<BF16>(pe[0], pe[1])),
                         "r"(pack2<BF16>(pe[2], pe[3])), "r"(pack2<BF16>(pe[4], pe[5])), "r"(pack2<BF16>(pe[6], pe[7])));
          }
        }
      }
#ifndef B200_ATTN_NO_TURNS
      if (!(t == 1 && j == n_kv - 1)) named_bar_arrive(3 - t, 256);
#endif
      l_run += rs0 + rs1;
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(p_full(t));
    }

    // ---- output: O_t / l -> fp16/bf16 -> warp-private staging (this tile's P buffer is free now) -> coalesced stores
    mbar_wait(pv_done(t), (uint32_t)(n_kv - 1) & 1u);
    tc_fence_after();
    const float inv = 1.0f / l_run;
    const uint32_t stg = p_tile + (uint32_t)quad * 8192u;  // 32 rows x 256 B per warp
    __syncwarp();
#pragma unroll 1
    for (int c = 0; c < 128; c += 32) {
      uint32_t w[32];
      tmem_ld_32x32(o_addr + (uint32_t)c, w);
      tmem_ld_wait();
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const uint32_t o0 = pack2<BF16>(__uint_as_float(w[g * 8 + 0]) * inv, __uint_as_float(w[g * 8 + 1]) * inv);
        const uint32_t o1 = pack2<BF16>(__uint_as_float(w[g * 8 + 2]) * inv, __uint_as_float(w[g * 8 + 3]) * inv);
        const uint32_t o2 = pack2<BF16>(__uint_as_float(w[g * 8 + 4]) * inv, __uint_as_float(w[g * 8 + 5]) * inv);
        const uint32_t o3 = pack2<BF16>(__uint_as_float(w[g * 8 + 6]) * inv, __uint_as_float(w[g * 8 + 7]) * inv);
        const uint32_t chunk = (uint32_t)(c >> 3) + (uint32_t)g;  // 16B chunk index inside the 256B row
        const uint32_t addr = stg + (uint32_t)lane * 256u + ((chunk ^ (uint32_t)(lane & 7)) << 4);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(o0), "r"(o1), "r"(o2), "r"(o3) : "memory");
      }
    }
    __syncwarp();
    {
      const int piece = lane & 15;
      char* obase = reinterpret_cast<char*>(p.O) + ((size_t)b * p.o_stride_b + (size_t)h * 128) * 2;
#pragma unroll
      for (int jj = 0; jj < 16; ++jj) {
        const int rl = jj * 2 + (lane >> 4);
        const uint32_t addr = stg + (uint32_t)rl * 256u + ((((uint32_t)piece) ^ (uint32_t)(rl & 7)) << 4);
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
static int launch_attn128(const CUtensorMap& mQ, const CUtensorMap& mK, const CUtensorMap& mV, const Attn128Params& p,
                          cudaStream_t stream) {
  const size_t smem = (size_t)kTile128 * (2 + kRing128 + 2) + 1024 + 256;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(attn128_kernel<BF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      set_error("attention128: smem attr: %s", cudaGetErrorString(e));
      return B200_ECUDA;
    }
    attr_done = true;
  }
  const int grid = p.q_tiles * p.H * p.B;
  {
    cudaError_t e = launch_pdl(attn128_kernel<BF16>, dim3(grid), dim3(384), smem, stream, 1, mQ, mK, mV, p);
    if (e != cudaSuccess) {
      set_error("attention128: launch failed: %s", cudaGetErrorString(e));
      return B200_ECUDA;
    }
  }
  B200_CHECK_LAUNCH("attention128");
  return B200_OK;
}

int attention128s_dispatch(const void* q, const void* k, const void* v, void* o, const b200_attn_desc* d, cudaStream_t st);

// called from b200_attention (attention.cu) for Dh == 128.  B200_ATTN128_VER: 2 (default) = the small-CTA kernel of
// attention64s.cu built for Dh = 128 (one query tile per CTA, 64-key blocks, P in TMEM, two CTAs per SM), 0 = this file's
// kernel (two tiles per CTA, one CTA per SM, P through shared memory), kept for A/B.
int attention128_dispatch(const void* q, const void* k, const void* v, void* o, const b200_attn_desc* d, cudaStream_t st) {
  {
    static int ver = -1;
    if (ver < 0) {
      const char* e = getenv("B200_ATTN128_VER");
      ver = e ? atoi(e) : 2;
    }
    if (ver == 2) return attention128s_dispatch(q, k, v, o, d, st);
  }
  Attn128Params p;
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
  p.idesc_pv = make_idesc_f16(128, 128, bf, false, true);
  const uint64_t cols = (uint64_t)d->H * 128;
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
  return bf ? launch_attn128<true>(mQ, mK, mV, p, st) : launch_attn128<false>(mQ, mK, mV, p, st);
}

}  // namespace b200
