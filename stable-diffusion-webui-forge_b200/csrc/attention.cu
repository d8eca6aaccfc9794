// b200forge — multi-head attention forward for sm_100a (FlashAttention-style, tcgen05 + TMEM + TMA).
//
//   O[b, q, h, :] = softmax_k( Q[b, q, h, :] . K[b, k, h, :] * scale ) V[b, k, h, :]
//
// One CTA per (batch, head, 128-query tile); 192 threads:
//   warp 0     TMA producer: Q tile once, then K_j / V_j tiles through a 3-slot ring
//   warp 1     MMA issuer + TMEM owner:  S = Q K_j^T  (M=128, N=BKV, K=Dh)  into TMEM columns [0,128)
//                                        O_j = P_j V_j (M=128, N=Dh, K=BKV)  into TMEM columns [128,128+Dh)
//              V is consumed straight from its [key, Dh] layout as an MN-major B operand (no transpose).
//   warps 2-5  softmax: thread = query row. Reads S from TMEM, online max / exp2 / row-sum in fp32,
//              writes P (fp16/bf16) into shared memory in the 128B-swizzled K-major layout the MMA expects,
//              then folds O_j into its fp32 register accumulator with the running rescale.
// At Dh = 64 two CTAs are co-resident per SM (99 KB smem, 256 TMEM columns each) so one CTA's softmax
// overlaps the other's MMAs; QK_{j+1} is also issued while softmax_j's O fold is still running.
#include "common.cuh"
#include "host_util.h"
#include <stdlib.h>

namespace b200 {

struct AttnKParams {
  int B, H, Lq, Lk;
  int BKV, n_kv, q_tiles;
  float scale_log2;
  void* O;
  long long o_stride_b, o_stride_l;
  uint32_t idesc_qk, idesc_pv;
};

__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

static constexpr int kAtomBytes = 128 * 128;  // 128 rows x 64 halfs, one swizzle-128B "atom" column
static constexpr int kRing = 3;

template <int DH, bool BF16>
__global__ void __launch_bounds__(192, DH == 64 ? 2 : 1)
attn_kernel(const __grid_constant__ CUtensorMap mapQ, const __grid_constant__ CUtensorMap mapK,
            const __grid_constant__ CUtensorMap mapV, const AttnKParams p) {
  constexpr int NA = DH / 64;                 // 64-wide atoms along the head dim
  constexpr int kTileBytes = NA * kAtomBytes; // one Q / K / V tile
  constexpr uint32_t kTmemCols = 256;
  constexpr uint32_t kOCol = 128;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t q_smem = base;
  const uint32_t ring_smem = base + kTileBytes;
  const uint32_t p_smem = ring_smem + kRing * kTileBytes;
  const uint32_t bar_base = p_smem + 2 * kAtomBytes;
  // barriers: q_full, ring_full[3], ring_empty[3], s_full, p_full, o_full, o_free ; then tmem slot
  const uint32_t q_full = bar_base;
  auto ring_full = [&](int i) { return bar_base + 8u * (1 + i); };
  auto ring_empty = [&](int i) { return bar_base + 8u * (1 + kRing + i); };
  const uint32_t s_full = bar_base + 8u * (1 + 2 * kRing);
  const uint32_t p_full = s_full + 8u;
  const uint32_t o_full = s_full + 16u;
  const uint32_t o_free = s_full + 24u;
  const uint32_t tmem_slot = s_full + 32u;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int qt = blockIdx.x % p.q_tiles;
  const int h = (blockIdx.x / p.q_tiles) % p.H;
  const int b = blockIdx.x / (p.q_tiles * p.H);
  const int q0 = qt * 128;
  const int BKV = p.BKV;
  const int n_kv = p.n_kv;

  pdl_launch_dependents();
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&mapQ);
    tma_prefetch_desc(&mapK);
    tma_prefetch_desc(&mapV);
    mbar_init(q_full, 1);
    for (int i = 0; i < kRing; ++i) {
      mbar_init(ring_full(i), 1);
      mbar_init(ring_empty(i), 1);
    }
    mbar_init(s_full, 1);
    mbar_init(p_full, 128);
    mbar_init(o_full, 1);
    mbar_init(o_free, 128);
    mbar_fence_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  pdl_wait();  // set-up done; q / k / v are the predecessor's output

  if (warp == 0 && lane == 0) {
    // ------------------------------------------------------------------ TMA producer
    mbar_expect_tx(q_full, (uint32_t)kTileBytes);
#pragma unroll
    for (int a = 0; a < NA; ++a) tma_load_3d(q_smem + a * kAtomBytes, &mapQ, q_full, h * DH + a * 64, q0, b);
    const uint32_t kv_bytes = (uint32_t)(NA * BKV * 128);
    int slot = 0;
    uint32_t phase = 0;
    for (int t = 0; t < 2 * n_kv; ++t) {  // t even: K_{t/2}, t odd: V_{t/2}
      const int j = t >> 1;
      mbar_wait(ring_empty(slot), phase ^ 1u);
      mbar_expect_tx(ring_full(slot), kv_bytes);
      const CUtensorMap* m = (t & 1) ? &mapV : &mapK;
#pragma unroll
      for (int a = 0; a < NA; ++a)
        tma_load_3d(ring_smem + slot * kTileBytes + a * kAtomBytes, m, ring_full(slot), h * DH + a * 64, j * BKV, b);
      if (++slot == kRing) { slot = 0; phase ^= 1u; }
    }
  } else if (warp == 1 && lane == 0) {
    // ------------------------------------------------------------------ MMA issuer
    const uint32_t s_tmem = tmem_base;
    const uint32_t o_tmem = tmem_base + kOCol;
    int slot = 0;
    uint32_t phase = 0;
    auto issue_qk = [&]() {
      mbar_wait(ring_full(slot), phase);
      tc_fence_after();
      const uint32_t k_smem = ring_smem + slot * kTileBytes;
#pragma unroll
      for (int k = 0; k < DH / 16; ++k) {
        const uint32_t off = (uint32_t)(k >> 2) * kAtomBytes + (uint32_t)(k & 3) * 32u;
        umma_f16(s_tmem, make_smem_desc_sw128(q_smem + off, 0, 1024), make_smem_desc_sw128(k_smem + off, 0, 1024),
                 p.idesc_qk, k != 0 ? 1u : 0u);
      }
      umma_commit(ring_empty(slot));  // K tile free once the QK MMAs retire
      umma_commit(s_full);
      if (++slot == kRing) { slot = 0; phase ^= 1u; }
    };
    mbar_wait(q_full, 0);
    issue_qk();
    for (int j = 0; j < n_kv; ++j) {
      mbar_wait(p_full, (uint32_t)j & 1u);         // P_j staged in smem, S_j fully consumed
      mbar_wait(ring_full(slot), phase);            // V_j landed
      if (j > 0) mbar_wait(o_free, (uint32_t)(j - 1) & 1u);  // O_{j-1} folded into registers
      tc_fence_after();
      const uint32_t v_smem = ring_smem + slot * kTileBytes;
      const int ksteps = BKV >> 4;
      for (int kk = 0; kk < ksteps; ++kk) {
        const uint64_t adesc =
            make_smem_desc_sw128(p_smem + (uint32_t)(kk >> 2) * kAtomBytes + (uint32_t)(kk & 3) * 32u, 0, 1024);
        // MN-major B: [key][64 dh] atoms; 16 keys = 2048 bytes; next dh atom kAtomBytes away
        const uint64_t bdesc = make_smem_desc_sw128(v_smem + (uint32_t)kk * 2048u, kAtomBytes, 1024);
        umma_f16(o_tmem, adesc, bdesc, p.idesc_pv, kk != 0 ? 1u : 0u);
      }
      umma_commit(ring_empty(slot));  // V tile free
      umma_commit(o_full);
      if (++slot == kRing) { slot = 0; phase ^= 1u; }
      if (j + 1 < n_kv) issue_qk();
    }
  } else if (warp >= 2) {
    // ------------------------------------------------------------------ softmax + output
    const int quad = warp & 3;  // TMEM lane quadrant accessible to this warp
    const int r = quad * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
    const uint32_t s_addr = tmem_base + lane_addr;
    const uint32_t o_addr = tmem_base + kOCol + lane_addr;
    const uint32_t p_row = p_smem + (uint32_t)r * 128u;
    const uint32_t sw = (uint32_t)(r & 7);
    float o_acc[DH];
#pragma unroll
    for (int i = 0; i < DH; ++i) o_acc[i] = 0.f;
    float m_run = -INFINITY, l_run = 0.f;
    const float sl2 = p.scale_log2;

    for (int j = 0; j < n_kv; ++j) {
      int nvalid = p.Lk - j * BKV;
      if (nvalid > BKV) nvalid = BKV;
      mbar_wait(s_full, (uint32_t)j & 1u);
      tc_fence_after();
      // pass 1: row max
      float mx = -INFINITY;
      for (int c = 0; c < BKV; c += 32) {
        uint32_t v[32];
        tmem_ld_32x32(s_addr + (uint32_t)c, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (c + i < nvalid) mx = fmaxf(mx, __uint_as_float(v[i]));
      }
      const float m_new = fmaxf(m_run, mx * sl2);
      const float alpha = ex2f(m_run - m_new);
      // pass 2: p = exp2(s*scale - m), row sum, stage P
      float rs = 0.f;
      for (int c = 0; c < BKV; c += 32) {
        uint32_t v[32];
        tmem_ld_32x32(s_addr + (uint32_t)c, v);
        tmem_ld_wait();
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          float p0 = (c + i < nvalid) ? ex2f(fmaf(__uint_as_float(v[i]), sl2, -m_new)) : 0.f;
          float p1 = (c + i + 1 < nvalid) ? ex2f(fmaf(__uint_as_float(v[i + 1]), sl2, -m_new)) : 0.f;
          rs += p0 + p1;
          pk[i >> 1] = pack2<BF16>(p0, p1);
        }
        const uint32_t atom_off = (uint32_t)(c >> 6) * kAtomBytes;
        const uint32_t chunk0 = (uint32_t)(c & 63) >> 3;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          if (c + g * 8 < BKV) {
            const uint32_t addr = p_row + atom_off + (((chunk0 + g) ^ sw) << 4);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(pk[g * 4]), "r"(pk[g * 4 + 1]),
                         "r"(pk[g * 4 + 2]), "r"(pk[g * 4 + 3])
                         : "memory");
          }
        }
      }
      l_run = l_run * alpha + rs;
      m_run = m_new;
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(p_full);
      // fold O_j
      mbar_wait(o_full, (uint32_t)j & 1u);
      tc_fence_after();
#pragma unroll
      for (int c = 0; c < DH; c += 32) {
        uint32_t v[32];
        tmem_ld_32x32(o_addr + (uint32_t)c, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) o_acc[c + i] = fmaf(o_acc[c + i], alpha, __uint_as_float(v[i]));
      }
      tc_fence_before();
      mbar_arrive(o_free);
    }
    const int q = q0 + r;
    if (q < p.Lq) {
      const float inv = 1.0f / l_run;
      char* dst = reinterpret_cast<char*>(p.O) +
                  ((size_t)b * p.o_stride_b + (size_t)q * p.o_stride_l + (size_t)h * DH) * 2;
#pragma unroll
      for (int g = 0; g < DH / 8; ++g) {
        uint4 o;
        o.x = pack2<BF16>(o_acc[g * 8 + 0] * inv, o_acc[g * 8 + 1] * inv);
        o.y = pack2<BF16>(o_acc[g * 8 + 2] * inv, o_acc[g * 8 + 3] * inv);
        o.z = pack2<BF16>(o_acc[g * 8 + 4] * inv, o_acc[g * 8 + 5] * inv);
        o.w = pack2<BF16>(o_acc[g * 8 + 6] * inv, o_acc[g * 8 + 7] * inv);
        *reinterpret_cast<uint4*>(dst + g * 16) = o;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

template <int DH, bool BF16>
static int launch_attn(const CUtensorMap& mQ, const CUtensorMap& mK, const CUtensorMap& mV, const AttnKParams& p,
                       cudaStream_t stream) {
  constexpr int NA = DH / 64;
  const size_t smem = (size_t)NA * kAtomBytes * (1 + kRing) + 2 * kAtomBytes + 1024 + 128;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(attn_kernel<DH, BF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      set_error("attention: smem attr: %s", cudaGetErrorString(e));
      return B200_ECUDA;
    }
    attr_done = true;
  }
  const int grid = p.q_tiles * p.H * p.B;
  {
    cudaError_t e = launch_pdl(attn_kernel<DH, BF16>, dim3(grid), dim3(192), smem, stream, 1, mQ, mK, mV, p);
    if (e != cudaSuccess) {
      set_error("attention: launch failed: %s", cudaGetErrorString(e));
      return B200_ECUDA;
    }
  }
  B200_CHECK_LAUNCH("attention");
  return B200_OK;
}

}  // namespace b200

namespace b200 {
int attention64_dispatch(const void* q, const void* k, const void* v, void* o, const b200_attn_desc* d, cudaStream_t st);
int attention128_dispatch(const void* q, const void* k, const void* v, void* o, const b200_attn_desc* d, cudaStream_t st);
int attention64s_dispatch(const void* q, const void* k, const void* v, void* o, const b200_attn_desc* d, cudaStream_t st);
}
using namespace b200;

extern "C" int b200_attention(const void* q, const void* k, const void* v, void* o, const b200_attn_desc* d,
                              b200_stream_t s) {
  B200_CHECK_ARG(q && k && v && o && d, "attention: null argument");
  B200_CHECK_ARG(d->B > 0 && d->H > 0 && d->Lq > 0 && d->Lk > 0, "attention: bad shape");
  if (d->Dh != 64 && d->Dh != 128) {
    set_error("attention: head dim %d unsupported (64 or 128)", d->Dh);
    return B200_EUNSUPPORTED;
  }
  B200_CHECK_ARG(d->dtype == B200_F16 || d->dtype == B200_BF16, "attention: dtype");
  B200_CHECK_ARG(d->q_stride_l % 8 == 0 && d->k_stride_l % 8 == 0 && d->v_stride_l % 8 == 0 &&
                     d->o_stride_l % 8 == 0 && d->q_stride_b % 8 == 0 && d->k_stride_b % 8 == 0 &&
                     d->v_stride_b % 8 == 0 && d->o_stride_b % 8 == 0,
                 "attention: strides must be multiples of 8 elements");
  B200_CHECK_ARG((reinterpret_cast<uintptr_t>(o) & 15) == 0, "attention: output not 16-byte aligned");
  static int use_old = -1;
  if (use_old < 0) {
    const char* e = getenv("B200_ATTN_V1");
    use_old = (e && e[0] == '1') ? 1 : 0;
  }
  if (d->Dh == 128 && !use_old && d->Lk > 128) return attention128_dispatch(q, k, v, o, d, static_cast<cudaStream_t>(s));
  if (d->Dh == 64) {
    // a single key block (cross-attention, Lk = 77) is latency- not throughput-bound: the one-tile kernel below keeps
    // two CTAs resident per SM and measured faster there (97 vs 125 us at B=16, H=10, Lq=4096)
    if (!use_old && d->Lk > 128) return attention64_dispatch(q, k, v, o, d, static_cast<cudaStream_t>(s));
    // a single key block (cross-attention, 77 keys) is latency-bound: the small-CTA kernel (3 CTAs / SM) measured 76 / 41 us
    // at B16 H10 Lq4096 / B16 H20 Lq1024 against 95 / 49 for the one-tile kernel below (SDPA 129 / 65); B200_ATTN64_CROSS=o
    // keeps the old routing
    static int cross_small = -1;
    if (cross_small < 0) {
      const char* e = getenv("B200_ATTN64_CROSS");
      cross_small = (e && e[0] == 'o') ? 0 : 1;
    }
    if (!use_old && cross_small) return attention64s_dispatch(q, k, v, o, d, static_cast<cudaStream_t>(s));
  }
  AttnKParams p;
  memset(&p, 0, sizeof(p));
  p.B = d->B;
  p.H = d->H;
  p.Lq = d->Lq;
  p.Lk = d->Lk;
  p.BKV = d->Lk >= 128 ? 128 : ((d->Lk + 15) / 16) * 16;
  p.n_kv = (d->Lk + p.BKV - 1) / p.BKV;
  p.q_tiles = (d->Lq + 127) / 128;
  p.scale_log2 = d->scale * 1.4426950408889634f;
  p.O = o;
  p.o_stride_b = d->o_stride_b;
  p.o_stride_l = d->o_stride_l;
  const bool bf = d->dtype == B200_BF16;
  p.idesc_qk = make_idesc_f16(128, p.BKV, bf, false, false);
  p.idesc_pv = make_idesc_f16(128, d->Dh, bf, false, true);

  const uint64_t cols = (uint64_t)d->H * d->Dh;
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
  cudaStream_t st = static_cast<cudaStream_t>(s);
  if (d->Dh == 64) return bf ? launch_attn<64, true>(mQ, mK, mV, p, st) : launch_attn<64, false>(mQ, mK, mV, p, st);
  return bf ? launch_attn<128, true>(mQ, mK, mV, p, st) : launch_attn<128, false>(mQ, mK, mV, p, st);
}
