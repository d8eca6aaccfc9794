// b200forge — persistent warp-specialised GEMM / implicit-GEMM 3x3 convolution for sm_100a.
//
//   C[M, N] = epilogue( A[M, K] * B[N, K]^T )        fp16 or bf16 operands, fp32 accumulation in TMEM
//
// Two builds of the same kernel:
//   CG = 1  one CTA per SM, 128 x BN tiles (tcgen05.mma cta_group::1, M = 128)
//   CG = 2  CTA pairs (cluster of 2) cooperate on 256 x BN tiles (cta_group::2, M = 256): each CTA loads its own
//           128 A rows and HALF of the B tile, the pair's tensor cores read both halves — 1.5x fewer shared-memory
//           and L2 bytes per FLOP than CG = 1, which is what lifts the kernel off the smem-bandwidth ceiling.
// One CTA per SM, static round-robin over output tiles (BN runtime, multiple of 32, <= 256):
//   warp 0   TMA producer: A tile (128 x 64) and B tile (BN x 64) per k-chunk into a ring of smem stages
//   warp 1   MMA issuer: one thread issues tcgen05.mma (M=128, N=BN, K=16) x4 per stage; tcgen05.commit
//            releases the stage and, after the last chunk, publishes the TMEM accumulator
//   warp 2   TMEM allocator (512 columns = two 256-column accumulator stages)
//   warps 4-7 epilogue: tcgen05.ld the accumulator (thread = row), add bias / time-embedding row /
//            residual, apply SiLU / GELU / GEGLU, convert and store — overlapped with the next tile's MMAs
//
// The A operand is fetched in one of two ways:
//   mode 0  plain row-major [M, K] (optionally the channel concatenation of two matrices)
//   mode 1  3x3 stride-1 pad-1 convolution on NHWC: k-chunk = (64-channel slice, filter tap); the tile of
//           128 output pixels is a (tile_n x tile_h x tile_w) box and the tap shifts the box by (ky-1, kx-1);
//           TMA zero-fills the out-of-image part, which is exactly the convolution's zero padding.
#include "common.cuh"
#include "host_util.h"
#include <stdlib.h>

namespace b200 {

struct GemmKParams {
  int M, N;
  int num_k_chunks;
  int mode;
  int split_chunk;     // mode 0: first chunk served by source 2; mode 1: chunks of source 1 per tap
  int chunks_per_tap;  // mode 1
  int tile_w, tile_h, tile_n, tiles_w, tiles_h;
  int BN, tiles_m, tiles_n, num_stages;
  uint32_t idesc;
  void* C;
  int ldc;
  const void* bias;
  int bias_along_m;
  const void* residual;
  int ldr;
  int n_out;  // output columns (N, or N/2 for GEGLU)
  const void* rowvec;
  int ld_rowvec;
  int rows_per_vec;
  int epilogue;
  // LayerNorm folded into the GEMM (A is the *un-normalised* activation, B is W.diag(gamma)):
  //   y = rstd[m] * (acc - mean[m] * ln_c[n]) + ln_d[n],  ln_c = rowsum(W.diag(gamma)),  ln_d = W.beta (+ bias)
  // Row statistics travel as PARTIALS, one float4 (count, mean, M2 = sum of squared deviations, 0) per (row, N tile,
  // epilogue-warp half) of the producer: written once each (no atomics, no zero-fill, bit-reproducible) and merged by the
  // consumer with the parallel-variance formula, so a large row mean does not cancel (sumsq/K - mean^2 did).
  const float4* ln_stats;  // [ln_parts, M] partials of each A row, produced by the previous GEMM's epilogue
  int ln_parts;
  const float* ln_c;
  const float* ln_d;
  float ln_eps;
  float4* row_stats_out;  // [2 * tiles_n, M]: partial statistics of this GEMM's own output rows
  // two row segments with their own weights (Flux double-stream blocks: txt rows and img rows of one joint
  // [B, L_txt + L_img, C] activation): rows with (m % seg_period) >= seg_split use mapB2 / bias2 / rowvec2.
  int seg_period, seg_split;
  const void* bias2;
  const void* rowvec2;
  // tile order: 0 = consecutive tiles walk M (the B tile is shared by the CTAs running at the same time, A is re-streamed
  // once per N tile), 1 = consecutive tiles walk N (A rows shared, B re-streamed once per group of M tiles).  The host
  // re-streams whichever operand is smaller, i.e. the one that stays in L2.
  int n_fast;
  float alpha;     // EXT: accumulators are multiplied by alpha before anything else (0 = off)
  int rowvec_mul;  // rowvec multiplies (per-sample gate) instead of being added
  int act_col0;    // the activation applies to output columns >= act_col0 only
  // mode 1, generic tiling (FEAT bit 2, experimental): tile_n = 1, tile_w / tile_h powers of two, tiles may overhang the
  // image (loads zero-fill, stores are masked); the epilogue decodes (image, y, x) per row instead of assuming that a
  // tile's 128 pixels are consecutive in NHWC order
  int img_n, img_h, img_w, tile_w_log2;
  // mode 1 + generic tiling, nearest x2 upsample folded into the convolution (b200_conv3x3_up2x): the 3x3 filter on the
  // upsampled image is, per output parity (py, px), a 2x2 filter on the LOW-RES image (taps at dy in {py-1, py}, dx in
  // {px-1, px}; weights pre-summed on the host) — 16 instead of 36 MACs per output pixel and input channel, and the 4x tensor
  // is never materialised.  M tiles enumerate (parity, low-res tile); B rows of parity q start at q * N; output row
  // (img, y, x) of parity (py, px) is pixel (2y + py, 2x + px) of the [img_n, 2 img_h, 2 img_w] output.
  int up2x, tiles_lr;
  // K-split of the LAST, partly filled wave of tiles (persistent kernel, static round-robin: T tiles on U units leave
  // T mod U units busy and the rest idle for a whole tile time — 4.32 waves at M = 16384, N = 1280 cost 5): the T mod U tail
  // tiles are cut into split_S shares of the k-chunk range, one per unit; shares 1.. dump their raw fp32 accumulators into a
  // workspace (column-major 128-row blocks: coalesced for the thread-per-row TMEM layout) and raise a per-tile flag, share 0
  // waits for the flag, adds the partials in a fixed order (bit-reproducible) and runs the normal epilogue.
  // last N tile narrower than BN (N = 640 or 1920 with 256-wide tiles): the MMA runs with N = bn_last (multiple of 32) instead
  // of multiplying zero-filled weight rows, the epilogue walks bn_last columns; a CTA pair splits the bn_last rows evenly
  int bn_last;
  uint32_t idesc_last;
  int split_S;        // 0 = off
  int split_first;    // first tail tile (= T when off): tiles below it are whole
  int split_rem;      // tail tiles
  float* split_ws;    // [split_rem][split_S - 1][CG] blocks of 256 x 128 floats
  int* split_flags;   // [split_rem][CG], self-resetting
};

#ifndef B200_GEMM_SETMAXNREG
#define B200_GEMM_SETMAXNREG 1
#endif
static constexpr int kATileBytes = 128 * 128;  // 128 rows x 64 halfs
// Two epilogue warps per TMEM lane quadrant.  Three (512 threads, setmaxnreg 48 / 152) were measured and are SLOWER on every
// flavour (UNet step 118.3 vs 114.5 ms, LayerNorm-fold N3840 1187 vs 1296 TF/s): the mainloop sits on the shared-memory port
// (TMA fill + MMA operand reads = 128 B/clk) and the epilogue's staging traffic (residual transpose + output transpose,
// ~4 KB each way per 32-column chunk and warp) competes for it — more concurrent epilogue warps take more of it.
static constexpr int kThreads = 384;   // 4 control warps + 8 epilogue warps (two per TMEM lane quadrant)
static constexpr int kEpiThreads = 256;
static constexpr uint32_t kStageBufs = 8;          // one private epilogue staging tile per epilogue warp
static constexpr uint32_t kStageBufBytes = 32 * 64;  // 32 rows x 32 cols (64 B, swizzled)

__device__ __forceinline__ void tmem_st_32x32_g(uint32_t taddr, const uint32_t (&v)[32]) {
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

template <bool BF16>
__device__ __forceinline__ void epi_add_vec8(const void* base, size_t elem_off, float (&x)[8]) {
  uint4 r = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const char*>(base) + elem_off * 2));
  float2 f;
  f = unpack2<BF16>(r.x); x[0] += f.x; x[1] += f.y;
  f = unpack2<BF16>(r.y); x[2] += f.x; x[3] += f.y;
  f = unpack2<BF16>(r.z); x[4] += f.x; x[5] += f.y;
  f = unpack2<BF16>(r.w); x[6] += f.x; x[7] += f.y;
}

template <bool BF16>
__device__ __forceinline__ void add_u4(const uint4& r, float (&x)[8]) {
  float2 f;
  f = unpack2<BF16>(r.x); x[0] += f.x; x[1] += f.y;
  f = unpack2<BF16>(r.y); x[2] += f.x; x[3] += f.y;
  f = unpack2<BF16>(r.z); x[4] += f.x; x[5] += f.y;
  f = unpack2<BF16>(r.w); x[6] += f.x; x[7] += f.y;
}
template <bool BF16>
__device__ __forceinline__ void mul_u4(const uint4& r, float (&x)[8]) {
  float2 f;
  f = unpack2<BF16>(r.x); x[0] *= f.x; x[1] *= f.y;
  f = unpack2<BF16>(r.y); x[2] *= f.x; x[3] *= f.y;
  f = unpack2<BF16>(r.z); x[4] *= f.x; x[5] *= f.y;
  f = unpack2<BF16>(r.w); x[6] *= f.x; x[7] *= f.y;
}
template <bool BF16>
__device__ __forceinline__ void add_smem8(uint32_t addr, float (&x)[8]) {
  uint4 r;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(addr));
  add_u4<BF16>(r, x);
}

__device__ __forceinline__ void lds_f4(uint32_t addr, float (&f)[4]) {
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(f[0]), "=f"(f[1]), "=f"(f[2]), "=f"(f[3]) : "r"(addr));
}
__device__ __forceinline__ void add_smem8_f32(uint32_t addr, float (&x)[8]) {  // x += 8 fp32 values at addr
  float a[4], b[4];
  lds_f4(addr, a);
  lds_f4(addr + 16u, b);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    x[i] += a[i];
    x[4 + i] += b[i];
  }
}

// x[i] = rstd * (x[i] - mean * c[col+i]) + d[col+i]   (c, d: fp32 rows of the current tile in smem)
__device__ __forceinline__ void ln_apply8(uint32_t c_smem, uint32_t d_smem, int col, float mean, float rstd, float (&x)[8]) {
  float cv[8], dv[8];
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(cv[0]), "=f"(cv[1]), "=f"(cv[2]), "=f"(cv[3]) : "r"(c_smem + 4u * col));
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(cv[4]), "=f"(cv[5]), "=f"(cv[6]), "=f"(cv[7]) : "r"(c_smem + 4u * col + 16u));
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(dv[0]), "=f"(dv[1]), "=f"(dv[2]), "=f"(dv[3]) : "r"(d_smem + 4u * col));
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(dv[4]), "=f"(dv[5]), "=f"(dv[6]), "=f"(dv[7]) : "r"(d_smem + 4u * col + 16u));
#pragma unroll
  for (int i = 0; i < 8; ++i) x[i] = fmaf(rstd, fmaf(-mean, cv[i], x[i]), dv[i]);
}

// FEAT selects the epilogue build so that each caller only carries the state it uses (the epilogue sits at the
// 168-register cap):  bit 0 (LNS) LayerNorm folding + output row statistics (UNet transformer blocks),
// bit 1 (EXT) the Flux-path features (row segments with two weight sets, multiplicative rowvec, partial activation,
// tanh GELU), bit 2 (GT) generic convolution tiling for image widths that are neither a power of two nor a multiple of
// 128 (B200_CONV_GENERAL=0 turns it off).  Convolutions and plain linears run the FEAT = 0 build.
// __maxnreg__ rather than __launch_bounds__: with setmaxnreg in the kernel ptxas takes the cap as the LAUNCH-time register
// count and lets the code after setmaxnreg.inc use the raised count (under __launch_bounds__ the 168 stayed a hard cap for the
// whole kernel and the epilogue spilled)
template <bool BF16, int CG, int FEAT>
#if B200_GEMM_SETMAXNREG
__global__ void __maxnreg__(168)
#else
__global__ void __launch_bounds__(kThreads, 1)
#endif
gemm_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapA2,
            const __grid_constant__ CUtensorMap mapB, const __grid_constant__ CUtensorMap mapB2,
            const GemmKParams p) {
  // CG == 2: launched with cluster dims (2,1,1); rank 0 of each pair is the MMA leader.
  constexpr bool LNS = (FEAT & 1) != 0;
  constexpr bool EXT = (FEAT & 2) != 0;
  constexpr bool GT = (FEAT & 4) != 0;
  constexpr bool RV = FEAT != 1;  // the UNet transformer build (LNS alone) carries no per-sample row vector (host-routed)
  const uint32_t cta_rank = (CG == 2) ? cluster_ctarank() : 0u;
  const int unit = (CG == 2) ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;       // pair (or CTA) index
  const int num_units = (CG == 2) ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int S = p.num_stages;
  const int BN = p.BN;
  const uint32_t b_tile_bytes = (uint32_t)(BN / CG) * 128u;  // CG == 2: this CTA holds half of the B tile
  const uint32_t a_base = smem_base;
  const uint32_t b_base = smem_base + (uint32_t)S * kATileBytes;
  const uint32_t stg_base = b_base + (uint32_t)S * b_tile_bytes;  // 1024-aligned: every tile size is a multiple of 1 KB
  const uint32_t bar_base = stg_base + kStageBufs * kStageBufBytes + 1024u + 2048u;  // + bias row + LN c/d rows
  // barrier layout: full[S], empty[S], tmem_full[2], tmem_empty[2], then the TMEM pointer slot
  auto full_bar = [&](int i) { return bar_base + (uint32_t)i * 8u; };
  auto empty_bar = [&](int i) { return bar_base + (uint32_t)(S + i) * 8u; };
  auto tfull_bar = [&](int i) { return bar_base + (uint32_t)(2 * S + i) * 8u; };
  auto tempty_bar = [&](int i) { return bar_base + (uint32_t)(2 * S + 2 + i) * 8u; };
  const uint32_t tmem_slot = bar_base + (uint32_t)(2 * S + 4) * 8u;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  pdl_launch_dependents();  // the next kernel's CTAs may take this SM as soon as this CTA leaves it
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&mapA);
    tma_prefetch_desc(&mapA2);
    tma_prefetch_desc(&mapB);
    tma_prefetch_desc(&mapB2);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < S; ++i) {
      mbar_init(full_bar(i), CG);  // CG == 2: one arrival per CTA's producer, bytes of both CTAs
      mbar_init(empty_bar(i), 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(tfull_bar(i), 1);
      mbar_init(tempty_bar(i), kEpiThreads * CG);
    }
    mbar_fence_init();
  }
  if (warp == 2) {
    if constexpr (CG == 2) {
      tmem_alloc_cg2(tmem_slot, 512);
      tmem_relinquish_cg2();
    } else {
      tmem_alloc(tmem_slot, 512);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  pdl_wait();  // barriers, TMEM and descriptors are ready: from here on the predecessor's output is read

  const int tiles_mu = (p.tiles_m + CG - 1) / CG;  // M tiles per unit: a pair covers two adjacent 128-row tiles
  const int nk = p.num_k_chunks;
  // work item i of this unit: whole tiles round-robin, then at most one share of a K-split tail tile.
  // role 0 = whole tile, 1 = share 0 of a tail tile (reduces the others' partials, runs the epilogue), 2 = partial share
  auto work_item = [&](int i, int& tile, int& kb, int& ke, int& role, int& tt, int& share) -> bool {
    tile = unit + i * num_units;
    kb = 0;
    ke = nk;
    role = 0;
    tt = 0;
    share = 0;
    if (tile < p.split_first) return true;
    if (p.split_S == 0) return false;
    const int j = tile - p.split_first;  // = unit at the first index past the whole tiles (split_first is a multiple of num_units)
    if (j >= p.split_rem * p.split_S) return false;
    share = j / p.split_rem;
    tt = j - share * p.split_rem;
    tile = p.split_first + tt;
    kb = (share * nk) / p.split_S;
    ke = ((share + 1) * nk) / p.split_S;
    role = share == 0 ? 1 : 2;
    return true;
  };

#if B200_GEMM_SETMAXNREG
  // Register split by warpgroup (setmaxnreg is warpgroup-wide; warps 0-3 = control, 4-11 = epilogue): the control warps keep
  // 56 registers, the epilogue warps take 224 (4 x 56 + 8 x 224 = 2016 of the 2048 x 32 the CTA was launched with).  The
  // epilogue of the K <= 1280 GEMMs is a latency chain (TMEM -> math -> staging -> stores) per 32-column chunk at 168 registers
  // with spills; with 224 the next chunk's TMEM load and the residual loads fly under the current chunk's arithmetic.
  // (each count governs the code of ITS branch only: the two roles must not share code after the instruction)
#endif
  if (warp < 4) {
#if B200_GEMM_SETMAXNREG
  asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
#endif
  if (warp == 0 && lane == 0) {
    // ------------------------------------------------------------------ TMA producer
    int stage = 0;
    uint32_t phase = 0;
    const uint32_t tx_bytes = ((uint32_t)kATileBytes + b_tile_bytes) * CG;
    for (int wi = 0;; ++wi) {
      int tile, kb, ke, role, tt, share;
      if (!work_item(wi, tile, kb, ke, role, tt, share)) break;
      const int m_unit = p.n_fast ? tile / p.tiles_n : tile % tiles_mu;
      const int n_blk = p.n_fast ? tile % p.tiles_n : tile / tiles_mu;
      const int m_blk = m_unit * CG + (int)cta_rank;
      // a pair's 256 rows never straddle a segment boundary (host-checked), so the weight set is per tile
      const CUtensorMap* mb = (EXT && p.seg_period && (m_unit * (CG * 128)) % p.seg_period >= p.seg_split) ? &mapB2 : &mapB;
      int cn = 0, ch = 0, cw = 0;
      int ntaps = 9, taps_w = 3, b_row0 = 0;
      if (p.mode == 1) {
        int mlr = m_blk;
        if (p.up2x) {  // (parity, low-res tile): 2x2 taps shifted by the parity, this parity's weight rows
          const int pa = m_blk / p.tiles_lr;
          mlr = m_blk - pa * p.tiles_lr;
          ntaps = 4;
          taps_w = 2;
          ch = pa >> 1;
          cw = pa & 1;
          b_row0 = pa * p.N;
        }
        const int tpi = p.tiles_w * p.tiles_h;
        const int img_grp = mlr / tpi;
        const int rem = mlr - img_grp * tpi;
        cn = img_grp * p.tile_n;
        ch += (rem / p.tiles_w) * p.tile_h;
        cw += (rem % p.tiles_w) * p.tile_w;
      }
      // (64-channel slice, tap) of k-chunk kc, kept incrementally: a division by the run-time tap count per chunk made the
      // single producer thread the bottleneck of the convolutions (+15-29 % on the VAE's, measured)
      int cc = 0, tap = 0, ky = 0, kx = 0;
      if (p.mode == 1) {
        cc = kb / ntaps;
        tap = kb - cc * ntaps;
        ky = tap / taps_w;
        kx = tap - ky * taps_w;
      }
      for (int kc = kb; kc < ke; ++kc) {
        mbar_wait(empty_bar(stage), phase ^ 1u);
        const uint32_t fb = full_bar(stage);
        const uint32_t a_dst = a_base + (uint32_t)stage * kATileBytes;
        const uint32_t b_dst = b_base + (uint32_t)stage * b_tile_bytes;
        int bcol = kc * 64;
        if constexpr (CG == 1) {
          mbar_expect_tx(fb, tx_bytes);
          if (p.mode == 0) {
            if (kc < p.split_chunk) tma_load_2d(a_dst, &mapA, fb, kc * 64, m_blk * 128);
            else tma_load_2d(a_dst, &mapA2, fb, (kc - p.split_chunk) * 64, m_blk * 128);
          } else {
            // k order = (64-channel slice, tap): the nine shifted boxes of one slice are fetched back to back, so
            // they hit L2 whatever the channel count (tap-major order re-streamed the whole activation from HBM nine
            // times once 74 pairs x 256 pixels x C channels outgrew L2).  The weight column follows the tap-major packing.
            bcol = (tap * p.chunks_per_tap + cc) * 64;
            if (cc < p.split_chunk) tma_load_4d(a_dst, &mapA, fb, cc * 64, cw + kx - 1, ch + ky - 1, cn);
            else tma_load_4d(a_dst, &mapA2, fb, (cc - p.split_chunk) * 64, cw + kx - 1, ch + ky - 1, cn);
          }
          tma_load_2d(b_dst, mb, fb, bcol, b_row0 + n_blk * BN);
        } else {
          // both CTAs' loads complete on the leader's barrier; the leader arms it for the pair's bytes
          if (cta_rank == 0) mbar_expect_tx(fb, tx_bytes);
          else mbar_arrive_remote(fb, 0);
          if (p.mode == 0) {
            if (kc < p.split_chunk) tma_load_2d_cg2(a_dst, &mapA, fb, kc * 64, m_blk * 128);
            else tma_load_2d_cg2(a_dst, &mapA2, fb, (kc - p.split_chunk) * 64, m_blk * 128);
          } else {
            bcol = (tap * p.chunks_per_tap + cc) * 64;  // (slice, tap) order, see the single-CTA branch
            if (cc < p.split_chunk) tma_load_4d_cg2(a_dst, &mapA, fb, cc * 64, cw + kx - 1, ch + ky - 1, cn);
            else tma_load_4d_cg2(a_dst, &mapA2, fb, (cc - p.split_chunk) * 64, cw + kx - 1, ch + ky - 1, cn);
          }
          tma_load_2d_cg2(b_dst, mb, fb, bcol, b_row0 + n_blk * BN + (int)cta_rank * ((n_blk == p.tiles_n - 1 ? p.bn_last : BN) / 2));
        }
        if (++stage == S) { stage = 0; phase ^= 1u; }
        if (++kx == taps_w) {
          kx = 0;
          ++ky;
        }
        if (++tap == ntaps) {
          tap = 0;
          ky = 0;
          ++cc;
        }
      }
    }
  } else if (warp == 1 && cta_rank == 0) {
    // ------------------------------------------------------------------ MMA issuer (pair leader only): the warp stays
    // converged and one elected lane issues, so descriptors are computed in uniform registers
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    const uint64_t adesc0 = make_smem_desc_sw128(a_base, 0, 1024);
    const uint64_t bdesc0 = make_smem_desc_sw128(b_base, 0, 1024);
    const uint32_t idesc_full = p.idesc;
    for (;; ++it) {
      int tile, kb, ke, role, tt, share;
      if (!work_item(it, tile, kb, ke, role, tt, share)) break;
      const int acc = it & 1;
      const uint32_t acc_phase = (uint32_t)(it >> 1) & 1u;
      mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)acc * 256u;
      const int n_blk_t = p.n_fast ? tile % p.tiles_n : tile / tiles_mu;
      const uint32_t idesc = (n_blk_t == p.tiles_n - 1) ? p.idesc_last : idesc_full;
      for (int kc = kb; kc < ke; ++kc) {
        mbar_wait(full_bar(stage), phase);
        tc_fence_after();
        const uint64_t adesc = adesc0 + (uint64_t)(((uint32_t)stage * kATileBytes) >> 4);
        const uint64_t bdesc = bdesc0 + (uint64_t)(((uint32_t)stage * b_tile_bytes) >> 4);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            // +32 bytes per 16-element k step inside the 128-byte swizzle atom (encoded >> 4)
            if constexpr (CG == 2)
              umma_f16_cg2(d_tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, ((kc - kb) | k) != 0 ? 1u : 0u);
            else
              umma_f16(d_tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, ((kc - kb) | k) != 0 ? 1u : 0u);
          }
          if constexpr (CG == 2) umma_commit_cg2(empty_bar(stage), 3); else umma_commit(empty_bar(stage));
          if (kc == ke - 1) {
            if constexpr (CG == 2) umma_commit_cg2(tfull_bar(acc), 3); else umma_commit(tfull_bar(acc));
          }
        }
        __syncwarp();
        if (++stage == S) { stage = 0; phase ^= 1u; }
      }
    }
  }
  } else {
#if B200_GEMM_SETMAXNREG
    asm volatile("setmaxnreg.inc.sync.aligned.u32 224;");
#endif
    // ------------------------------------------------------------------ epilogue
    // Per 32-column chunk: TMEM -> registers -> (bias, temb row, activation, residual) -> fp16/bf16 ->
    // warp-private 64B-swizzled staging tile in smem -> coalesced full-sector global stores.
    const int quad = warp & 3;          // the TMEM lane quadrant this warp may access
    const int chalf = (warp - 4) >> 2;  // the two warps of a quadrant take alternate 32-column chunks
    const int r = quad * 32 + lane;
    const uint32_t my_stg = stg_base + (uint32_t)(warp - 4) * 2048u;
    const uint32_t bias_smem = stg_base + kStageBufs * kStageBufBytes;  // 256 halfs
    const uint32_t lnc_smem = bias_smem + 1024u;                        // 256 floats
    const uint32_t lnd_smem = lnc_smem + 1024u;                         // 256 floats
    const bool ln = LNS && p.ln_stats != nullptr;
    float4* const row_stats_out = LNS ? p.row_stats_out : nullptr;
    const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
    const uint32_t sw = (uint32_t)(lane >> 1) & 3u;
    const bool geglu = p.epilogue == B200_EPI_GEGLU;
    const int ncols_out = geglu ? (BN >> 1) : BN;
    for (int it = 0;; ++it) {
      int tile, kb, ke, role, tt, share;
      if (!work_item(it, tile, kb, ke, role, tt, share)) break;
      const int m_unit = p.n_fast ? tile / p.tiles_n : tile % tiles_mu;
      const int n_blk = p.n_fast ? tile % p.tiles_n : tile / tiles_mu;
      const int m_blk = m_unit * CG + (int)cta_rank;
      const int acc = it & 1;
      const uint32_t acc_phase = (uint32_t)(it >> 1) & 1u;
      const int bn_t = (n_blk == p.tiles_n - 1) ? p.bn_last : BN;  // accumulator columns of this tile
      if (role == 2) {
        // ---- partial share of a K-split tail tile: raw accumulators -> workspace (element (col, row) at col * 128 + row: a
        // warp's store of one register is 128 contiguous bytes), release TMEM, publish
        mbar_wait(tfull_bar(acc), acc_phase);
        tc_fence_after();
        float* const blk = p.split_ws + ((size_t)(tt * (p.split_S - 1) + (share - 1)) * CG + cta_rank) * (256 * 128) + r;
        const uint32_t t_addr_p = tmem_base + (uint32_t)acc * 256u + ((uint32_t)(quad * 32) << 16);
        for (int c = chalf * 32; c < bn_t; c += 64) {
          uint32_t v[32];
          tmem_ld_32x32(t_addr_p + (uint32_t)c, v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) blk[(size_t)(c + i) * 128] = __uint_as_float(v[i]);
        }
        tc_fence_before();
        if (CG == 2 && cta_rank != 0) mbar_arrive_remote(tempty_bar(acc), 0);
        else mbar_arrive(tempty_bar(acc));
        __threadfence();
        named_bar_sync(1, kEpiThreads);
        if (threadIdx.x == 128) atomicAdd(p.split_flags + tt * CG + (int)cta_rank, 1);
        continue;
      }
      if (role == 1) {
        // ---- share 0 of a K-split tail tile: before anything else of the epilogue is live in registers, add the other
        // shares' partial accumulators into this CTA's TMEM accumulator (fixed order).  All loads of a chunk are in flight
        // together — a dependent chain of small batches made this pass cost more than the split saved.
        mbar_wait(tfull_bar(acc), acc_phase);
        tc_fence_after();
        if (threadIdx.x == 128) {
          volatile int* f = p.split_flags + tt * CG + (int)cta_rank;
          while (*f < p.split_S - 1) __nanosleep(20);
          *f = 0;  // self-resetting: the next launch's shares only start after this grid has completed
          __threadfence();
        }
        named_bar_sync(1, kEpiThreads);
        const float* const src0 = p.split_ws + ((size_t)(tt * (p.split_S - 1)) * CG + cta_rank) * (256 * 128) + r;
        const float* const src1 = src0 + (size_t)CG * (256 * 128);
        const bool two = p.split_S == 3;
        const uint32_t t_addr_p = tmem_base + (uint32_t)acc * 256u + ((uint32_t)(quad * 32) << 16);
        for (int c = chalf * 32; c < bn_t; c += 64) {
          uint32_t v[32];
          float pa[32], pb[32];
          tmem_ld_32x32(t_addr_p + (uint32_t)c, v);
#pragma unroll
          for (int i = 0; i < 32; ++i) pa[i] = __ldcg(src0 + (size_t)(c + i) * 128);  // L2 loads: written by other SMs
          if (two) {
#pragma unroll
            for (int i = 0; i < 32; ++i) pb[i] = __ldcg(src1 + (size_t)(c + i) * 128);
          }
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            float x = __uint_as_float(v[i]) + pa[i];
            if (two) x += pb[i];
            v[i] = __float_as_uint(x);
          }
          tmem_st_32x32_g(t_addr_p + (uint32_t)c, v);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      }
      int m = m_blk * 128 + r;
      bool row_ok = m < p.M;
      // folded LayerNorm: merge the producer's partial statistics of this row (Chan et al.): fetched before the wait on
      // the accumulator so that the loads fly under the tile's MMAs
      float ln_mean = 0.f, ln_rstd = 1.f;
      if (ln && row_ok) {
        const float4* sp = p.ln_stats + m;  // [part][row]: a warp's 32 rows are 512 contiguous bytes per part
        float cnt = 0.f, wsum = 0.f;
        for (int i = 0; i < p.ln_parts; ++i) {
          const float4 q = __ldg(sp + (size_t)i * p.M);
          cnt += q.x;
          wsum = fmaf(q.x, q.y, wsum);
        }
        ln_mean = wsum / cnt;
        float m2 = 0.f;
        for (int i = 0; i < p.ln_parts; ++i) {
          const float4 q = __ldg(sp + (size_t)i * p.M);  // second pass hits L1
          const float dm = q.y - ln_mean;
          m2 += fmaf(q.x * dm, dm, q.z);
        }
        ln_rstd = rsqrtf(m2 / cnt + p.ln_eps);
      }
      const uint32_t t_addr = tmem_base + (uint32_t)acc * 256u + lane_addr;
      int gt_img = 0, gt_y0 = 0, gt_x0 = 0;  // GT: image and origin of this tile
      int gt_py = 0, gt_px = 0;              // up2x: output parity of this tile
      // GT: output row of pixel (img, y, x); with the folded upsample the tile's pixels land on one parity of the 2x grid
      auto gt_row = [&](int img, int y, int x, int py, int px) -> int {
        if (p.up2x) return (img * (2 * p.img_h) + 2 * y + py) * (2 * p.img_w) + 2 * x + px;
        return (img * p.img_h + y) * p.img_w + x;
      };
      if constexpr (GT) {
        int mlr = m_blk;
        if (p.up2x) {
          const int pa = m_blk / p.tiles_lr;
          mlr = m_blk - pa * p.tiles_lr;
          gt_py = pa >> 1;
          gt_px = pa & 1;
        }
        const int tpi = p.tiles_w * p.tiles_h;
        gt_img = mlr / tpi;
        const int rem = mlr - gt_img * tpi;
        const int ty = rem / p.tiles_w;
        gt_y0 = ty * p.tile_h;
        gt_x0 = (rem - ty * p.tiles_w) * p.tile_w;
        const int y = gt_y0 + (r >> p.tile_w_log2), x = gt_x0 + (r & (p.tile_w - 1));
        row_ok = gt_img < p.img_n && y < p.img_h && x < p.img_w;
        m = gt_row(gt_img, y, x, gt_py, gt_px);
      }
      const int n0 = n_blk * BN;            // first accumulator column of this tile (weight row index)
      const int out_n0 = n_blk * ncols_out; // first output column
      const bool seg1 = EXT && p.seg_period && (m_unit * (CG * 128)) % p.seg_period >= p.seg_split;
      const void* const bias_p = seg1 ? p.bias2 : p.bias;
      const void* const rowvec_p = seg1 ? p.rowvec2 : p.rowvec;
      const bool act_on = !EXT || n0 >= p.act_col0;
      float bias_m = 0.f;
      if (bias_p && p.bias_along_m && row_ok) bias_m = ld1<BF16>(bias_p, m);
      const size_t rv_off = (rowvec_p && row_ok) ? (size_t)(m / p.rows_per_vec) * p.ld_rowvec : 0;
      // per-column bias of this tile -> smem once (the per-chunk global loads were the epilogue's critical path)
      const bool col_bias = bias_p && !p.bias_along_m;
      // output row statistics: packed sums of (x - pivot), (x - pivot)^2 with the thread's first value as the pivot
      float st_piv = 0.f, st_cnt = 0.f;
      f32x2_t st_s2 = pk2(0.f, 0.f), st_q2 = pk2(0.f, 0.f);
      if (col_bias || ln) {
        // this tile's bias / LN rows -> smem, all of it BEFORE the wait on the accumulator: the global loads are issued first,
        // the first barrier (every warp is done with the previous tile's rows) then only guards the st.shared
        const int e0 = (int)(threadIdx.x - 128) * 8;
        const int e = (int)threadIdx.x - 128;
        float cv = 0.f, dv = 0.f;
        if (ln && e < BN && n0 + e < p.N) {
          cv = __ldg(p.ln_c + n0 + e);
          dv = __ldg(p.ln_d + n0 + e);
        }
        uint4 bv = make_uint4(0, 0, 0, 0);
        if (col_bias && e0 < BN && n0 + e0 < p.N)
          bv = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const char*>(bias_p) + (size_t)(n0 + e0) * 2));
        named_bar_sync(1, kEpiThreads);
        if (ln && e < BN) {
          asm volatile("st.shared.f32 [%0], %1;" ::"r"(lnc_smem + 4u * e), "f"(cv) : "memory");
          asm volatile("st.shared.f32 [%0], %1;" ::"r"(lnd_smem + 4u * e), "f"(dv) : "memory");
        }
        if (col_bias && e0 < BN) {  // the bias row is kept in fp32: the chunks add it with packed adds, no conversions
          const float2 b0 = unpack2<BF16>(bv.x), b1 = unpack2<BF16>(bv.y), b2 = unpack2<BF16>(bv.z), b3 = unpack2<BF16>(bv.w);
          asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(bias_smem + (uint32_t)e0 * 4u), "f"(b0.x), "f"(b0.y),
                       "f"(b1.x), "f"(b1.y)
                       : "memory");
          asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(bias_smem + (uint32_t)e0 * 4u + 16u), "f"(b2.x), "f"(b2.y),
                       "f"(b3.x), "f"(b3.y)
                       : "memory");
        }
        named_bar_sync(1, kEpiThreads);
      }
      // global row of local row rl of this warp's quadrant (the coalesced residual loads and the coalesced stores walk rows
      // as rl = j*8 + lane/4 with 16-byte pieces lane%4)
      auto row_of = [&](int rl, int& mm) -> bool {
        mm = m_blk * 128 + quad * 32 + rl;
        bool ok = mm < p.M;
        if constexpr (GT) {
          const int rr = quad * 32 + rl;
          const int y = gt_y0 + (rr >> p.tile_w_log2), x = gt_x0 + (rr & (p.tile_w - 1));
          ok = gt_img < p.img_n && y < p.img_h && x < p.img_w;
          mm = gt_row(gt_img, y, x, gt_py, gt_px);
        }
        return ok;
      };
      const bool any_rs = p.residual != nullptr;  // warp-uniform
      const int piece = lane & 3;
      // residual: COALESCED — each load instruction covers 8 rows x 64 contiguous bytes (the thread-per-row pattern of the
      // TMEM layout touched 32 rows per instruction: 32 L1 wavefronts each, which saturated the load pipe on the K <= 1280
      // shapes); the tile is transposed to thread-per-row through the warp's staging buffer below.  The loads run one chunk
      // ahead of the arithmetic (the first chunk's before the wait on the accumulator): a DRAM round trip per chunk in the
      // dependent chain made the K <= 1280 producer GEMMs epilogue-bound.
      auto load_res = [&](int c, uint4 (&dst)[4]) {
        const int col = out_n0 + c + piece * 8;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          int mm;
          const bool ok = row_of(j * 8 + (lane >> 2), mm) && col < p.n_out;
          dst[j] = ok ? *reinterpret_cast<const uint4*>(reinterpret_cast<const char*>(p.residual) + ((size_t)mm * p.ldr + col) * 2)
                      : make_uint4(0, 0, 0, 0);
        }
      };
      uint4 rsc[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) rsc[j] = make_uint4(0, 0, 0, 0);
      if (any_rs) {
        // the residual usually comes from DRAM (written several kernels ago): pull the NEXT tile's lines into L2 now, a whole
        // mainloop ahead of their loads (prefetch instructions hold no registers), and the first tile's at the start
        auto prefetch_res = [&](int t) {
          const int t_mu = p.n_fast ? t / p.tiles_n : t % tiles_mu;
          const int t_nb = p.n_fast ? t % p.tiles_n : t / tiles_mu;
          const int t_mb = t_mu * CG + (int)cta_rank;
          const int e = (int)threadIdx.x - 128;  // 256 epilogue threads: two per tile row
          const int rr = e >> 1;
          int mm = t_mb * 128 + rr;
          bool ok = mm < p.M;
          if constexpr (GT) {
            int t_lr = t_mb, py = 0, px = 0;
            if (p.up2x) {
              const int pa = t_mb / p.tiles_lr;
              t_lr = t_mb - pa * p.tiles_lr;
              py = pa >> 1;
              px = pa & 1;
            }
            const int tpi = p.tiles_w * p.tiles_h;
            const int img = t_lr / tpi;
            const int rem = t_lr - img * tpi;
            const int ty = rem / p.tiles_w;
            const int y = ty * p.tile_h + (rr >> p.tile_w_log2), x = (rem - ty * p.tiles_w) * p.tile_w + (rr & (p.tile_w - 1));
            ok = img < p.img_n && y < p.img_h && x < p.img_w;
            mm = gt_row(img, y, x, py, px);
          }
          if (!ok) return;
          const char* row = reinterpret_cast<const char*>(p.residual) + ((size_t)mm * p.ldr + (size_t)t_nb * ncols_out) * 2;
          const int row_bytes = min(ncols_out, p.n_out - t_nb * ncols_out) * 2;
          for (int off = (e & 1) * 128; off < row_bytes; off += 256)
            asm volatile("prefetch.global.L2 [%0];" ::"l"(row + off));
        };
        if (it == 0) prefetch_res(tile);
        if (role == 0 && tile + num_units < p.split_first) prefetch_res(tile + num_units);
        if (chalf * 32 < (geglu ? ncols_out : bn_t)) load_res(chalf * 32, rsc);
      }
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();

      const int ncols_t = geglu ? ncols_out : bn_t;
      uint32_t v[32];  // accumulator chunk: loaded one chunk ahead (the next load is issued as soon as these registers are consumed)
      if (chalf * 32 < ncols_t) tmem_ld_32x32(t_addr + (uint32_t)(chalf * 32), v);
      for (int c = chalf * 32; c < ncols_t; c += 64) {
        // issue this chunk's global operand loads first so their latency overlaps the TMEM read
        uint4 rv[4], rsn[4];
        const bool has_rv = RV && rowvec_p && row_ok && !geglu;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          rv[g] = make_uint4(0, 0, 0, 0);
          const int n = n0 + c + g * 8;
          if (has_rv && n < p.N)
            rv[g] = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const char*>(rowvec_p) + (rv_off + (size_t)n) * 2));
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) rsn[j] = make_uint4(0, 0, 0, 0);
        if (any_rs && c + 64 < ncols_t) load_res(c + 64, rsn);
        f32x2_t xp[16];
        if (geglu) {
          uint32_t vg[32];
          tmem_ld_32x32(t_addr + (uint32_t)(ncols_out + c), vg);
          tmem_ld_wait();
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            float xv[8], gt[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              xv[i] = __uint_as_float(v[g * 8 + i]);
              gt[i] = __uint_as_float(vg[g * 8 + i]);
            }
            if (ln) {
              ln_apply8(lnc_smem, lnd_smem, c + g * 8, ln_mean, ln_rstd, xv);
              ln_apply8(lnc_smem, lnd_smem, ncols_out + c + g * 8, ln_mean, ln_rstd, gt);
            }
            if (col_bias) {
              add_smem8_f32(bias_smem + (uint32_t)(c + g * 8) * 4u, xv);
              add_smem8_f32(bias_smem + (uint32_t)(ncols_out + c + g * 8) * 4u, gt);
            }
#pragma unroll
            for (int i = 0; i < 8; i += 2)
              xp[g * 4 + (i >> 1)] = mul2(pk2(xv[i], xv[i + 1]), gelu_erf_x2(pk2(gt[i], gt[i + 1])));
          }
          if (c + 64 < ncols_t) tmem_ld_32x32(t_addr + (uint32_t)(c + 64), v);
        } else {
          tmem_ld_wait();
#pragma unroll
          for (int k = 0; k < 16; ++k) xp[k] = pk2(__uint_as_float(v[2 * k]), __uint_as_float(v[2 * k + 1]));
          if (c + 64 < ncols_t) tmem_ld_32x32(t_addr + (uint32_t)(c + 64), v);
          if (EXT && p.alpha != 0.f) {
            const f32x2_t a2 = pk2(p.alpha, p.alpha);
#pragma unroll
            for (int k = 0; k < 16; ++k) xp[k] = mul2(xp[k], a2);
          }
          if (ln) {  // x = rstd * (x - mean * c) + d on pairs
            const f32x2_t nmean2 = pk2(-ln_mean, -ln_mean), rstd2 = pk2(ln_rstd, ln_rstd);
#pragma unroll
            for (int h = 0; h < 8; ++h) {
              float cv[4], dv[4];
              lds_f4(lnc_smem + 4u * (uint32_t)(c + h * 4), cv);
              lds_f4(lnd_smem + 4u * (uint32_t)(c + h * 4), dv);
              xp[2 * h] = fma2(fma2(pk2(cv[0], cv[1]), nmean2, xp[2 * h]), rstd2, pk2(dv[0], dv[1]));
              xp[2 * h + 1] = fma2(fma2(pk2(cv[2], cv[3]), nmean2, xp[2 * h + 1]), rstd2, pk2(dv[2], dv[3]));
            }
          }
          if (col_bias) {
#pragma unroll
            for (int h = 0; h < 8; ++h) {
              float bv[4];
              lds_f4(bias_smem + 4u * (uint32_t)(c + h * 4), bv);
              xp[2 * h] = add2(xp[2 * h], pk2(bv[0], bv[1]));
              xp[2 * h + 1] = add2(xp[2 * h + 1], pk2(bv[2], bv[3]));
            }
          } else if (bias_p) {
            const f32x2_t b2 = pk2(bias_m, bias_m);
#pragma unroll
            for (int k = 0; k < 16; ++k) xp[k] = add2(xp[k], b2);
          }
          if (has_rv) {
            const bool mulv = EXT && p.rowvec_mul;
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              const uint32_t w4[4] = {rv[g].x, rv[g].y, rv[g].z, rv[g].w};
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const float2 f = unpack2<BF16>(w4[i]);
                xp[g * 4 + i] = mulv ? mul2(xp[g * 4 + i], pk2(f.x, f.y)) : add2(xp[g * 4 + i], pk2(f.x, f.y));
              }
            }
          }
          if (act_on && p.epilogue != B200_EPI_NONE) {
#pragma unroll
            for (int k = 0; k < 16; ++k) {
              float xa, xb;
              upk2(xp[k], xa, xb);
              if (p.epilogue == B200_EPI_SILU) {
                xa = silu_f(xa);
                xb = silu_f(xb);
              } else if (p.epilogue == B200_EPI_GELU) {
                upk2(gelu_erf_x2(pk2(xa, xb)), xa, xb);
              } else if (EXT && p.epilogue == B200_EPI_GELU_TANH) {
                xa = gelu_tanh_f(xa);
                xb = gelu_tanh_f(xb);
              }
              xp[k] = pk2(xa, xb);
            }
          }
        }
        __syncwarp();  // the previous chunk's coalesced stores have read the staging buffer
        if (any_rs) {
          // residual tile -> staging (rows as loaded) -> each thread reads its own row's 64 bytes
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int rl = j * 8 + (lane >> 2);
            const uint32_t addr = my_stg + (uint32_t)rl * 64u + ((((uint32_t)piece) ^ (((uint32_t)rl >> 1) & 3u)) << 4);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(rsc[j].x), "r"(rsc[j].y), "r"(rsc[j].z),
                         "r"(rsc[j].w)
                         : "memory");
          }
          __syncwarp();
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const uint32_t addr = my_stg + (uint32_t)lane * 64u + ((((uint32_t)g) ^ sw) << 4);
            uint32_t w4[4];
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(w4[0]), "=r"(w4[1]), "=r"(w4[2]), "=r"(w4[3]) : "r"(addr));
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float2 f = unpack2<BF16>(w4[i]);
              xp[g * 4 + i] = add2(xp[g * 4 + i], pk2(f.x, f.y));
            }
          }
          __syncwarp();  // every lane has its residual row before the buffer is reused for the output tile
        }
        if (row_stats_out && out_n0 + c < p.n_out) {  // row statistics for the next GEMM's folded LayerNorm (n_out % 32 == 0)
          if (st_cnt == 0.f) {
            float dummy;
            upk2(xp[0], st_piv, dummy);
          }
          st_cnt += 32.f;
          const f32x2_t np2 = pk2(-st_piv, -st_piv);
#pragma unroll
          for (int k = 0; k < 16; ++k) {
            const f32x2_t d2 = add2(xp[k], np2);
            st_s2 = add2(st_s2, d2);
            st_q2 = fma2(d2, d2, st_q2);
          }
        }
        // stage this warp's 32 rows x 64 B in its private (64B-swizzled) smem tile, then write them out as
        // full 32-byte sectors: each store instruction covers 8 rows x 64 contiguous bytes
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint32_t o[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            float xa, xb;
            upk2(xp[g * 4 + i], xa, xb);
            o[i] = pack2<BF16>(xa, xb);
          }
          const uint32_t addr = my_stg + (uint32_t)lane * 64u + ((((uint32_t)g) ^ sw) << 4);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(o[0]), "r"(o[1]), "r"(o[2]), "r"(o[3])
                       : "memory");
        }
        __syncwarp();
        {
          const int col = out_n0 + c + piece * 8;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int rl = j * 8 + (lane >> 2);
            const uint32_t addr = my_stg + (uint32_t)rl * 64u + ((((uint32_t)piece) ^ (((uint32_t)rl >> 1) & 3u)) << 4);
            uint32_t o0, o1, o2, o3;
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(o0), "=r"(o1), "=r"(o2), "=r"(o3) : "r"(addr));
            int mm;
            const bool mm_ok = row_of(rl, mm);
            if (mm_ok && col < p.n_out) {
              uint4 o = make_uint4(o0, o1, o2, o3);
              *reinterpret_cast<uint4*>(reinterpret_cast<char*>(p.C) + ((size_t)mm * p.ldc + col) * 2) = o;
            }
          }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) rsc[j] = rsn[j];
      }
      if (row_stats_out && row_ok) {  // one partial per (row, N tile, warp half), written exactly once; [part][row] layout:
        float s0, s1, q0, q1;         // consecutive lanes (rows) write consecutive 16-byte slots
        upk2(st_s2, s0, s1);
        upk2(st_q2, q0, q1);
        const float st_sum = s0 + s1, st_sq = q0 + q1;
        const float inv = st_cnt > 0.f ? 1.0f / st_cnt : 0.f;
        const float ds = st_sum * inv;
        row_stats_out[(size_t)(2 * n_blk + chalf) * p.M + m] =
            make_float4(st_cnt, st_piv + ds, fmaxf(fmaf(-st_sum, ds, st_sq), 0.f), 0.f);
      }
      tc_fence_before();
      if (CG == 2 && cta_rank != 0) mbar_arrive_remote(tempty_bar(acc), 0);
      else mbar_arrive(tempty_bar(acc));
    }
  }

  tc_fence_before();
  if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    if constexpr (CG == 2) tmem_dealloc_cg2(tmem_base, 512); else tmem_dealloc(tmem_base, 512);
  }
}

// ---------------------------------------------------------------------------------------------- host
static int pick_block_n(int N, int epilogue) {
  const int step = (epilogue == B200_EPI_GEGLU) ? 64 : 32;
  // N >= 512: always 256 wide, even with a partly empty last tile: measured on B200 (M = 65536, N = 640) a 256-wide tile with
  // 17 % of its columns wasted still beats four exact 160-wide tiles by 17-26 % (the A tile is re-read once per N tile and
  // the wider MMA amortises it), and N = 1920 gains 5-10 % over 192
  if (N >= 512) return 256;
  if (N >= 256) {
    for (int bn = 256; bn >= 128; bn -= step)
      if (N % bn == 0) return bn;
    return 256;  // tail tile handled by TMA zero fill + guarded stores
  }
  int bn = (N + step - 1) / step * step;
  return bn;
}

// Tile order (GemmKParams::n_fast): re-stream the smaller operand.  B200_GEMM_RASTER = m | n forces one order.
static int raster_n_fast(size_t a_elems, size_t b_elems, int tiles_n) {
  static int forced = -1;
  if (forced < 0) {
    const char* e = getenv("B200_GEMM_RASTER");
    forced = !e ? 0 : (e[0] == 'n' ? 2 : 1);
  }
  if (forced) return forced == 2;
  return (tiles_n > 1 && a_elems > b_elems) ? 1 : 0;
}

static bool use_pair_kernel() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("B200_GEMM_CG");
    v = (e && e[0] == '1') ? 0 : 1;
  }
  return v == 1;
}

// Workspace of the K-split tail (GemmKParams::split_*), caller-owned (b200_gemm_desc::workspace): [num_sms ints of flags,
// padded to 1 KB][num_sms blocks of 256 x 128 floats].  B200_GEMM_SPLITK=0 turns the split off.
static bool split_enabled() {
  static int enabled = -1;
  if (enabled < 0) {
    const char* e = getenv("B200_GEMM_SPLITK");
    enabled = (e && e[0] == '0') ? 0 : 1;
  }
  return enabled == 1;
}
static size_t split_flag_bytes() { return ((size_t)num_sms() * sizeof(int) + 1023) & ~(size_t)1023; }

template <bool BF16, int CG, int FEAT>
static int launch_gemm_t(const CUtensorMap& mapA, const CUtensorMap& mapA2, const CUtensorMap& mapB,
                         const CUtensorMap& mapB2, GemmKParams& p, cudaStream_t stream) {
  const int stage_bytes = kATileBytes + (p.BN / CG) * 128;
  const int staging = (int)(kStageBufs * kStageBufBytes) + 1024 + 2048;
  int S = (222 * 1024 - staging) / stage_bytes;
  if (S > 8) S = 8;
  if (S < 2) S = 2;
  p.num_stages = S;
  const size_t smem = (size_t)S * stage_bytes + staging + 1024 + (2 * S + 4) * 8 + 16;
  const int total = ((p.tiles_m + CG - 1) / CG) * p.tiles_n;
  int units = num_sms() / CG;
  if (units <= 0) {
    set_error("gemm: no device");
    return B200_ENODEVICE;
  }
  if (units > total) units = total;
  // K-split of the last, partly filled wave: rem tail tiles x S shares <= units, every share at least four k-chunks
  p.split_S = 0;
  p.split_first = total;
  p.split_rem = 0;
  {
    const int rem = total % units;
    if (rem > 0 && 2 * rem <= units && p.epilogue != B200_EPI_GEGLU) {
      int S = units / rem;
      if (S > 3) S = 3;
      // worth it only when the k-chunks a tail unit no longer runs outweigh the exchange (~4 us: dump, flag, one L2 round
      // trip per accumulator chunk; a 256 x 256 x 64 chunk is ~0.45 us): measured break-even near 24 chunks
      static int min_saved = -1;
      if (min_saved < 0) {
        const char* e = getenv("B200_GEMM_SPLITK_MIN");
        min_saved = e ? atoi(e) : 24;
      }
      if (S >= 2 && p.num_k_chunks - p.num_k_chunks / S < min_saved) S = 1;
      if (S >= 2 && p.split_ws && split_enabled()) {  // split_ws arrives holding the caller's workspace pointer
        p.split_S = S;
        p.split_first = total - rem;
        p.split_rem = rem;
        p.split_flags = reinterpret_cast<int*>(p.split_ws);
        p.split_ws = reinterpret_cast<float*>(reinterpret_cast<char*>(p.split_ws) + split_flag_bytes());
      }
    }
  }
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(gemm_kernel<BF16, CG, FEAT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) {
      set_error("gemm: smem attr: %s", cudaGetErrorString(e));
      return B200_ECUDA;
    }
    attr_done = true;
  }
  {
    cudaError_t e = launch_pdl(gemm_kernel<BF16, CG, FEAT>, dim3(units * CG), dim3(kThreads), smem, stream, CG, mapA, mapA2, mapB,
                               mapB2, p);
    if (e != cudaSuccess) {
      set_error("gemm: launch failed: %s", cudaGetErrorString(e));
      return B200_ECUDA;
    }
  }
  B200_CHECK_LAUNCH("gemm");
  return B200_OK;
}

// The B tensor map's box height depends on the cluster mode, so the mode is decided before the maps are built.
static int gemm_cg(const GemmKParams& p) { return (use_pair_kernel() && p.tiles_m >= 2 && p.BN % 32 == 0) ? 2 : 1; }

template <int FEAT>
static int launch_gemm_e(const CUtensorMap& mapA, const CUtensorMap& mapA2, const CUtensorMap& mapB,
                         const CUtensorMap& mapB2, GemmKParams& p, int dtype, int cg, cudaStream_t stream) {
  {
    // width of the last N tile, in whole 32-column epilogue chunks (GEGLU tiles are always full: N % BN == 0 is required)
    int last = p.N - (p.tiles_n - 1) * p.BN;
    last = (last + 31) / 32 * 32;
    if (last > p.BN || last <= 0 || p.epilogue == B200_EPI_GEGLU) last = p.BN;
    static int narrow = -1;
    if (narrow < 0) {
      const char* e = getenv("B200_GEMM_NARROW_LAST");
      narrow = (e && e[0] == '0') ? 0 : 1;
    }
    if (!narrow) last = p.BN;
    p.bn_last = last;
    p.idesc_last = make_idesc_f16(cg == 2 ? 256 : 128, last, dtype == B200_BF16, false, false);
  }
  if (cg == 2) {
    p.idesc = make_idesc_f16(256, p.BN, dtype == B200_BF16, false, false);
    return dtype == B200_BF16 ? launch_gemm_t<true, 2, FEAT>(mapA, mapA2, mapB, mapB2, p, stream)
                              : launch_gemm_t<false, 2, FEAT>(mapA, mapA2, mapB, mapB2, p, stream);
  }
  p.idesc = make_idesc_f16(128, p.BN, dtype == B200_BF16, false, false);
  return dtype == B200_BF16 ? launch_gemm_t<true, 1, FEAT>(mapA, mapA2, mapB, mapB2, p, stream)
                            : launch_gemm_t<false, 1, FEAT>(mapA, mapA2, mapB, mapB2, p, stream);
}

static int launch_gemm(const CUtensorMap& mapA, const CUtensorMap& mapA2, const CUtensorMap& mapB,
                       const CUtensorMap& mapB2, GemmKParams& p, int dtype, int cg, cudaStream_t stream) {
  const bool lns = p.ln_stats != nullptr || p.row_stats_out != nullptr;
  const bool ext = p.seg_period != 0 || p.rowvec_mul != 0 || p.act_col0 != 0 || p.epilogue == B200_EPI_GELU_TANH || p.alpha != 0.f ||
                   (lns && p.rowvec != nullptr);  // the LNS-only build has the row vector compiled out
  if (p.tile_w_log2 >= 0 && p.mode == 1 && p.img_n > 0) return launch_gemm_e<4>(mapA, mapA2, mapB, mapB2, p, dtype, cg, stream);
  switch ((lns ? 1 : 0) | (ext ? 2 : 0)) {
    case 0: return launch_gemm_e<0>(mapA, mapA2, mapB, mapB2, p, dtype, cg, stream);
    case 1: return launch_gemm_e<1>(mapA, mapA2, mapB, mapB2, p, dtype, cg, stream);
    case 2: return launch_gemm_e<2>(mapA, mapA2, mapB, mapB2, p, dtype, cg, stream);
    default: return launch_gemm_e<3>(mapA, mapA2, mapB, mapB2, p, dtype, cg, stream);
  }
}

}  // namespace b200

using namespace b200;

extern "C" size_t b200_gemm_workspace_bytes(void) { return split_flag_bytes() + (size_t)num_sms() * 256 * 128 * sizeof(float); }

extern "C" int b200_gemm_row_stats_parts(int N, int epilogue, int block_n) {
  if (N <= 0) return 0;
  const int bn = block_n > 0 ? block_n : pick_block_n(N, epilogue);
  return 2 * ((N + bn - 1) / bn);
}

extern "C" int b200_gemm(const void* A, const void* B, void* C, const b200_gemm_desc* d, b200_stream_t s) {
  B200_CHECK_ARG(A && B && C && d, "gemm: null argument");
  B200_CHECK_ARG(d->M > 0 && d->N > 0 && d->K > 0, "gemm: bad shape %d %d %d", d->M, d->N, d->K);
  B200_CHECK_ARG(d->N % 8 == 0 && d->K % 8 == 0, "gemm: N (%d) and K (%d) must be multiples of 8", d->N, d->K);
  B200_CHECK_ARG(d->lda % 8 == 0 && d->ldb % 8 == 0 && d->ldc % 8 == 0, "gemm: leading dims must be multiples of 8");
  B200_CHECK_ARG(d->dtype == B200_F16 || d->dtype == B200_BF16, "gemm: dtype");
  B200_CHECK_ARG(!d->residual || d->ldr % 8 == 0, "gemm: ldr must be a multiple of 8");
  B200_CHECK_ARG(!d->rowvec || (d->rows_per_vec > 0 && d->ld_rowvec % 8 == 0), "gemm: rowvec layout");
  int K1 = d->K, K2 = 0;
  if (d->A2) {
    K1 = d->K1;
    K2 = d->K - d->K1;
    B200_CHECK_ARG(K1 > 0 && K2 > 0 && K1 % 64 == 0 && d->lda2 % 8 == 0, "gemm: concat split K1=%d K=%d", K1, d->K);
  }
  GemmKParams p;
  memset(&p, 0, sizeof(p));
  p.M = d->M;
  p.N = d->N;
  p.mode = 0;
  const int chunks1 = (K1 + 63) / 64, chunks2 = (K2 + 63) / 64;
  p.num_k_chunks = chunks1 + chunks2;
  p.split_chunk = d->A2 ? chunks1 : p.num_k_chunks;
  int bn = d->block_n > 0 ? d->block_n : pick_block_n(d->N, d->epilogue);
  const int bn_step = d->epilogue == B200_EPI_GEGLU ? 64 : 32;
  B200_CHECK_ARG(bn % bn_step == 0 && bn <= 256, "gemm: block_n %d invalid", bn);
  if (d->epilogue == B200_EPI_GEGLU)
    B200_CHECK_ARG(d->N % bn == 0, "gemm: GEGLU needs N (%d) divisible by block_n (%d)", d->N, bn);
  p.BN = bn;
  p.tiles_m = (d->M + 127) / 128;
  p.tiles_n = (d->N + bn - 1) / bn;
  p.idesc = make_idesc_f16(128, bn, d->dtype == B200_BF16, false, false);
  p.C = C;
  p.ldc = d->ldc;
  p.bias = d->bias;
  p.bias_along_m = d->bias_along_m;
  p.residual = d->residual;
  p.ldr = d->ldr;
  p.n_out = d->epilogue == B200_EPI_GEGLU ? d->N / 2 : d->N;
  p.rowvec = d->rowvec;
  p.ld_rowvec = d->ld_rowvec;
  p.rows_per_vec = d->rows_per_vec > 0 ? d->rows_per_vec : 1;
  p.epilogue = d->epilogue;
  p.ln_stats = reinterpret_cast<const float4*>(d->ln_stats);
  p.ln_parts = d->ln_stats_parts;
  p.ln_c = d->ln_c;
  p.ln_d = d->ln_d;
  p.ln_eps = d->ln_eps;
  p.row_stats_out = reinterpret_cast<float4*>(d->row_stats_out);
  B200_CHECK_ARG((reinterpret_cast<uintptr_t>(d->workspace) & 15) == 0, "gemm: workspace must be 16-byte aligned");
  p.split_ws = reinterpret_cast<float*>(d->workspace);
  B200_CHECK_ARG(!d->ln_stats || (d->ln_c && d->ln_d && !d->A2 && d->ln_stats_parts > 0),
                 "gemm: LayerNorm folding needs ln_c, ln_d and ln_stats_parts (single A source)");
  B200_CHECK_ARG(!d->row_stats_out || (d->epilogue != B200_EPI_GEGLU && d->N % 32 == 0),
                 "gemm: row_stats_out needs N (%d) to be a multiple of 32 and excludes the GEGLU epilogue", d->N);
  p.rowvec_mul = d->rowvec_mul;
  p.act_col0 = d->act_col0;
  p.alpha = (d->alpha == 1.0f) ? 0.f : d->alpha;
  B200_CHECK_ARG(p.alpha == 0.f || d->epilogue != B200_EPI_GEGLU, "gemm: alpha excludes the GEGLU epilogue");
  p.n_fast = raster_n_fast((size_t)d->M * d->K, (size_t)d->N * d->K, p.tiles_n);
  B200_CHECK_ARG(d->act_col0 >= 0 && d->act_col0 % bn == 0, "gemm: act_col0 (%d) must be a multiple of block_n (%d)", d->act_col0, bn);
  if (d->B2) {
    // tiles of 256 rows (CTA pairs) must not straddle a segment boundary
    B200_CHECK_ARG(d->seg_period > 0 && d->seg_split > 0 && d->seg_split < d->seg_period && d->seg_period % 256 == 0 &&
                       d->seg_split % 256 == 0 && d->M % d->seg_period == 0,
                   "gemm: row segments (period %d, split %d) must be multiples of 256 rows", d->seg_period, d->seg_split);
    B200_CHECK_ARG(!d->ln_stats && !d->bias_along_m, "gemm: row segments exclude LayerNorm folding and bias_along_m");
    p.seg_period = d->seg_period;
    p.seg_split = d->seg_split;
    p.bias2 = d->bias2;
    p.rowvec2 = d->rowvec2;
    B200_CHECK_ARG((d->bias != nullptr) == (d->bias2 != nullptr) && (d->rowvec != nullptr) == (d->rowvec2 != nullptr),
                   "gemm: both row segments need the same set of epilogue operands");
  }

  CUtensorMap mA, mA2, mB, mB2;
  {
    uint64_t dims[2] = {(uint64_t)K1, (uint64_t)d->M};
    uint64_t str[1] = {(uint64_t)d->lda * 2};
    uint32_t box[2] = {64, 128};
    int rc = make_tmap(&mA, d->dtype, A, 2, dims, str, box);
    if (rc) return rc;
  }
  if (d->A2) {
    uint64_t dims[2] = {(uint64_t)K2, (uint64_t)d->M};
    uint64_t str[1] = {(uint64_t)d->lda2 * 2};
    uint32_t box[2] = {64, 128};
    int rc = make_tmap(&mA2, d->dtype, d->A2, 2, dims, str, box);
    if (rc) return rc;
  } else {
    mA2 = mA;
  }
  const int cg = gemm_cg(p);
  {
    uint64_t dims[2] = {(uint64_t)d->K, (uint64_t)d->N};
    uint64_t str[1] = {(uint64_t)d->ldb * 2};
    uint32_t box[2] = {64, (uint32_t)(bn / cg)};
    int rc = make_tmap(&mB, d->dtype, B, 2, dims, str, box);
    if (rc) return rc;
    if (d->B2) {
      rc = make_tmap(&mB2, d->dtype, d->B2, 2, dims, str, box);
      if (rc) return rc;
    } else {
      mB2 = mB;
    }
  }
  return launch_gemm(mA, mA2, mB, mB2, p, d->dtype, cg, static_cast<cudaStream_t>(s));
}

static int conv3x3_impl(const void* x1, const void* x2, const void* w_packed, void* y, const b200_conv3x3_desc* d,
                        b200_stream_t s, bool up2x) {
  B200_CHECK_ARG(x1 && w_packed && y && d, "conv3x3: null argument");
  B200_CHECK_ARG(d->N > 0 && d->H > 0 && d->W > 0 && d->Cout > 0, "conv3x3: bad shape");
  B200_CHECK_ARG(d->C1 > 0 && d->C1 % 64 == 0 && d->C2 >= 0 && d->C2 % 64 == 0,
                 "conv3x3: C1 (%d) / C2 (%d) must be multiples of 64 (use b200_im2col3x3 + b200_gemm otherwise)",
                 d->C1, d->C2);
  B200_CHECK_ARG((d->C2 == 0) == (x2 == nullptr), "conv3x3: x2 / C2 mismatch");
  B200_CHECK_ARG(d->Cout % 8 == 0, "conv3x3: Cout must be a multiple of 8");
  B200_CHECK_ARG(d->dtype == B200_F16 || d->dtype == B200_BF16, "conv3x3: dtype");
  // output tile = tile_n images x tile_h rows x tile_w columns = 128 consecutive NHWC pixels
  int tile_w = d->W < 128 ? d->W : 128;
  int tile_h = tile_w > 0 && 128 % tile_w == 0 ? 128 / tile_w : 1;
  if (tile_h > d->H) tile_h = d->H;
  int tile_n = (tile_w * tile_h) > 0 && 128 % (tile_w * tile_h) == 0 ? 128 / (tile_w * tile_h) : 0;
  const bool exact = !up2x && 128 % tile_w == 0 && d->W % tile_w == 0 && d->H % tile_h == 0 && tile_n > 0 &&
                     (tile_n == 1 || (tile_w == d->W && tile_h == d->H));
  int gt_log2 = -1;
  if (!exact) {
    static int general = -1;  // B200_CONV_GENERAL=0: only exactly tiling sizes (the caller falls back to im2col + GEMM)
    if (general < 0) {
      const char* e = getenv("B200_CONV_GENERAL");
      general = (e && e[0] == '0') ? 0 : 1;
    }
    if (!general && !up2x) {
      set_error("conv3x3: %dx%d does not tile into 128-pixel boxes and generic tiling is disabled (B200_CONV_GENERAL=0)", d->H, d->W);
      return B200_EUNSUPPORTED;
    }
    // generic tiling: the widest power-of-two column count that divides W (<= 128), rows to make 128 pixels; tiles
    // overhang the bottom edge (masked stores, zero-filled loads)
    tile_w = 1;
    while (tile_w < 128 && d->W % (tile_w * 2) == 0) tile_w *= 2;
    tile_h = 128 / tile_w;
    tile_n = 1;
    gt_log2 = 0;
    while ((1 << gt_log2) < tile_w) ++gt_log2;
  }
  const int C = d->C1 + d->C2;
  GemmKParams p;
  memset(&p, 0, sizeof(p));
  p.M = d->N * d->H * d->W;
  p.N = d->Cout;
  p.mode = 1;
  const int ntaps = up2x ? 4 : 9;
  p.chunks_per_tap = C / 64;
  p.split_chunk = d->C1 / 64;
  p.num_k_chunks = ntaps * p.chunks_per_tap;
  p.tile_w = tile_w;
  p.tile_h = tile_h;
  p.tile_n = tile_n;
  p.tiles_w = (d->W + tile_w - 1) / tile_w;
  p.tiles_h = (d->H + tile_h - 1) / tile_h;
  p.tile_w_log2 = gt_log2;
  if (gt_log2 >= 0) {
    p.img_n = d->N;
    p.img_h = d->H;
    p.img_w = d->W;
  }
  int bn = d->block_n > 0 ? d->block_n : pick_block_n(d->Cout, d->epilogue);
  B200_CHECK_ARG(bn % 32 == 0 && bn <= 256, "conv3x3: block_n %d invalid", bn);
  B200_CHECK_ARG(d->epilogue != B200_EPI_GEGLU, "conv3x3: GEGLU epilogue not supported");
  p.BN = bn;
  p.tiles_m = ((d->N + tile_n - 1) / tile_n) * p.tiles_w * p.tiles_h;
  if (up2x) {
    // four parities, each over all low-res tiles; an even tile count per parity keeps a CTA pair (two adjacent M tiles, one
    // shared B tile) inside one parity — the padding tile lies past the last image (loads zero-fill, stores masked)
    p.up2x = 1;
    p.tiles_lr = (p.tiles_m + 1) & ~1;
    p.tiles_m = 4 * p.tiles_lr;
    p.M = 4 * d->N * d->H * d->W;
  }
  p.tiles_n = (d->Cout + bn - 1) / bn;
  p.idesc = make_idesc_f16(128, bn, d->dtype == B200_BF16, false, false);
  p.C = y;
  p.ldc = d->Cout;
  p.bias = d->bias;
  p.residual = d->residual;
  p.ldr = d->ldr;
  p.n_out = d->Cout;
  B200_CHECK_ARG(!d->residual || d->ldr % 8 == 0, "conv3x3: ldr");
  p.rowvec = d->temb;
  p.ld_rowvec = d->ld_temb;
  B200_CHECK_ARG(!d->temb || d->ld_temb % 8 == 0, "conv3x3: ld_temb");
  p.split_ws = reinterpret_cast<float*>(d->workspace);
  p.rows_per_vec = (up2x ? 4 : 1) * d->H * d->W;
  p.epilogue = d->epilogue;
  p.n_fast = raster_n_fast((size_t)p.M * C, (size_t)d->Cout * ntaps * C, p.tiles_n);  // activations vs weights (each tap re-reads the same activation rows)

  CUtensorMap mA, mA2, mB;
  auto make4 = [&](CUtensorMap* m, const void* base, int Csrc) {
    uint64_t dims[4] = {(uint64_t)Csrc, (uint64_t)d->W, (uint64_t)d->H, (uint64_t)d->N};
    uint64_t str[3] = {(uint64_t)Csrc * 2, (uint64_t)d->W * Csrc * 2, (uint64_t)d->H * d->W * Csrc * 2};
    uint32_t box[4] = {64, (uint32_t)tile_w, (uint32_t)tile_h, (uint32_t)tile_n};
    return make_tmap(m, d->dtype, base, 4, dims, str, box);
  };
  int rc = make4(&mA, x1, d->C1);
  if (rc) return rc;
  if (x2) {
    rc = make4(&mA2, x2, d->C2);
    if (rc) return rc;
  } else {
    mA2 = mA;
  }
  const int cg = gemm_cg(p);
  {
    uint64_t dims[2] = {(uint64_t)(ntaps * C), (uint64_t)((up2x ? 4 : 1) * d->Cout)};
    uint64_t str[1] = {(uint64_t)(ntaps * C) * 2};
    uint32_t box[2] = {64, (uint32_t)(bn / cg)};
    rc = make_tmap(&mB, d->dtype, w_packed, 2, dims, str, box);
    if (rc) return rc;
  }
  return launch_gemm(mA, mA2, mB, mB, p, d->dtype, cg, static_cast<cudaStream_t>(s));
}

extern "C" int b200_conv3x3(const void* x1, const void* x2, const void* w_packed, void* y,
                            const b200_conv3x3_desc* d, b200_stream_t s) {
  return conv3x3_impl(x1, x2, w_packed, y, d, s, false);
}

extern "C" int b200_conv3x3_up2x(const void* x1, const void* x2, const void* w_packed4, void* y,
                                 const b200_conv3x3_desc* d, b200_stream_t s) {
  return conv3x3_impl(x1, x2, w_packed4, y, d, s, true);
}
