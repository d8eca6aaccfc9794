// b200forge — HBM-bound kernels of the Flux (DiT) denoise path (backend/nn/flux.py): modulated LayerNorm,
// query/key RMSNorm + rotary embedding in place on the fused QKV projection, 2x2 patchify / unpatchify.
// Token tensors are row-major [rows, C]; one joint activation holds txt rows then img rows of every sample, and
// "segment" arguments (period = rows per sample, split = txt rows) pick the per-stream parameters by row.
// Every global access is a 16-byte vector.
#include "common.cuh"
#include "host_util.h"

namespace b200 {

template <bool BF16>
__device__ __forceinline__ void ld8(const void* p, size_t elem_off, float (&x)[8]) {
  uint4 r = *reinterpret_cast<const uint4*>(reinterpret_cast<const char*>(p) + elem_off * 2);
  float2 f;
  f = unpack2<BF16>(r.x); x[0] = f.x; x[1] = f.y;
  f = unpack2<BF16>(r.y); x[2] = f.x; x[3] = f.y;
  f = unpack2<BF16>(r.z); x[4] = f.x; x[5] = f.y;
  f = unpack2<BF16>(r.w); x[6] = f.x; x[7] = f.y;
}
template <bool BF16>
__device__ __forceinline__ void st8(void* p, size_t elem_off, const float (&x)[8]) {
  uint4 o;
  o.x = pack2<BF16>(x[0], x[1]);
  o.y = pack2<BF16>(x[2], x[3]);
  o.z = pack2<BF16>(x[4], x[5]);
  o.w = pack2<BF16>(x[6], x[7]);
  *reinterpret_cast<uint4*>(reinterpret_cast<char*>(p) + elem_off * 2) = o;
}
template <bool BF16>
__device__ __forceinline__ float round_act(float v) {  // value after a store in the activation dtype
  if constexpr (BF16) return __bfloat162float(__float2bfloat16_rn(v));
  else return __half2float(__float2half_rn(v));
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ------------------------------------------------------------------------------------------ modulated LayerNorm
// y = (1 + scale[b]) * LayerNorm(x) + shift[b]   (no affine, flux.py:188,190,279,313 with 211-212,232-233,255,259,287)
// one 128-thread block per row; the row lives in registers (NV vectors of 8 per thread), variance by the two-pass
// formula, block reductions through shared memory.
__device__ __forceinline__ float block_sum_128(float v, float* sh) {
  v = warp_sum(v);
  const int w = threadIdx.x >> 5;
  __syncthreads();  // sh may still be read from the previous reduction
  if ((threadIdx.x & 31) == 0) sh[w] = v;
  __syncthreads();
  return (sh[0] + sh[1]) + (sh[2] + sh[3]);
}

template <bool BF16, int NV>
__global__ void __launch_bounds__(128) adaln_kernel(const void* __restrict__ x, void* __restrict__ y, int rows, int C,
                                                    float eps, const void* __restrict__ shift0,
                                                    const void* __restrict__ scale0, const void* __restrict__ shift1,
                                                    const void* __restrict__ scale1, int ld_mod, int seg_period,
                                                    int seg_split) {
  __shared__ float sh[4];
  const int row = blockIdx.x;
  const int nvec = C >> 3;
  float v[NV][8];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int cv = threadIdx.x + 128 * i;
    if (cv < nvec) {
      ld8<BF16>(x, (size_t)row * C + (size_t)cv * 8, v[i]);
#pragma unroll
      for (int j = 0; j < 8; ++j) s += v[i][j];
    }
  }
  const float mean = block_sum_128(s, sh) / (float)C;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    if (threadIdx.x + 128 * i < nvec) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float d = v[i][j] - mean;
        q = fmaf(d, d, q);
      }
    }
  }
  const float rstd = rsqrtf(block_sum_128(q, sh) / (float)C + eps);
  const int b = row / seg_period;
  const bool g1 = (row - b * seg_period) >= seg_split;
  const void* shp = g1 ? shift1 : shift0;
  const void* scp = g1 ? scale1 : scale0;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int cv = threadIdx.x + 128 * i;
    if (cv < nvec) {
      float a[8], c[8], o[8];
      ld8<BF16>(shp, (size_t)b * ld_mod + (size_t)cv * 8, a);
      ld8<BF16>(scp, (size_t)b * ld_mod + (size_t)cv * 8, c);
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = fmaf(1.0f + c[j], (v[i][j] - mean) * rstd, a[j]);
      st8<BF16>(y, (size_t)row * C + (size_t)cv * 8, o);
    }
  }
}

// ------------------------------------------------------------------------------------------ RMSNorm over rows
// y = x * rsqrt(mean(x^2) + eps) * scale   (RMSNorm, flux.py:115-126, as used on [.., hidden] rows by Chroma's Approximator,
// backend/nn/chroma.py:14-28).  Same organisation as adaln_kernel: one 128-thread block per row, row in registers.
template <bool BF16, int NV>
__global__ void __launch_bounds__(128) rmsnorm_rows_kernel(const void* __restrict__ x, const void* __restrict__ scale,
                                                           void* __restrict__ y, int C, float eps) {
  __shared__ float sh[4];
  const int row = blockIdx.x;
  const int nvec = C >> 3;
  float v[NV][8];
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int cv = threadIdx.x + 128 * i;
    if (cv < nvec) {
      ld8<BF16>(x, (size_t)row * C + (size_t)cv * 8, v[i]);
#pragma unroll
      for (int j = 0; j < 8; ++j) q = fmaf(v[i][j], v[i][j], q);
    }
  }
  const float rn = rsqrtf(block_sum_128(q, sh) / (float)C + eps);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int cv = threadIdx.x + 128 * i;
    if (cv < nvec) {
      float w[8], o[8];
      ld8<BF16>(scale, (size_t)cv * 8, w);
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = v[i][j] * rn * w[j];
      st8<BF16>(y, (size_t)row * C + (size_t)cv * 8, o);
    }
  }
}

// ------------------------------------------------------------------------------------------ QK RMSNorm + RoPE
// In place on the q and k thirds of a fused QKV row [3, H, 128] (flux.py:128-139 QKNorm, 15-18 + 45-51 rope):
//   t = rms_norm(x) * scale   (rounded to the activation dtype, as the reference's torch.rms_norm output is)
//   out[2i]   = cos_i * t[2i] - sin_i * t[2i+1];   out[2i+1] = sin_i * t[2i] + cos_i * t[2i+1]   (fp32, then rounded)
// one warp per row; a half-warp (16 lanes x 8 elements) covers one head of 128, so each pass handles two heads.
template <bool BF16>
__global__ void __launch_bounds__(256) qk_norm_rope_kernel(void* __restrict__ qkv, int rows, int H, int ld,
                                                           const void* __restrict__ qs0, const void* __restrict__ ks0,
                                                           const void* __restrict__ qs1, const void* __restrict__ ks1,
                                                           const float* __restrict__ cos_t, const float* __restrict__ sin_t,
                                                           int seg_period, int seg_split, float eps) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int hl = lane & 15;   // position inside the head: elements [8*hl, 8*hl + 8)
  const int hsel = lane >> 4; // which of the two heads of this pass
  const int pos = row % seg_period;
  const bool g1 = pos >= seg_split;
  float qs[8], ks[8], cs[4], sn[4];
  ld8<BF16>(g1 ? qs1 : qs0, (size_t)hl * 8, qs);
  ld8<BF16>(g1 ? ks1 : ks0, (size_t)hl * 8, ks);
  {
    const float4 c4 = *reinterpret_cast<const float4*>(cos_t + (size_t)pos * 64 + hl * 4);
    const float4 s4 = *reinterpret_cast<const float4*>(sin_t + (size_t)pos * 64 + hl * 4);
    cs[0] = c4.x; cs[1] = c4.y; cs[2] = c4.z; cs[3] = c4.w;
    sn[0] = s4.x; sn[1] = s4.y; sn[2] = s4.z; sn[3] = s4.w;
  }
  const int nheads = 2 * H;  // q heads then k heads: k starts H*128 elements after q
  for (int h0 = 0; h0 < nheads; h0 += 2) {
    const int h = h0 + hsel;
    const bool ok = h < nheads;
    float v[8];
    const size_t off = (size_t)row * ld + (size_t)h * 128 + (size_t)hl * 8;
    if (ok) ld8<BF16>(qkv, off, v);
    else {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = 0.f;
    }
    float ss = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) ss = fmaf(v[j], v[j], ss);
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);  // within the half-warp
    const float rn = rsqrtf(ss * (1.0f / 128.0f) + eps);
    const bool is_k = h >= H;
    float t[8], o[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) t[j] = round_act<BF16>(v[j] * rn * (is_k ? ks[j] : qs[j]));
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      o[2 * i] = cs[i] * t[2 * i] - sn[i] * t[2 * i + 1];
      o[2 * i + 1] = sn[i] * t[2 * i] + cs[i] * t[2 * i + 1];
    }
    if (ok) st8<BF16>(qkv, off, o);
  }
}

// ------------------------------------------------------------------------------------------ patchify / unpatchify
// flux.py:398-399: x [B, C, 2h, 2w] (fp32 or activation dtype) -> tokens [B*h*w, ld] with feature (c, ph, pw) at c*4 + ph*2 + pw
template <bool BF16, bool IN_F32>
__global__ void flux_patchify_kernel(const void* __restrict__ x, void* __restrict__ out, int B, int C, int h, int w, int ld) {
  const size_t total = (size_t)B * h * w * C;  // one thread per (token, channel): 4 inputs -> 4 contiguous outputs
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(t % C);
    const size_t tok = t / C;
    const int px = (int)(tok % w);
    const int py = (int)((tok / w) % h);
    const int b = (int)(tok / ((size_t)w * h));
    const size_t base = (((size_t)b * C + c) * (2 * h) + 2 * py) * (2 * w) + 2 * px;
    float v00, v01, v10, v11;
    if constexpr (IN_F32) {
      const float* xf = reinterpret_cast<const float*>(x);
      const float2 r0 = *reinterpret_cast<const float2*>(xf + base);
      const float2 r1 = *reinterpret_cast<const float2*>(xf + base + 2 * w);
      v00 = r0.x; v01 = r0.y; v10 = r1.x; v11 = r1.y;
    } else {
      v00 = ld1<BF16>(x, base); v01 = ld1<BF16>(x, base + 1);
      v10 = ld1<BF16>(x, base + 2 * w); v11 = ld1<BF16>(x, base + 2 * w + 1);
    }
    uint2 o;
    o.x = pack2<BF16>(v00, v01);
    o.y = pack2<BF16>(v10, v11);
    *reinterpret_cast<uint2*>(reinterpret_cast<char*>(out) + (tok * ld + (size_t)c * 4) * 2) = o;
  }
}

// flux.py:412: tokens [B*h*w, ld] -> image; NHWC [B, 2h, 2w, C] in the activation dtype (what the fused sampler step
// reads) or NCHW fp32 [B, C, 2h, 2w] (what KModel.apply_model returns after .float(), k_model.py:44)
template <bool BF16, bool OUT_NCHW_F32>
__global__ void flux_unpatchify_kernel(const void* __restrict__ tok_in, void* __restrict__ out, int B, int C, int h, int w,
                                       int ld) {
  const size_t total = (size_t)B * h * w * C;
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(t % C);
    const size_t tok = t / C;
    const int px = (int)(tok % w);
    const int py = (int)((tok / w) % h);
    const int b = (int)(tok / ((size_t)w * h));
    const uint2 r = *reinterpret_cast<const uint2*>(reinterpret_cast<const char*>(tok_in) + (tok * ld + (size_t)c * 4) * 2);
    const float2 r0 = unpack2<BF16>(r.x), r1 = unpack2<BF16>(r.y);
    if constexpr (OUT_NCHW_F32) {
      float* of = reinterpret_cast<float*>(out);
      const size_t base = (((size_t)b * C + c) * (2 * h) + 2 * py) * (2 * w) + 2 * px;
      *reinterpret_cast<float2*>(of + base) = r0;
      *reinterpret_cast<float2*>(of + base + 2 * w) = r1;
    } else {
      const size_t p00 = (((size_t)b * (2 * h) + 2 * py) * (2 * w) + 2 * px) * C + c;
      st1<BF16>(out, p00, r0.x);
      st1<BF16>(out, p00 + C, r0.y);
      st1<BF16>(out, p00 + (size_t)2 * w * C, r1.x);
      st1<BF16>(out, p00 + (size_t)2 * w * C + C, r1.y);
    }
  }
}

static inline int grid_cap(size_t total, int block) {
  size_t g = (total + block - 1) / block;
  const size_t cap = (size_t)num_sms() * 16;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

}  // namespace b200

using namespace b200;

#define DISPATCH_BF(dtype, ...)       \
  do {                                \
    if ((dtype) == B200_BF16) {       \
      constexpr bool BF = true;       \
      __VA_ARGS__;                    \
    } else {                          \
      constexpr bool BF = false;      \
      __VA_ARGS__;                    \
    }                                 \
  } while (0)

extern "C" int b200_adaln(const void* x, void* y, int rows, int C, float eps, const void* shift0, const void* scale0,
                          const void* shift1, const void* scale1, int ld_mod, int seg_period, int seg_split, int dtype,
                          b200_stream_t s) {
  B200_CHECK_ARG(x && y && shift0 && scale0 && rows > 0 && C > 0, "adaln: bad arguments");
  B200_CHECK_ARG(C % 8 == 0 && C <= 4096 && ld_mod % 8 == 0, "adaln: C (%d) must be a multiple of 8 and <= 4096", C);
  B200_CHECK_ARG(dtype == B200_F16 || dtype == B200_BF16, "adaln: dtype");
  if (seg_period <= 0) { seg_period = rows; seg_split = rows; }
  if (!shift1) { shift1 = shift0; scale1 = scale0; }
  const int nv = (C / 8 + 127) / 128;
#define ADALN_LAUNCH(NV) \
  DISPATCH_BF(dtype, (adaln_kernel<BF, NV><<<rows, 128, 0, (cudaStream_t)s>>>(x, y, rows, C, eps, shift0, scale0, shift1, scale1, ld_mod, seg_period, seg_split)))
  if (nv <= 1) ADALN_LAUNCH(1);
  else if (nv <= 2) ADALN_LAUNCH(2);
  else if (nv <= 3) ADALN_LAUNCH(3);
  else ADALN_LAUNCH(4);
#undef ADALN_LAUNCH
  B200_CHECK_LAUNCH("adaln");
  return B200_OK;
}

extern "C" int b200_qk_norm_rope(void* qkv, int rows, int H, int Dh, int ld, const void* q_scale0, const void* k_scale0,
                                 const void* q_scale1, const void* k_scale1, const float* cos_t, const float* sin_t,
                                 int seg_period, int seg_split, float eps, int dtype, b200_stream_t s) {
  B200_CHECK_ARG(qkv && q_scale0 && k_scale0 && cos_t && sin_t && rows > 0 && H > 0, "qk_norm_rope: bad arguments");
  if (Dh != 128) {
    set_error("qk_norm_rope: head dim %d not supported (128 only)", Dh);
    return B200_EUNSUPPORTED;
  }
  B200_CHECK_ARG(ld % 8 == 0 && ld >= 2 * H * Dh && seg_period > 0, "qk_norm_rope: layout");
  B200_CHECK_ARG(dtype == B200_F16 || dtype == B200_BF16, "qk_norm_rope: dtype");
  if (!q_scale1) { q_scale1 = q_scale0; k_scale1 = k_scale0; }
  DISPATCH_BF(dtype, (qk_norm_rope_kernel<BF><<<(rows + 7) / 8, 256, 0, (cudaStream_t)s>>>(
                         qkv, rows, H, ld, q_scale0, k_scale0, q_scale1, k_scale1, cos_t, sin_t, seg_period, seg_split, eps)));
  B200_CHECK_LAUNCH("qk_norm_rope");
  return B200_OK;
}

extern "C" int b200_flux_patchify(const void* x, void* tokens, int B, int C, int H, int W, int ld, int in_is_f32, int dtype,
                                  b200_stream_t s) {
  B200_CHECK_ARG(x && tokens && B > 0 && C > 0 && H > 0 && W > 0, "flux_patchify: bad arguments");
  if ((H | W) & 1) {
    set_error("flux_patchify: odd latent size %dx%d (circular padding path) not supported", H, W);
    return B200_EUNSUPPORTED;
  }
  B200_CHECK_ARG(ld >= 4 * C && ld % 4 == 0, "flux_patchify: ld");
  B200_CHECK_ARG(dtype == B200_F16 || dtype == B200_BF16, "flux_patchify: dtype");
  const size_t total = (size_t)B * (H / 2) * (W / 2) * C;
  const int grid = grid_cap(total, 256);
  if (in_is_f32) DISPATCH_BF(dtype, (flux_patchify_kernel<BF, true><<<grid, 256, 0, (cudaStream_t)s>>>(x, tokens, B, C, H / 2, W / 2, ld)));
  else DISPATCH_BF(dtype, (flux_patchify_kernel<BF, false><<<grid, 256, 0, (cudaStream_t)s>>>(x, tokens, B, C, H / 2, W / 2, ld)));
  B200_CHECK_LAUNCH("flux_patchify");
  return B200_OK;
}

extern "C" int b200_flux_unpatchify(const void* tokens, void* out, int B, int C, int H, int W, int ld, int out_nchw_f32,
                                    int dtype, b200_stream_t s) {
  B200_CHECK_ARG(tokens && out && B > 0 && C > 0 && H > 0 && W > 0, "flux_unpatchify: bad arguments");
  B200_CHECK_ARG(((H | W) & 1) == 0 && ld >= 4 * C && ld % 4 == 0, "flux_unpatchify: layout");
  B200_CHECK_ARG(dtype == B200_F16 || dtype == B200_BF16, "flux_unpatchify: dtype");
  const size_t total = (size_t)B * (H / 2) * (W / 2) * C;
  const int grid = grid_cap(total, 256);
  if (out_nchw_f32) DISPATCH_BF(dtype, (flux_unpatchify_kernel<BF, true><<<grid, 256, 0, (cudaStream_t)s>>>(tokens, out, B, C, H / 2, W / 2, ld)));
  else DISPATCH_BF(dtype, (flux_unpatchify_kernel<BF, false><<<grid, 256, 0, (cudaStream_t)s>>>(tokens, out, B, C, H / 2, W / 2, ld)));
  B200_CHECK_LAUNCH("flux_unpatchify");
  return B200_OK;
}

extern "C" int b200_rmsnorm_rows(const void* x, const void* scale, void* y, int rows, int C, float eps, int dtype,
                                 b200_stream_t s) {
  B200_CHECK_ARG(x && scale && y && rows > 0 && C > 0, "rmsnorm_rows: bad arguments");
  B200_CHECK_ARG(C % 8 == 0 && C <= 8192, "rmsnorm_rows: C (%d) must be a multiple of 8 and <= 8192", C);
  B200_CHECK_ARG(dtype == B200_F16 || dtype == B200_BF16, "rmsnorm_rows: dtype");
  const int nv = (C / 8 + 127) / 128;
#define RMS_LAUNCH(NV) DISPATCH_BF(dtype, (rmsnorm_rows_kernel<BF, NV><<<rows, 128, 0, (cudaStream_t)s>>>(x, scale, y, C, eps)))
  if (nv <= 1) RMS_LAUNCH(1);
  else if (nv <= 2) RMS_LAUNCH(2);
  else if (nv <= 4) RMS_LAUNCH(4);
  else if (nv <= 6) RMS_LAUNCH(6);
  else RMS_LAUNCH(8);
#undef RMS_LAUNCH
  B200_CHECK_LAUNCH("rmsnorm_rows");
  return B200_OK;
}
