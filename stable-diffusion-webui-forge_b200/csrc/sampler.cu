// b200forge — fused denoise epilogue + classifier-free guidance + k-diffusion sampler update.
// One launch per sampler step over the fp32 latent [B, C, H, W]:
//   eps_u, eps_c  (UNet output, channels-last, fp16/bf16)
//   D_u = x - eps_u*sigma ; D_c = x - eps_c*sigma          (k_prediction.py:81-92, epsilon)
//   D   = D_u + (D_c - D_u) * cfg_scale                     (sampling_function.py:312)
//   euler / euler-a: x += (x - D)/sigma * dt [+ noise * s_noise*sigma_up]   (k_diffusion/sampling.py:131-158,
//                                                  with the to_d override of modules/sd_schedulers.py:10-15)
//   dpm++ 2m: x = c_x*x + c_d*D + c_old*D_old ; D_old = D   (k_diffusion/sampling.py:660-670)
// The arithmetic order follows the reference expression by expression so fp32 results match closely.
#include "common.cuh"
#include "host_util.h"

namespace b200 {

template <bool BF16>
__global__ void __launch_bounds__(256) sampler_step_kernel(float* __restrict__ x, const void* __restrict__ eps,
                                                           const float* __restrict__ noise,
                                                           float* __restrict__ denoised,
                                                           float* __restrict__ old_denoised, b200_step_desc d) {
  const size_t HW = (size_t)d.H * d.W;
  const size_t total = (size_t)d.B * d.C * HW;
  const float sigma = d.sigma;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t hw = i % HW;
    size_t t = i / HW;
    const int c = (int)(t % d.C);
    const int b = (int)(t / d.C);
    const float xv = x[i];
    const size_t e_c = (((size_t)(d.has_uncond ? d.B + b : b)) * HW + hw) * d.ld_eps + c;
    const float ec = ld1<BF16>(eps, e_c);
    float Dc, D;
    if (d.prediction == 1) {  // v-prediction, sigma_data = 1
      const float s2 = sigma * sigma + 1.0f;
      Dc = xv / s2 - ec * sigma / sqrtf(s2);
    } else {
      Dc = xv - ec * sigma;
    }
    if (d.has_uncond) {
      const size_t e_u = ((size_t)b * HW + hw) * d.ld_eps + c;
      const float eu = ld1<BF16>(eps, e_u);
      float Du;
      if (d.prediction == 1) {
        const float s2 = sigma * sigma + 1.0f;
        Du = xv / s2 - eu * sigma / sqrtf(s2);
      } else {
        Du = xv - eu * sigma;
      }
      D = Du + (Dc - Du) * d.cfg_scale;
    } else {
      D = Dc;
    }
    denoised[i] = D;
    float xn;
    if (d.kind == B200_STEP_EULER) {
      const float dd = (xv - D) / sigma;
      xn = xv + dd * d.dt;
      if (d.noise_scale != 0.f) xn = xn + noise[i] * d.noise_scale;
    } else {
      xn = d.c_x * xv + d.c_d * D;
      if (d.c_old != 0.f) xn += d.c_old * old_denoised[i];
      old_denoised[i] = D;
    }
    x[i] = xn;
  }
}

// denoised[b,c,h,w] = x - eps*sigma_b (epsilon) from the channels-last UNet output; used by the
// model_function_wrapper plug point, which must hand Forge a plain NCHW fp32 `denoised` tensor.
template <bool BF16>
__global__ void __launch_bounds__(256) eps_to_denoised_kernel(const float* __restrict__ x, const void* __restrict__ eps,
                                                              const float* __restrict__ sigma, float* __restrict__ out,
                                                              int N, int C, int H, int W, int ld_eps, int prediction) {
  const size_t HW = (size_t)H * W;
  const size_t total = (size_t)N * C * HW;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t hw = i % HW;
    size_t t = i / HW;
    const int c = (int)(t % C);
    const int n = (int)(t / C);
    const float sg = sigma[n];
    const float e = ld1<BF16>(eps, ((size_t)n * HW + hw) * ld_eps + c);
    const float xv = x[i];
    float d;
    if (prediction == 1) {
      const float s2 = sg * sg + 1.0f;
      d = xv / s2 - e * sg / sqrtf(s2);
    } else {
      d = xv - e * sg;
    }
    out[i] = d;
  }
}

// the k-diffusion update alone (the `model` callable already returned the CFG-combined denoised):
// used when the samplers are installed under an unmodified CFGDenoiser (plug point P4).
__global__ void __launch_bounds__(256) sampler_update_kernel(float* __restrict__ x, const float* __restrict__ denoised,
                                                             const float* __restrict__ noise,
                                                             float* __restrict__ old_denoised, b200_step_desc d) {
  const size_t total = (size_t)d.B * d.C * d.H * d.W;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const float xv = x[i];
    const float D = denoised[i];
    float xn;
    if (d.kind == B200_STEP_EULER) {
      const float dd = (xv - D) / d.sigma;
      xn = xv + dd * d.dt;
      if (d.noise_scale != 0.f) xn = xn + noise[i] * d.noise_scale;
    } else if (d.kind == B200_STEP_LINEAR) {
      // generic linear step of the two-evaluation samplers (Heun, DPM2, DPM++ 2S): every operand is read-only
      xn = d.c_x * xv + d.c_d * D;
      if (d.c_old != 0.f) xn += d.c_old * old_denoised[i];
      if (d.noise_scale != 0.f) xn += d.noise_scale * noise[i];
    } else {
      xn = d.c_x * xv + d.c_d * D;
      if (d.c_old != 0.f) xn += d.c_old * old_denoised[i];
      old_denoised[i] = D;
    }
    x[i] = xn;
  }
}

}  // namespace b200

using namespace b200;

extern "C" int b200_eps_to_denoised(const float* x, const void* eps, const float* sigma, float* out, int N, int C, int H,
                                    int W, int ld_eps, int prediction, int eps_dtype, b200_stream_t s) {
  B200_CHECK_ARG(x && eps && sigma && out && N > 0 && C > 0 && H > 0 && W > 0 && ld_eps >= C, "eps_to_denoised: bad arguments");
  const size_t total = (size_t)N * C * H * W;
  size_t g = (total + 255) / 256;
  const size_t cap = (size_t)num_sms() * 8;
  if (g > cap) g = cap;
  if (eps_dtype == B200_BF16)
    eps_to_denoised_kernel<true><<<(int)g, 256, 0, (cudaStream_t)s>>>(x, eps, sigma, out, N, C, H, W, ld_eps, prediction);
  else
    eps_to_denoised_kernel<false><<<(int)g, 256, 0, (cudaStream_t)s>>>(x, eps, sigma, out, N, C, H, W, ld_eps, prediction);
  B200_CHECK_LAUNCH("eps_to_denoised");
  return B200_OK;
}

extern "C" int b200_sampler_update(float* x, const float* denoised, const float* noise, float* old_denoised,
                                   const b200_step_desc* d, b200_stream_t s) {
  B200_CHECK_ARG(x && denoised && d, "sampler_update: null argument");
  B200_CHECK_ARG(d->B > 0 && d->C > 0 && d->H > 0 && d->W > 0, "sampler_update: bad shape");
  B200_CHECK_ARG(d->kind == B200_STEP_EULER || d->kind == B200_STEP_DPMPP_2M || d->kind == B200_STEP_LINEAR, "sampler_update: kind");
  B200_CHECK_ARG(d->kind != B200_STEP_LINEAR || d->c_old == 0.f || old_denoised, "sampler_update: c_old without a third operand");
  B200_CHECK_ARG(d->kind != B200_STEP_DPMPP_2M || old_denoised, "sampler_update: dpm++ 2m needs old_denoised");
  B200_CHECK_ARG(d->noise_scale == 0.f || noise, "sampler_update: noise_scale without noise");
  B200_CHECK_ARG(d->kind != B200_STEP_EULER || d->sigma > 0.f, "sampler_update: sigma must be positive");
  const size_t total = (size_t)d->B * d->C * d->H * d->W;
  size_t g = (total + 255) / 256;
  const size_t cap = (size_t)num_sms() * 8;
  if (g > cap) g = cap;
  sampler_update_kernel<<<(int)g, 256, 0, (cudaStream_t)s>>>(x, denoised, noise, old_denoised, *d);
  B200_CHECK_LAUNCH("sampler_update");
  return B200_OK;
}

extern "C" int b200_sampler_step(float* x, const void* eps, const float* noise, float* denoised, float* old_denoised,
                                 const b200_step_desc* d, b200_stream_t s) {
  B200_CHECK_ARG(x && eps && denoised && d, "sampler_step: null argument");
  B200_CHECK_ARG(d->B > 0 && d->C > 0 && d->H > 0 && d->W > 0 && d->ld_eps >= d->C, "sampler_step: bad shape");
  B200_CHECK_ARG(d->kind == B200_STEP_EULER || d->kind == B200_STEP_DPMPP_2M, "sampler_step: kind");
  B200_CHECK_ARG(d->kind != B200_STEP_DPMPP_2M || old_denoised, "sampler_step: dpm++ 2m needs old_denoised");
  B200_CHECK_ARG(d->noise_scale == 0.f || noise, "sampler_step: noise_scale without noise");
  B200_CHECK_ARG(d->sigma > 0.f, "sampler_step: sigma must be positive");
  const size_t total = (size_t)d->B * d->C * d->H * d->W;
  size_t g = (total + 255) / 256;
  const size_t cap = (size_t)num_sms() * 8;
  if (g > cap) g = cap;
  if (d->eps_dtype == B200_BF16)
    sampler_step_kernel<true><<<(int)g, 256, 0, (cudaStream_t)s>>>(x, eps, noise, denoised, old_denoised, *d);
  else
    sampler_step_kernel<false><<<(int)g, 256, 0, (cudaStream_t)s>>>(x, eps, noise, denoised, old_denoised, *d);
  B200_CHECK_LAUNCH("sampler_step");
  return B200_OK;
}
