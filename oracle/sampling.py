"""ORACLE — test infrastructure only (see oracle/ops.py header).

fp32 restatement of the reference's denoise-step arithmetic around the UNet:
noise parameterisation (backend/modules/k_prediction.py), KModel.apply_model (backend/modules/k_model.py),
the cond/uncond batching + CFG combine (backend/sampling/sampling_function.py), sigma schedules
(k_diffusion/external.py, k_diffusion/sampling.py) and the Euler / Euler-ancestral / DPM++ 2M loops
(k_diffusion/sampling.py, with the to_d override of modules/sd_schedulers.py:10-15).
"""
from __future__ import annotations

import math
from typing import Callable, Optional

import torch


# ------------------------------------------------------------------------------------------- schedule
def make_sigmas_scaled_linear(linear_start: float = 0.00085, linear_end: float = 0.012, timesteps: int = 1000):
    """backend/modules/k_prediction.py:18-22,127-133 (beta_schedule 'linear' = scaled-linear betas in fp64,
    sigma = sqrt((1 - acp)/acp) -> fp32)."""
    betas = torch.linspace(linear_start ** 0.5, linear_end ** 0.5, timesteps, dtype=torch.float64) ** 2
    alphas_cumprod = torch.cumprod(1.0 - betas, dim=0)
    sigmas = ((1 - alphas_cumprod) / alphas_cumprod) ** 0.5
    return sigmas  # fp64; the reference registers sigmas.float() and sigmas.log().float() (log taken in fp64)


class EpsPrediction:
    """backend/modules/k_prediction.py:113-167 (Prediction, prediction_type='epsilon', sigma_data=1)."""

    def __init__(self):
        s64 = make_sigmas_scaled_linear()
        self.sigmas = s64.float()
        self.log_sigmas = s64.log().float()
        self.sigma_data = 1.0

    @property
    def sigma_min(self):
        return self.sigmas[0]

    @property
    def sigma_max(self):
        return self.sigmas[-1]

    def timestep(self, sigma: torch.Tensor) -> torch.Tensor:
        """:148-151 — index of the nearest log-sigma."""
        log_sigma = sigma.log()
        dists = log_sigma.to(self.log_sigmas.device) - self.log_sigmas[:, None]
        return dists.abs().argmin(dim=0).view(sigma.shape).to(sigma.device)

    def sigma(self, timestep: torch.Tensor) -> torch.Tensor:
        """:153-159 — log-linear interpolation."""
        t = torch.clamp(timestep.float(), min=0, max=(len(self.sigmas) - 1))
        low_idx = t.floor().long()
        high_idx = t.ceil().long()
        w = t.frac()
        log_sigma = (1 - w) * self.log_sigmas[low_idx] + w * self.log_sigmas[high_idx]
        return log_sigma.exp()

    def calculate_input(self, sigma: torch.Tensor, noise: torch.Tensor) -> torch.Tensor:
        """:74-79"""
        sigma = sigma.view(sigma.shape[:1] + (1,) * (noise.ndim - 1))
        return noise / (sigma ** 2 + self.sigma_data ** 2) ** 0.5

    def calculate_denoised(self, sigma, model_output, model_input):
        """:81-92 (epsilon branch)"""
        sigma = sigma.view(sigma.shape[:1] + (1,) * (model_output.ndim - 1))
        return model_input - model_output * sigma

    def noise_scaling(self, sigma, noise, latent_image, max_denoise=False):
        """:94-104"""
        if max_denoise:
            noise = noise * torch.sqrt(1.0 + sigma ** 2.0)
        else:
            noise = noise * sigma
        return noise + latent_image


def append_zero(x):
    return torch.cat([x, x.new_zeros([1])])


def get_sigmas_uniform(pred: EpsPrediction, n: int) -> torch.Tensor:
    """k_diffusion/external.py:62-67 ForgeScheduleLinker.get_sigmas(n) — the 'Automatic' schedule of
    Euler / Euler a (modules/sd_schedulers.py uniform)."""
    t_max = len(pred.sigmas) - 1
    t = torch.linspace(t_max, 0, n)
    return append_zero(pred.sigma(t))


def get_sigmas_karras(n: int, sigma_min: float, sigma_max: float, rho: float = 7.0) -> torch.Tensor:
    """k_diffusion/sampling.py:19-25"""
    ramp = torch.linspace(0, 1, n)
    min_inv_rho = sigma_min ** (1 / rho)
    max_inv_rho = sigma_max ** (1 / rho)
    sigmas = (max_inv_rho + ramp * (min_inv_rho - max_inv_rho)) ** rho
    return append_zero(sigmas)


# ------------------------------------------------------------------------------------------- per-step math
def cfg_denoised_eps(x, eps_uncond, eps_cond, sigma: float, cond_scale: float):
    """k_model.py:45-46 + k_prediction.py:92 (denoised = x - eps*sigma, per branch) then
    sampling_function.py:312 (uncond + (cond - uncond) * scale)."""
    d_u = x - eps_uncond * sigma
    d_c = x - eps_cond * sigma
    return d_u + (d_c - d_u) * cond_scale


def to_d(x, sigma, denoised):
    """modules/sd_schedulers.py:10-15 override of k_diffusion.sampling.to_d."""
    return (x - denoised) / sigma


def get_ancestral_step(sigma_from: float, sigma_to: float, eta: float = 1.0):
    """k_diffusion/sampling.py:53-60"""
    if not eta:
        return sigma_to, 0.0
    sigma_up = min(sigma_to, eta * (sigma_to ** 2 * (sigma_from ** 2 - sigma_to ** 2) / sigma_from ** 2) ** 0.5)
    sigma_down = (sigma_to ** 2 - sigma_up ** 2) ** 0.5
    return sigma_down, sigma_up


def euler_step(x, denoised, sigma: float, sigma_next: float):
    """k_diffusion/sampling.py:131-136 with s_churn = 0 (sigma_hat = sigma)."""
    d = to_d(x, sigma, denoised)
    return x + d * (sigma_next - sigma)


def euler_ancestral_step(x, denoised, sigma: float, sigma_next: float, noise, eta: float = 1.0, s_noise: float = 1.0):
    """k_diffusion/sampling.py:149-158"""
    sigma_down, sigma_up = get_ancestral_step(sigma, sigma_next, eta)
    d = to_d(x, sigma, denoised)
    x = x + d * (sigma_down - sigma)
    if sigma_next > 0:
        x = x + noise * s_noise * sigma_up
    return x


def dpmpp_2m_coeffs(sigma_prev: Optional[float], sigma: float, sigma_next: float, has_old: bool):
    """Scalar coefficients of k_diffusion/sampling.py:660-669 written as x' = c_x x + c_d D + c_old D_old."""
    t, t_next = -math.log(sigma), (-math.log(sigma_next) if sigma_next > 0 else math.inf)
    h = t_next - t
    ratio = sigma_next / sigma  # sigma_fn(t_next)/sigma_fn(t)
    em = -math.expm1(-h) if math.isfinite(h) else 1.0  # -(-h).expm1()
    if not has_old or sigma_next == 0:
        return ratio, em, 0.0
    h_last = t - (-math.log(sigma_prev))
    r = h_last / h
    return ratio, em * (1 + 1 / (2 * r)), -em * (1 / (2 * r))


def dpmpp_2m_step(x, denoised, old_denoised, sigma_prev, sigma: float, sigma_next: float):
    """k_diffusion/sampling.py:658-669 (tensor form, as the reference writes it)."""
    sig, sig_n = torch.tensor(sigma), torch.tensor(sigma_next)
    t, t_next = sig.log().neg(), sig_n.log().neg()
    h = t_next - t
    if old_denoised is None or sigma_next == 0:
        return (t_next.neg().exp() / t.neg().exp()) * x - (-h).expm1() * denoised
    h_last = t - torch.tensor(sigma_prev).log().neg()
    r = h_last / h
    denoised_d = (1 + 1 / (2 * r)) * denoised - (1 / (2 * r)) * old_denoised
    return (t_next.neg().exp() / t.neg().exp()) * x - (-h).expm1() * denoised_d


# ------------------------------------------------------------------------------------------- denoiser + loops
class VPrediction(EpsPrediction):
    """Prediction(prediction_type='v_prediction') — SD2.x 768-v (backend/modules/k_prediction.py:81-92, v branch, sigma_data 1)."""

    def calculate_denoised(self, sigma, model_output, model_input):
        sigma = sigma.view(sigma.shape[:1] + (1,) * (model_output.ndim - 1))
        sd = self.sigma_data
        return model_input * sd ** 2 / (sigma ** 2 + sd ** 2) - model_output * sigma * sd / (sigma ** 2 + sd ** 2) ** 0.5


class Denoiser:
    """KModel.apply_model (backend/modules/k_model.py:25-46) + calc_cond_uncond_batch / CFG
    (backend/sampling/sampling_function.py:154-322) for the plain txt2img case: one cond and one uncond
    entry, both batched into a single UNet call ordered [uncond, cond] (:186-188 reverses the run list).

    unet(x[2B,4,h,w], t[2B], context[2B,77,ctx], y[2B,adm]|None) -> eps[2B,4,h,w]; `compute_dtype` is the
    dtype the reference casts the UNet inputs to (k_model.py:34-36)."""

    def __init__(self, unet: Callable, pred: EpsPrediction, cond: dict, uncond: dict, cond_scale: float,
                 compute_dtype: torch.dtype = torch.float32):
        self.unet, self.pred, self.cond, self.uncond = unet, pred, cond, uncond
        self.cond_scale = cond_scale
        self.compute_dtype = compute_dtype

    def apply_model(self, x, sigma, context, y):
        xc = self.pred.calculate_input(sigma, x).to(self.compute_dtype)
        t = self.pred.timestep(sigma).float()
        out = self.unet(xc, t, context.to(self.compute_dtype), None if y is None else y.to(self.compute_dtype)).float()
        return self.pred.calculate_denoised(sigma, out, x)

    def __call__(self, x, sigma):
        """x [B,4,h,w] fp32, sigma [B] -> CFG-combined denoised (sampling_function_inner, :292-322)."""
        xin = torch.cat([x, x])
        sig = torch.cat([sigma, sigma])
        ctx = torch.cat([self.uncond["crossattn"], self.cond["crossattn"]])
        y = None
        if self.cond.get("vector") is not None:
            y = torch.cat([self.uncond["vector"], self.cond["vector"]])
        out = self.apply_model(xin, sig, ctx, y)
        uncond_pred, cond_pred = out.chunk(2)
        return uncond_pred + (cond_pred - uncond_pred) * self.cond_scale


def sample_euler(model, x, sigmas, callback=None):
    """k_diffusion/sampling.py:119-137 with s_churn=0 (gamma = 0)."""
    s_in = x.new_ones([x.shape[0]])
    for i in range(len(sigmas) - 1):
        denoised = model(x, sigmas[i] * s_in)
        if callback is not None:
            callback(i, x, denoised)
        x = euler_step(x, denoised, float(sigmas[i]), float(sigmas[i + 1]))
    return x


def sample_euler_ancestral(model, x, sigmas, noise_sampler, eta=1.0, s_noise=1.0, callback=None):
    """k_diffusion/sampling.py:140-159; noise_sampler() -> N(0,1) like x, drawn only when sigma_next > 0."""
    s_in = x.new_ones([x.shape[0]])
    for i in range(len(sigmas) - 1):
        denoised = model(x, sigmas[i] * s_in)
        if callback is not None:
            callback(i, x, denoised)
        sn = float(sigmas[i + 1])
        noise = noise_sampler() if sn > 0 else None
        x = euler_ancestral_step(x, denoised, float(sigmas[i]), sn, noise, eta, s_noise)
    return x


def sample_dpmpp_2m(model, x, sigmas, callback=None):
    """k_diffusion/sampling.py:648-671"""
    s_in = x.new_ones([x.shape[0]])
    old = None
    for i in range(len(sigmas) - 1):
        denoised = model(x, sigmas[i] * s_in)
        if callback is not None:
            callback(i, x, denoised)
        x = dpmpp_2m_step(x, denoised, old, float(sigmas[i - 1]) if i > 0 else None, float(sigmas[i]),
                          float(sigmas[i + 1]))
        old = denoised
    return x


def image_rng_noise(shape, seeds, device="cpu"):
    """modules/rng.py:113-177 (ImageRNG with randn_source='GPU'/'CPU'): one torch.Generator per seed,
    torch.randn(shape) per image, stacked.  Returns (first_noise, next_fn) where next_fn() draws the
    per-step ancestral noise from the same generators (ImageRNG.next)."""
    gens = [torch.Generator(device=device).manual_seed(int(s)) for s in seeds]

    def draw():
        return torch.stack([torch.randn(shape, generator=g, device=device) for g in gens])

    return draw(), draw


# --------------------------------------------------------------------------------------------- Flux (flow matching)
# PredictionFlux (backend/modules/k_prediction.py:285-322) builds its sigma table with two helpers of the third-party
# package `diffusers` (not vendored under /root/reference, not installed here; Forge pins diffusers==0.31.0 in
# requirements_versions.txt): pipelines.flux.pipeline_flux.calculate_shift and
# FlowMatchEulerDiscreteScheduler.time_shift.  Both are restated from the published algorithm — PARITY UNPINNED for the
# table itself (no reference output to compare against in this container); everything downstream of the table
# (simple_scheduler, the Euler update, `const` prediction) follows reference code that is present.
def flux_calculate_shift(image_seq_len: int, base_seq_len: int = 256, max_seq_len: int = 4096, base_shift: float = 0.5,
                         max_shift: float = 1.15) -> float:
    m = (max_shift - base_shift) / (max_seq_len - base_seq_len)
    b = base_shift - m * base_seq_len
    return image_seq_len * m + b


def flux_sigma_table(seq_len: int = 4096, pseudo_timestep_range: int = 10000, mu=None) -> torch.Tensor:
    """k_prediction.py:292-301: sigmas = time_shift(mu, 1.0, arange(1, N+1)/N), time_shift = e^mu / (e^mu + (1/t - 1))."""
    if mu is None:
        mu = flux_calculate_shift(seq_len)
    t = torch.arange(1, pseudo_timestep_range + 1, 1) / pseudo_timestep_range
    return math.exp(mu) / (math.exp(mu) + (1 / t - 1) ** 1.0)


def simple_scheduler(n: int, table: torch.Tensor) -> torch.Tensor:
    """modules/sd_schedulers.py:81-87 (the default scheduler Forge selects for Flux)."""
    ss = len(table) / n
    sigs = [float(table[-(1 + int(x * ss))]) for x in range(n)]
    return torch.FloatTensor(sigs + [0.0])


def const_noise_scaling(sigma, noise, latent):
    """k_prediction.py:94-96 for prediction_type 'const'."""
    return sigma * noise + (1.0 - sigma) * latent


def const_denoised(x, model_output, sigma):
    """k_prediction.py:81-92 'const' branch: model_input - model_output * sigma (calculate_input is the identity, :74-76)."""
    return x - model_output * sigma


# --------------------------------------------------------------------------------------------- two-evaluation samplers
def _anc_t(sig_from: torch.Tensor, sig_to: torch.Tensor, eta: float):
    """get_ancestral_step (k_diffusion/sampling.py:53-60) on 0-dim tensors."""
    if not eta:
        return sig_to, sig_to.new_zeros(())
    up = torch.minimum(sig_to, eta * (sig_to ** 2 * (sig_from ** 2 - sig_to ** 2) / sig_from ** 2) ** 0.5)
    return (sig_to ** 2 - up ** 2) ** 0.5, up


def sample_heun(model, x, sigmas):
    """k_diffusion/sampling.py:188-214 with s_churn = 0."""
    s_in = x.new_ones([x.shape[0]])
    for i in range(len(sigmas) - 1):
        denoised = model(x, sigmas[i] * s_in)
        d = to_d(x, sigmas[i], denoised)
        dt = sigmas[i + 1] - sigmas[i]
        if sigmas[i + 1] == 0:
            x = x + d * dt
        else:
            x_2 = x + d * dt
            d_2 = to_d(x_2, sigmas[i + 1], model(x_2, sigmas[i + 1] * s_in))
            x = x + (d + d_2) / 2 * dt
    return x


def sample_dpm_2(model, x, sigmas):
    """k_diffusion/sampling.py:217-246 with s_churn = 0."""
    s_in = x.new_ones([x.shape[0]])
    for i in range(len(sigmas) - 1):
        denoised = model(x, sigmas[i] * s_in)
        d = to_d(x, sigmas[i], denoised)
        if sigmas[i + 1] == 0:
            x = x + d * (sigmas[i + 1] - sigmas[i])
        else:
            sigma_mid = sigmas[i].log().lerp(sigmas[i + 1].log(), 0.5).exp()
            x_2 = x + d * (sigma_mid - sigmas[i])
            d_2 = to_d(x_2, sigma_mid, model(x_2, sigma_mid * s_in))
            x = x + d_2 * (sigmas[i + 1] - sigmas[i])
    return x


def sample_dpm_2_ancestral(model, x, sigmas, noise_sampler, eta=1.0, s_noise=1.0):
    """k_diffusion/sampling.py:249-276; noise_sampler() -> N(0,1) like x, drawn only in the second-order branch."""
    s_in = x.new_ones([x.shape[0]])
    for i in range(len(sigmas) - 1):
        denoised = model(x, sigmas[i] * s_in)
        sigma_down, sigma_up = _anc_t(sigmas[i], sigmas[i + 1], eta)
        d = to_d(x, sigmas[i], denoised)
        if sigma_down == 0:
            x = x + d * (sigma_down - sigmas[i])
        else:
            sigma_mid = sigmas[i].log().lerp(sigma_down.log(), 0.5).exp()
            x_2 = x + d * (sigma_mid - sigmas[i])
            d_2 = to_d(x_2, sigma_mid, model(x_2, sigma_mid * s_in))
            x = x + d_2 * (sigma_down - sigmas[i])
            x = x + noise_sampler() * s_noise * sigma_up
    return x


def sample_dpmpp_2s_ancestral(model, x, sigmas, noise_sampler, eta=1.0, s_noise=1.0):
    """k_diffusion/sampling.py:573-603."""
    s_in = x.new_ones([x.shape[0]])
    for i in range(len(sigmas) - 1):
        denoised = model(x, sigmas[i] * s_in)
        sigma_down, sigma_up = _anc_t(sigmas[i], sigmas[i + 1], eta)
        if sigma_down == 0:
            x = x + to_d(x, sigmas[i], denoised) * (sigma_down - sigmas[i])
        else:
            t, t_next = sigmas[i].log().neg(), sigma_down.log().neg()
            h = t_next - t
            s = t + 0.5 * h
            x_2 = (s.neg().exp() / t.neg().exp()) * x - (-h * 0.5).expm1() * denoised
            denoised_2 = model(x_2, s.neg().exp() * s_in)
            x = (t_next.neg().exp() / t.neg().exp()) * x - (-h).expm1() * denoised_2
        if sigmas[i + 1] > 0:
            x = x + noise_sampler() * s_noise * sigma_up
    return x


def toy_denoiser(x, sigma):
    """A cheap deterministic stand-in for the CFG denoiser used to pin sampler arithmetic (fixtures: samplers_toy.pt):
    smooth and non-linear in x, sigma-dependent, shape-preserving."""
    s = sigma.view(-1, 1, 1, 1)
    return x / (1.0 + s * s) + 0.3 * torch.tanh(x * 0.5) * s / (1.0 + s)
