"""Host side of the denoise loop: noise schedule, per-step scalar tables and the fused step launches.

Mirrors (same names, argument meaning and results) the reference pieces that surround the UNet:
  * `Prediction`                 backend/modules/k_prediction.py:113-167  (epsilon parameterisation)
  * `get_sigmas_uniform/karras`  k_diffusion/external.py:62-67, k_diffusion/sampling.py:19-25
  * `get_ancestral_step`         k_diffusion/sampling.py:53-60
  * `sample_euler`, `sample_euler_ancestral`, `sample_dpmpp_2m`   k_diffusion/sampling.py:119-159, 648-671
The reference evaluates the sigma-dependent scalars as 0-dim GPU tensors inside the loop (two to three
device->host syncs per step); here the whole schedule is reduced to Python floats once and
each step is: one graph replay of the UNet forward + ONE `b200_sampler_step` launch (CFG + update fused).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Callable, List, Optional, Sequence

import torch

from . import ops


class Prediction:
    """Epsilon-prediction schedule (scaled-linear betas), backend/modules/k_prediction.py:113-167."""

    def __init__(self, linear_start: float = 0.00085, linear_end: float = 0.012, timesteps: int = 1000,
                 prediction_type: str = "epsilon"):
        betas = torch.linspace(linear_start ** 0.5, linear_end ** 0.5, timesteps, dtype=torch.float64) ** 2
        acp = torch.cumprod(1.0 - betas, dim=0)
        sig = ((1 - acp) / acp) ** 0.5
        self.sigmas = sig.float()
        self.log_sigmas = sig.log().float()
        self.prediction_type = prediction_type
        self.sigma_data = 1.0
        self._log_sigmas_dev: dict = {}

    @property
    def sigma_min(self) -> float:
        return float(self.sigmas[0])

    @property
    def sigma_max(self) -> float:
        return float(self.sigmas[-1])

    def timestep(self, sigma: torch.Tensor) -> torch.Tensor:
        """index of the nearest log-sigma (k_prediction.py:148-151), evaluated on sigma's device like the reference (whose
        table is a module buffer): the host for the schedule, the GPU when Forge's sampler calls the P3 wrapper per step."""
        ls = self.log_sigmas
        if sigma.device != ls.device:
            ls = self._log_sigmas_dev.get(sigma.device)
            if ls is None:
                ls = self._log_sigmas_dev[sigma.device] = self.log_sigmas.to(sigma.device)
        dists = sigma.float().log() - ls[:, None]
        return dists.abs().argmin(dim=0).view(sigma.shape)

    def sigma(self, timestep: torch.Tensor) -> torch.Tensor:
        t = torch.clamp(timestep.float(), min=0, max=len(self.sigmas) - 1)
        lo, hi, w = t.floor().long(), t.ceil().long(), t.frac()
        return ((1 - w) * self.log_sigmas[lo] + w * self.log_sigmas[hi]).exp()

    def noise_scaling(self, sigma, noise, latent_image=None, max_denoise: bool = False):
        """k_prediction.py:94-104"""
        noise = noise * (torch.sqrt(1.0 + sigma ** 2.0) if max_denoise else sigma)
        return noise if latent_image is None else noise + latent_image


class FluxPrediction:
    """PredictionFlux (backend/modules/k_prediction.py:285-322): prediction type 'const' (model input = x, denoised =
    x - sigma * output, timestep = sigma) over a 10000-entry sigma table shifted by mu.  mu and the shift come from the
    third-party `diffusers` helpers the reference imports (calculate_shift, FlowMatchEulerDiscreteScheduler.time_shift),
    restated here from their published form."""

    def __init__(self, seq_len: int = 4096, base_seq_len: int = 256, max_seq_len: int = 4096, base_shift: float = 0.5,
                 max_shift: float = 1.15, pseudo_timestep_range: int = 10000, mu: Optional[float] = None):
        if mu is None:
            m = (max_shift - base_shift) / (max_seq_len - base_seq_len)
            mu = seq_len * m + (base_shift - m * base_seq_len)
        self.mu = mu
        t = torch.arange(1, pseudo_timestep_range + 1, 1) / pseudo_timestep_range
        self.sigmas = math.exp(mu) / (math.exp(mu) + (1 / t - 1) ** 1.0)

    @property
    def sigma_min(self) -> float:
        return float(self.sigmas[0])

    @property
    def sigma_max(self) -> float:
        return float(self.sigmas[-1])

    def timestep(self, sigma: torch.Tensor) -> torch.Tensor:
        return sigma

    def noise_scaling(self, sigma, noise, latent_image=None, max_denoise: bool = False):
        out = sigma * noise
        if latent_image is not None:
            out = out + (1.0 - sigma) * latent_image
        return out


def get_sigmas_simple(table: torch.Tensor, n: int) -> torch.Tensor:
    """"Simple" scheduler (modules/sd_schedulers.py:81-87), Forge's default for Flux."""
    ss = len(table) / n
    return torch.tensor([float(table[-(1 + int(x * ss))]) for x in range(n)] + [0.0], dtype=torch.float32)


def get_sigmas_uniform(pred: Prediction, n: int) -> torch.Tensor:
    """ForgeScheduleLinker.get_sigmas (k_diffusion/external.py:62-67): 'Automatic' for Euler / Euler a."""
    t = torch.linspace(len(pred.sigmas) - 1, 0, n)
    s = pred.sigma(t)
    return torch.cat([s, s.new_zeros([1])])


def get_sigmas_karras(n: int, sigma_min: float, sigma_max: float, rho: float = 7.0) -> torch.Tensor:
    """k_diffusion/sampling.py:19-25 (default schedule of DPM++ 2M, sd_samplers_kdiffusion.py:15,129)."""
    ramp = torch.linspace(0, 1, n)
    mi, ma = sigma_min ** (1 / rho), sigma_max ** (1 / rho)
    s = (ma + ramp * (mi - ma)) ** rho
    return torch.cat([s, s.new_zeros([1])])


def get_ancestral_step(sigma_from: float, sigma_to: float, eta: float = 1.0):
    """k_diffusion/sampling.py:53-60"""
    if not eta:
        return sigma_to, 0.0
    sigma_up = min(sigma_to, eta * (sigma_to ** 2 * (sigma_from ** 2 - sigma_to ** 2) / sigma_from ** 2) ** 0.5)
    sigma_down = (sigma_to ** 2 - sigma_up ** 2) ** 0.5
    return sigma_down, sigma_up


@dataclass
class StepPlan:
    """Host-precomputed scalars of one sampler step (what b200_step_desc carries)."""
    kind: int
    sigma: float
    dt: float = 0.0
    noise_scale: float = 0.0
    c_x: float = 0.0
    c_d: float = 0.0
    c_old: float = 0.0


def plan_euler(sigmas: Sequence[float]) -> List[StepPlan]:
    """sample_euler with s_churn = 0: x += (x - D)/sigma_i * (sigma_{i+1} - sigma_i)."""
    f = lambda v: torch.tensor(float(v), dtype=torch.float32)  # noqa: E731
    return [StepPlan(ops.STEP_EULER, float(f(sigmas[i])), dt=float(f(sigmas[i + 1]) - f(sigmas[i])))
            for i in range(len(sigmas) - 1)]


def plan_euler_ancestral(sigmas: Sequence[float], eta: float = 1.0, s_noise: float = 1.0) -> List[StepPlan]:
    """sample_euler_ancestral: dt = sigma_down - sigma_i; noise added with s_noise*sigma_up while sigma_{i+1} > 0.
    The reference evaluates these scalars with fp32 tensor arithmetic; the same roundings are applied here."""
    out = []
    f = lambda v: torch.tensor(v, dtype=torch.float32)  # noqa: E731
    for i in range(len(sigmas) - 1):
        sf, st = f(float(sigmas[i])), f(float(sigmas[i + 1]))
        if eta:
            up = torch.minimum(st, eta * (st ** 2 * (sf ** 2 - st ** 2) / sf ** 2) ** 0.5)
            down = (st ** 2 - up ** 2) ** 0.5
        else:
            up, down = f(0.0), st
        out.append(StepPlan(ops.STEP_EULER, float(sf), dt=float(down - sf),
                            noise_scale=float(s_noise * up) if float(st) > 0 else 0.0))
    return out


def plan_dpmpp_2m(sigmas: Sequence[float]) -> List[StepPlan]:
    """sample_dpmpp_2m (k_diffusion/sampling.py:648-671) as x' = c_x x + c_d D + c_old D_old."""
    out = []
    f = lambda v: torch.tensor(v, dtype=torch.float32)  # noqa: E731
    n = len(sigmas) - 1
    for i in range(n):
        s, sn = f(float(sigmas[i])), f(float(sigmas[i + 1]))
        t, t_next = -s.log(), -sn.log()
        h = t_next - t
        ratio = float((-t_next).exp() / (-t).exp())
        em = float(-(-h).expm1())
        if i == 0 or float(sn) == 0:
            out.append(StepPlan(ops.STEP_DPMPP_2M, float(s), c_x=ratio, c_d=em, c_old=0.0))
        else:
            h_last = t - (-f(float(sigmas[i - 1])).log())
            r = h_last / h
            out.append(StepPlan(ops.STEP_DPMPP_2M, float(s), c_x=ratio, c_d=em * float(1 + 1 / (2 * r)),
                                c_old=-em * float(1 / (2 * r))))
    return out


SAMPLERS = {
    # name -> (plan builder, default schedule, needs per-step noise)
    "euler": (plan_euler, "uniform", False),
    "euler_a": (plan_euler_ancestral, "uniform", True),
    "dpmpp_2m": (plan_dpmpp_2m, "karras", False),
}


def make_sigmas(pred: Prediction, sampler: str, steps: int) -> torch.Tensor:
    kind = SAMPLERS[sampler][1]
    if kind == "karras":
        return get_sigmas_karras(steps, pred.sigma_min, pred.sigma_max)
    return get_sigmas_uniform(pred, steps)


def run_sampler(eps_fn: Callable[[int], torch.Tensor], x: torch.Tensor, plan: List[StepPlan], *, cfg_scale: float,
                has_uncond: bool, noise_fn: Optional[Callable[[int], torch.Tensor]] = None,
                callback: Optional[Callable[[int, torch.Tensor, torch.Tensor], None]] = None,
                prediction: int = 0) -> torch.Tensor:
    """Drives the loop.  eps_fn(i) runs the UNet for step i on the current `x` (in place buffer) and returns
    the NHWC model output; each step then costs one fused launch.  `callback(i, x, denoised)` mirrors
    k-diffusion's callback dict (sd_samplers_common.py:265-272) with the pre-update x."""
    b, c, h, w = x.shape
    denoised = torch.empty_like(x)
    old = torch.zeros_like(x) if plan and plan[0].kind == ops.STEP_DPMPP_2M else None
    for i, st in enumerate(plan):
        eps = eps_fn(i)
        if callback is not None:
            x_before = x.clone()
        noise = noise_fn(i) if (st.noise_scale != 0.0 and noise_fn is not None) else None
        ops.sampler_step(x, eps, denoised, kind=st.kind, sigma=st.sigma, cfg_scale=cfg_scale, has_uncond=has_uncond,
                         dt=st.dt, noise=noise, noise_scale=st.noise_scale if noise is not None else 0.0,
                         old_denoised=old, c_x=st.c_x, c_d=st.c_d, c_old=st.c_old, prediction=prediction)
        if callback is not None:
            callback(i, x_before, denoised)
    return x
