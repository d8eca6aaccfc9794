"""Plug point P4 — k-diffusion sampler functions with the reference's exact call contract
(k_diffusion/sampling.py:119-137 sample_euler, :140-159 sample_euler_ancestral, :648-671 sample_dpmpp_2m, and the
two-evaluation samplers :188-214 sample_heun, :217-246 sample_dpm_2, :249-276 sample_dpm_2_ancestral,
:573-603 sample_dpmpp_2s_ancestral; multistep / SDE: :310-345 sample_lms, :606-646 sample_dpmpp_sde,
:675-724 sample_dpmpp_2m_sde, :727-778 sample_dpmpp_3m_sde):

    fn(model, x, sigmas, extra_args=None, callback=None, disable=None, ...) -> x

`model` is whatever Forge passes (CFGDenoiser.forward, modules/sd_samplers_cfg_denoiser.py:156-228) and is called
unchanged; what changes is the per-step arithmetic around it: the reference evaluates 4-8 tiny fp32 tensor kernels
plus 2-3 device->host syncs per step on 0-dim GPU sigmas, here the schedule is read to the host ONCE and each step's
update is a single `b200_sampler_update` launch.  Contract details that are preserved:
  * `callback({'x','i','sigma','sigma_hat','denoised'})` every step with the pre-update x (drives progress/interrupt);
  * noise is drawn through k-diffusion's module-global `torch.randn_like` (TorchHijack -> ImageRNG.next(),
    modules/sd_samplers_common.py:214-235) in the same order and count as the reference — sample_euler draws one
    (unused when s_churn = 0) tensor per step, sample_euler_ancestral draws only while sigma_next > 0;
  * `modules/sd_schedulers.py:10-15` to_d override: d = (x - denoised) / sigma.
Cases outside the fused path (s_churn > 0, Flux rectified-flow variant) defer to `reference_*` when the plug-in
recorded the originals, else raise.
"""
from __future__ import annotations

import math

import torch

from . import lib as _l
from . import ops, sampling

reference_sample_euler = None
reference_sample_euler_ancestral = None
reference_sample_dpmpp_2m = None
reference_sample_heun = None
reference_sample_dpm_2 = None
reference_sample_dpm_2_ancestral = None
reference_sample_dpmpp_2s_ancestral = None
reference_sample_lms = None
reference_sample_dpmpp_sde = None
reference_sample_dpmpp_2m_sde = None
reference_sample_dpmpp_3m_sde = None


def _randn_like(x):
    try:  # k-diffusion's own (possibly hijacked) torch handle keeps the per-image seeded RNG stream
        import k_diffusion.sampling as ks  # type: ignore
        return ks.torch.randn_like(x)
    except Exception:
        return torch.randn_like(x)


def _host_sigmas(sigmas):
    return [float(v) for v in sigmas.detach().float().cpu().tolist()]


def _fusable(x):
    return x.is_cuda and x.dtype == torch.float32 and x.dim() == 4


def _prep(x):
    return x if x.is_contiguous() else x.contiguous()


@torch.no_grad()
def sample_euler(model, x, sigmas, extra_args=None, callback=None, disable=None, s_churn=0., s_tmin=0.,
                 s_tmax=float('inf'), s_noise=1.):
    if s_churn > 0 or not _fusable(x):
        if reference_sample_euler is None:
            raise _l.B200Error(_l.E_UNSUPPORTED, "sample_euler: s_churn > 0 / non-CUDA-fp32 latents need the reference sampler")
        return reference_sample_euler(model, x, sigmas, extra_args, callback, disable, s_churn, s_tmin, s_tmax, s_noise)
    extra_args = {} if extra_args is None else extra_args
    s_in = x.new_ones([x.shape[0]])
    sig = _host_sigmas(sigmas)
    plan = sampling.plan_euler(sig)
    x = _prep(x).clone()
    for i, st in enumerate(plan):
        _randn_like(x)  # the reference draws eps every step even when gamma == 0 (sampling.py:126): keep the RNG stream aligned
        denoised = model(x, sigmas[i] * s_in, **extra_args)
        if callback is not None:
            callback({'x': x, 'i': i, 'sigma': sigmas[i], 'sigma_hat': sigmas[i], 'denoised': denoised})
        ops.sampler_update(x, _prep(denoised.float()), kind=ops.STEP_EULER, sigma=st.sigma, dt=st.dt)
    return x


@torch.no_grad()
def sample_euler_ancestral(model, x, sigmas, extra_args=None, callback=None, disable=None, eta=1., s_noise=1.,
                           noise_sampler=None):
    if not _fusable(x):
        if reference_sample_euler_ancestral is None:
            raise _l.B200Error(_l.E_UNSUPPORTED, "sample_euler_ancestral: non-CUDA-fp32 latents need the reference sampler")
        return reference_sample_euler_ancestral(model, x, sigmas, extra_args, callback, disable, eta, s_noise, noise_sampler)
    if _is_flux(model):  # the reference dispatches on the predictor type (k_diffusion/sampling.py:143-144)
        return _sample_euler_ancestral_rf(model, x, sigmas, extra_args, callback, eta, s_noise, noise_sampler)
    extra_args = {} if extra_args is None else extra_args
    if noise_sampler is None:
        noise_sampler = lambda sigma, sigma_next: _randn_like(x)  # noqa: E731  (default_noise_sampler, sampling.py:63-64)
    s_in = x.new_ones([x.shape[0]])
    sig = _host_sigmas(sigmas)
    plan = sampling.plan_euler_ancestral(sig, eta, s_noise)
    x = _prep(x).clone()
    for i, st in enumerate(plan):
        denoised = model(x, sigmas[i] * s_in, **extra_args)
        if callback is not None:
            callback({'x': x, 'i': i, 'sigma': sigmas[i], 'sigma_hat': sigmas[i], 'denoised': denoised})
        noise = None
        if sig[i + 1] > 0:
            noise = _prep(noise_sampler(sigmas[i], sigmas[i + 1]).float())
        ops.sampler_update(x, _prep(denoised.float()), kind=ops.STEP_EULER, sigma=st.sigma, dt=st.dt, noise=noise,
                           noise_scale=st.noise_scale if noise is not None else 0.0)
    return x


@torch.no_grad()
def sample_dpmpp_2m(model, x, sigmas, extra_args=None, callback=None, disable=None):
    if not _fusable(x):
        if reference_sample_dpmpp_2m is None:
            raise _l.B200Error(_l.E_UNSUPPORTED, "sample_dpmpp_2m: non-CUDA-fp32 latents need the reference sampler")
        return reference_sample_dpmpp_2m(model, x, sigmas, extra_args, callback, disable)
    extra_args = {} if extra_args is None else extra_args
    s_in = x.new_ones([x.shape[0]])
    sig = _host_sigmas(sigmas)
    plan = sampling.plan_dpmpp_2m(sig)
    x = _prep(x).clone()
    old = torch.zeros_like(x)
    for i, st in enumerate(plan):
        denoised = model(x, sigmas[i] * s_in, **extra_args)
        if callback is not None:
            callback({'x': x, 'i': i, 'sigma': sigmas[i], 'sigma_hat': sigmas[i], 'denoised': denoised})
        ops.sampler_update(x, _prep(denoised.float()), kind=ops.STEP_DPMPP_2M, sigma=max(st.sigma, 1e-30), old_denoised=old,
                           c_x=st.c_x, c_d=st.c_d, c_old=st.c_old)
    return x


# ------------------------------------------------------------------------------------------------------------------
# Two-evaluation samplers.  Scalars follow the reference's fp32 tensor arithmetic (computed here
# on fp32 CPU scalars, once); each stage's tensor update is one launch: the first stage is the Euler kernel on a copy of
# x, the second a 4-operand linear combination (B200_STEP_LINEAR).
def _f32(v):
    return torch.tensor(float(v), dtype=torch.float32)


def _ancestral(sig_from, sig_to, eta):
    """get_ancestral_step (k_diffusion/sampling.py:53-60) in fp32 like the reference's 0-dim tensors."""
    if not eta:
        return sig_to, _f32(0.0)
    up = torch.minimum(sig_to, eta * (sig_to ** 2 * (sig_from ** 2 - sig_to ** 2) / sig_from ** 2) ** 0.5)
    down = (sig_to ** 2 - up ** 2) ** 0.5
    return down, up


def _is_flux(model):
    try:
        from backend.modules.k_prediction import PredictionFlux  # type: ignore
        return isinstance(model.inner_model.predictor, PredictionFlux)
    except Exception:
        return False


def _defer(ref, name, *args):
    if ref is None:
        raise _l.B200Error(_l.E_UNSUPPORTED, f"{name}: this case needs the reference sampler")
    return ref(*args)


@torch.no_grad()
def sample_heun(model, x, sigmas, extra_args=None, callback=None, disable=None, s_churn=0., s_tmin=0.,
                s_tmax=float('inf'), s_noise=1.):
    """k_diffusion/sampling.py:188-214 (Karras Algorithm 2) with s_churn = 0."""
    if s_churn > 0 or not _fusable(x):
        return _defer(reference_sample_heun, "sample_heun", model, x, sigmas, extra_args, callback, disable, s_churn, s_tmin, s_tmax, s_noise)
    extra_args = {} if extra_args is None else extra_args
    s_in = x.new_ones([x.shape[0]])
    sig = [_f32(v) for v in _host_sigmas(sigmas)]
    x = _prep(x).clone()
    for i in range(len(sig) - 1):
        _randn_like(x)  # eps is drawn every step (sampling.py:196) even when unused
        s0, s1 = sig[i], sig[i + 1]
        denoised = _prep(model(x, sigmas[i] * s_in, **extra_args).float())
        if callback is not None:
            callback({'x': x, 'i': i, 'sigma': sigmas[i], 'sigma_hat': sigmas[i], 'denoised': denoised})
        dt = float(s1 - s0)
        if float(s1) == 0:
            ops.sampler_update(x, denoised, kind=ops.STEP_EULER, sigma=float(s0), dt=dt)
        else:
            x2 = x.clone()
            ops.sampler_update(x2, denoised, kind=ops.STEP_EULER, sigma=float(s0), dt=dt)          # x_2 = x + d * dt
            denoised2 = _prep(model(x2, sigmas[i + 1] * s_in, **extra_args).float())
            # x + (d + d_2)/2 * dt,  d = (x - D)/s0,  d_2 = (x_2 - D_2)/s1
            a, b = dt / (2.0 * float(s0)), dt / (2.0 * float(s1))
            ops.sampler_update(x, denoised2, kind=ops.STEP_LINEAR, sigma=1.0, c_x=1.0 + a, c_d=-b, old_denoised=denoised,
                               c_old=-a, noise=x2, noise_scale=b)
    return x


def _dpm2_stage(model, x, denoised, s0, s_target, sigmas_i_dev, s_in, extra_args):
    """Shared DPM-Solver-2 stage (sampling.py:237-244 / 266-274): midpoint in log-sigma, second evaluation, full step."""
    sigma_mid = s0.log().lerp(s_target.log(), 0.5).exp()
    dt1, dt2 = float(sigma_mid - s0), float(s_target - s0)
    x2 = x.clone()
    ops.sampler_update(x2, denoised, kind=ops.STEP_EULER, sigma=float(s0), dt=dt1)                  # x_2 = x + d * dt_1
    denoised2 = _prep(model(x2, float(sigma_mid) * s_in, **extra_args).float())
    c = dt2 / float(sigma_mid)
    ops.sampler_update(x, denoised2, kind=ops.STEP_LINEAR, sigma=1.0, c_x=1.0, c_d=-c, noise=x2, noise_scale=c)  # x + d_2 * dt_2


@torch.no_grad()
def sample_dpm_2(model, x, sigmas, extra_args=None, callback=None, disable=None, s_churn=0., s_tmin=0.,
                 s_tmax=float('inf'), s_noise=1.):
    """k_diffusion/sampling.py:217-246 with s_churn = 0."""
    if s_churn > 0 or not _fusable(x):
        return _defer(reference_sample_dpm_2, "sample_dpm_2", model, x, sigmas, extra_args, callback, disable, s_churn, s_tmin, s_tmax, s_noise)
    extra_args = {} if extra_args is None else extra_args
    s_in = x.new_ones([x.shape[0]])
    sig = [_f32(v) for v in _host_sigmas(sigmas)]
    x = _prep(x).clone()
    for i in range(len(sig) - 1):
        _randn_like(x)
        s0, s1 = sig[i], sig[i + 1]
        denoised = _prep(model(x, sigmas[i] * s_in, **extra_args).float())
        if callback is not None:
            callback({'x': x, 'i': i, 'sigma': sigmas[i], 'sigma_hat': sigmas[i], 'denoised': denoised})
        if float(s1) == 0:
            ops.sampler_update(x, denoised, kind=ops.STEP_EULER, sigma=float(s0), dt=float(s1 - s0))
        else:
            _dpm2_stage(model, x, denoised, s0, s1, sigmas[i], s_in, extra_args)
    return x


@torch.no_grad()
def sample_dpm_2_ancestral(model, x, sigmas, extra_args=None, callback=None, disable=None, eta=1., s_noise=1.,
                           noise_sampler=None):
    """k_diffusion/sampling.py:249-276."""
    if not _fusable(x):
        return _defer(reference_sample_dpm_2_ancestral, "sample_dpm_2_ancestral", model, x, sigmas, extra_args, callback, disable, eta, s_noise, noise_sampler)
    if _is_flux(model):  # k_diffusion/sampling.py:251-252
        return _sample_dpm_2_ancestral_rf(model, x, sigmas, extra_args, callback, eta, s_noise, noise_sampler)
    extra_args = {} if extra_args is None else extra_args
    if noise_sampler is None:
        noise_sampler = lambda sigma, sigma_next: _randn_like(x)  # noqa: E731
    s_in = x.new_ones([x.shape[0]])
    sig = [_f32(v) for v in _host_sigmas(sigmas)]
    x = _prep(x).clone()
    for i in range(len(sig) - 1):
        s0, s1 = sig[i], sig[i + 1]
        denoised = _prep(model(x, sigmas[i] * s_in, **extra_args).float())
        down, up = _ancestral(s0, s1, eta)
        if callback is not None:
            callback({'x': x, 'i': i, 'sigma': sigmas[i], 'sigma_hat': sigmas[i], 'denoised': denoised})
        if float(down) == 0:
            ops.sampler_update(x, denoised, kind=ops.STEP_EULER, sigma=float(s0), dt=float(down - s0))
        else:
            _dpm2_stage(model, x, denoised, s0, down, sigmas[i], s_in, extra_args)
            noise = _prep(noise_sampler(sigmas[i], sigmas[i + 1]).float())
            ops.sampler_update(x, noise, kind=ops.STEP_LINEAR, sigma=1.0, c_x=1.0, c_d=float(s_noise * up))
    return x


@torch.no_grad()
def sample_dpmpp_2s_ancestral(model, x, sigmas, extra_args=None, callback=None, disable=None, eta=1., s_noise=1.,
                              noise_sampler=None):
    """k_diffusion/sampling.py:573-603."""
    if _is_flux(model) or not _fusable(x):
        return _defer(reference_sample_dpmpp_2s_ancestral, "sample_dpmpp_2s_ancestral", model, x, sigmas, extra_args, callback, disable, eta, s_noise, noise_sampler)
    extra_args = {} if extra_args is None else extra_args
    if noise_sampler is None:
        noise_sampler = lambda sigma, sigma_next: _randn_like(x)  # noqa: E731
    s_in = x.new_ones([x.shape[0]])
    sig = [_f32(v) for v in _host_sigmas(sigmas)]
    x = _prep(x).clone()
    for i in range(len(sig) - 1):
        s0, s1 = sig[i], sig[i + 1]
        denoised = _prep(model(x, sigmas[i] * s_in, **extra_args).float())
        down, up = _ancestral(s0, s1, eta)
        if callback is not None:
            callback({'x': x, 'i': i, 'sigma': sigmas[i], 'sigma_hat': sigmas[i], 'denoised': denoised})
        if float(down) == 0:
            ops.sampler_update(x, denoised, kind=ops.STEP_EULER, sigma=float(s0), dt=float(down - s0))
        else:
            t, t_next = s0.log().neg(), down.log().neg()
            h = t_next - t
            s = t + 0.5 * h
            sig_s = s.neg().exp()
            x2 = x.clone()
            ops.sampler_update(x2, denoised, kind=ops.STEP_LINEAR, sigma=1.0, c_x=float(sig_s / t.neg().exp()),
                               c_d=float(-(-h * 0.5).expm1()))
            denoised2 = _prep(model(x2, float(sig_s) * s_in, **extra_args).float())
            ops.sampler_update(x, denoised2, kind=ops.STEP_LINEAR, sigma=1.0, c_x=float(t_next.neg().exp() / t.neg().exp()),
                               c_d=float(-(-h).expm1()))
        if float(s1) > 0:
            noise = _prep(noise_sampler(sigmas[i], sigmas[i + 1]).float())
            ops.sampler_update(x, noise, kind=ops.STEP_LINEAR, sigma=1.0, c_x=1.0, c_d=float(s_noise * up))
    return x


# ------------------------------------------------------------------------------------------------------------------
# Multistep / SDE samplers.  Every update is a linear combination of x, up to two denoised tensors and one noise tensor
# with host-computed coefficients (fp32 scalar arithmetic like the reference's 0-dim tensors), i.e. one or two
# B200_STEP_LINEAR launches per model evaluation.  The noise sampler is whatever Forge passes (BrownianTreeNoiseSampler,
# modules/sd_samplers_kdiffusion.py:196-213); without one the reference function builds its own tree, so that case is
# handed back.
def _lin(x, d, *, c_x, c_d, old=None, c_old=0.0, noise=None, c_noise=0.0):
    ops.sampler_update(x, d, kind=ops.STEP_LINEAR, sigma=1.0, c_x=float(c_x), c_d=float(c_d), old_denoised=old,
                       c_old=float(c_old) if old is not None else 0.0, noise=noise,
                       noise_scale=float(c_noise) if noise is not None else 0.0)


@torch.no_grad()
def sample_lms(model, x, sigmas, extra_args=None, callback=None, disable=None, order=4):
    """k_diffusion/sampling.py:310-345 (coefficients by the same scipy quadrature, :298-307)."""
    if not _fusable(x):
        return _defer(reference_sample_lms, "sample_lms", model, x, sigmas, extra_args, callback, disable, order)
    from scipy import integrate
    extra_args = {} if extra_args is None else extra_args
    s_in = x.new_ones([x.shape[0]])
    t = sigmas.detach().cpu().numpy()

    def coeff(cur_order, i, j):
        def fn(tau):
            prod = 1.
            for k in range(cur_order):
                if j == k:
                    continue
                prod *= (tau - t[i - k]) / (t[i - j] - t[i - k])
            return prod
        return integrate.quad(fn, t[i], t[i + 1], epsrel=1e-4)[0]

    x = _prep(x).clone()
    ds = []
    for i in range(len(t) - 1):
        denoised = _prep(model(x, sigmas[i] * s_in, **extra_args).float())
        d = x.clone()
        inv = 1.0 / float(t[i])
        _lin(d, denoised, c_x=inv, c_d=-inv)                      # d = (x - denoised) / sigma
        ds.append(d)
        if len(ds) > order:
            ds.pop(0)
        if callback is not None:
            callback({'x': x, 'i': i, 'sigma': sigmas[i], 'sigma_hat': sigmas[i], 'denoised': denoised})
        cur_order = min(i + 1, order)
        cs = [coeff(cur_order, i, j) for j in range(cur_order)]
        hist = list(reversed(ds))[:cur_order]
        _lin(x, hist[0], c_x=1.0, c_d=cs[0], old=hist[1] if cur_order > 1 else None, c_old=cs[1] if cur_order > 1 else 0.0,
             noise=hist[2] if cur_order > 2 else None, c_noise=cs[2] if cur_order > 2 else 0.0)
        if cur_order > 3:
            _lin(x, hist[3], c_x=1.0, c_d=cs[3])
    return x


@torch.no_grad()
def sample_dpmpp_sde(model, x, sigmas, extra_args=None, callback=None, disable=None, eta=1., s_noise=1., noise_sampler=None,
                     r=1 / 2):
    """k_diffusion/sampling.py:606-646."""
    if noise_sampler is None or not _fusable(x):
        return _defer(reference_sample_dpmpp_sde, "sample_dpmpp_sde", model, x, sigmas, extra_args, callback, disable, eta, s_noise, noise_sampler, r)
    extra_args = {} if extra_args is None else extra_args
    s_in = x.new_ones([x.shape[0]])
    sig = [_f32(v) for v in _host_sigmas(sigmas)]
    x = _prep(x).clone()
    for i in range(len(sig) - 1):
        s0, s1 = sig[i], sig[i + 1]
        denoised = _prep(model(x, sigmas[i] * s_in, **extra_args).float())
        if callback is not None:
            callback({'x': x, 'i': i, 'sigma': sigmas[i], 'sigma_hat': sigmas[i], 'denoised': denoised})
        if float(s1) == 0:
            ops.sampler_update(x, denoised, kind=ops.STEP_EULER, sigma=float(s0), dt=float(s1 - s0))
            continue
        t, t_next = s0.log().neg(), s1.log().neg()
        h = t_next - t
        s = t + h * r
        fac = 1 / (2 * r)
        sig_t, sig_s, sig_n = t.neg().exp(), s.neg().exp(), t_next.neg().exp()
        # step 1
        sd, su = _ancestral(sig_t, sig_s, eta)
        s_ = sd.log().neg()
        x2 = x.clone()
        n1 = _prep(noise_sampler(float(sig_t) * s_in[0], float(sig_s) * s_in[0]).float())
        _lin(x2, denoised, c_x=s_.neg().exp() / sig_t, c_d=-(t - s_).expm1(), noise=n1, c_noise=s_noise * su)
        denoised2 = _prep(model(x2, float(sig_s) * s_in, **extra_args).float())
        # step 2
        sd, su = _ancestral(sig_t, sig_n, eta)
        tn_ = sd.log().neg()
        b = -(t - tn_).expm1()
        n2 = _prep(noise_sampler(float(sig_t) * s_in[0], float(sig_n) * s_in[0]).float())
        _lin(x, denoised2, c_x=tn_.neg().exp() / sig_t, c_d=b * fac, old=denoised, c_old=b * (1 - fac), noise=n2,
             c_noise=s_noise * su)
    return x


@torch.no_grad()
def sample_dpmpp_2m_sde(model, x, sigmas, extra_args=None, callback=None, disable=None, eta=1., s_noise=1.,
                        noise_sampler=None, solver_type='midpoint'):
    """k_diffusion/sampling.py:675-724."""
    if solver_type not in {'heun', 'midpoint'}:
        raise ValueError('solver_type must be \'heun\' or \'midpoint\'')
    if (noise_sampler is None and eta) or not _fusable(x):
        return _defer(reference_sample_dpmpp_2m_sde, "sample_dpmpp_2m_sde", model, x, sigmas, extra_args, callback, disable, eta, s_noise, noise_sampler, solver_type)
    extra_args = {} if extra_args is None else extra_args
    s_in = x.new_ones([x.shape[0]])
    sig = [_f32(v) for v in _host_sigmas(sigmas)]
    x = _prep(x).clone()
    old, h_last = None, None
    for i in range(len(sig) - 1):
        s0, s1 = sig[i], sig[i + 1]
        denoised = _prep(model(x, sigmas[i] * s_in, **extra_args).float())
        if callback is not None:
            callback({'x': x, 'i': i, 'sigma': sigmas[i], 'sigma_hat': sigmas[i], 'denoised': denoised})
        if float(s1) == 0:
            _lin(x, denoised, c_x=0.0, c_d=1.0)                   # x = denoised
        else:
            t, s = -s0.log(), -s1.log()
            h = s - t
            eta_h = eta * h
            c_x = s1 / s0 * (-eta_h).exp()
            c_d = (-h - eta_h).expm1().neg()
            c2 = _f32(0.0)
            if old is not None:
                r = h_last / h
                if solver_type == 'heun':
                    c2 = ((-h - eta_h).expm1().neg() / (-h - eta_h) + 1) * (1 / r)
                else:
                    c2 = 0.5 * (-h - eta_h).expm1().neg() * (1 / r)
            noise, c_n = None, 0.0
            if eta:
                noise = _prep(noise_sampler(sigmas[i], sigmas[i + 1]).float())
                c_n = s1 * (-2 * eta_h).expm1().neg().sqrt() * s_noise
            _lin(x, denoised, c_x=c_x, c_d=c_d + c2, old=old, c_old=-c2, noise=noise, c_noise=c_n)
            h_last = h
        old = denoised
    return x


@torch.no_grad()
def sample_dpmpp_3m_sde(model, x, sigmas, extra_args=None, callback=None, disable=None, eta=1., s_noise=1.,
                        noise_sampler=None):
    """k_diffusion/sampling.py:727-778."""
    if (noise_sampler is None and eta) or not _fusable(x):
        return _defer(reference_sample_dpmpp_3m_sde, "sample_dpmpp_3m_sde", model, x, sigmas, extra_args, callback, disable, eta, s_noise, noise_sampler)
    extra_args = {} if extra_args is None else extra_args
    s_in = x.new_ones([x.shape[0]])
    sig = [_f32(v) for v in _host_sigmas(sigmas)]
    x = _prep(x).clone()
    d_1, d_2, h_1, h_2 = None, None, None, None
    for i in range(len(sig) - 1):
        s0, s1 = sig[i], sig[i + 1]
        denoised = _prep(model(x, sigmas[i] * s_in, **extra_args).float())
        if callback is not None:
            callback({'x': x, 'i': i, 'sigma': sigmas[i], 'sigma_hat': sigmas[i], 'denoised': denoised})
        if float(s1) == 0:
            _lin(x, denoised, c_x=0.0, c_d=1.0)
        else:
            t, s = -s0.log(), -s1.log()
            h = s - t
            h_eta = h * (eta + 1)
            c_x = torch.exp(-h_eta)
            c_d = (-h_eta).expm1().neg()
            c_1, c_2 = _f32(0.0), _f32(0.0)                       # coefficients of denoised_1, denoised_2
            if h_2 is not None:
                r0, r1 = h_1 / h, h_2 / h
                phi_2 = h_eta.neg().expm1() / h_eta + 1
                phi_3 = phi_2 / h_eta - 0.5
                # x += phi_2 * d1 - phi_3 * d2 with d1, d2 the divided differences of (denoised, denoised_1, denoised_2)
                a = phi_2
                bc = (phi_2 * r0 - phi_3) / (r0 + r1)
                c_d = c_d + (a + bc) / r0
                c_1 = -(a + bc) / r0 - bc / r1
                c_2 = bc / r1
            elif h_1 is not None:
                r = h_1 / h
                phi_2 = h_eta.neg().expm1() / h_eta + 1
                c_d = c_d + phi_2 / r
                c_1 = -phi_2 / r
            noise, c_n = None, 0.0
            if eta:
                noise = _prep(noise_sampler(sigmas[i], sigmas[i + 1]).float())
                c_n = s1 * (-2 * h * eta).expm1().neg().sqrt() * s_noise
            _lin(x, denoised, c_x=c_x, c_d=c_d, old=d_1, c_old=c_1, noise=noise, c_noise=c_n)
            if d_2 is not None and float(c_2) != 0.0:
                _lin(x, d_2, c_x=1.0, c_d=c_2)
            h_1, h_2 = h, h_1
        d_1, d_2 = denoised, d_1
    return x


# ------------------------------------------------------------------------------------------------------------------
# Rectified-flow variants the reference switches to for Flux (k_diffusion/sampling.py:162-186, 278-309): the down-step and
# the re-noising coefficient come from alpha = 1 - sigma instead of the variance-exploding ancestral step.
def _rf_coeffs(s0, s1, eta):
    downstep_ratio = 1 + (s1 / s0 - 1) * eta
    sigma_down = s1 * downstep_ratio
    alpha_ip1, alpha_down = 1 - s1, 1 - sigma_down
    renoise = (s1 ** 2 - sigma_down ** 2 * alpha_ip1 ** 2 / alpha_down ** 2) ** 0.5
    return sigma_down, alpha_ip1 / alpha_down, renoise


def _sample_euler_ancestral_rf(model, x, sigmas, extra_args, callback, eta, s_noise, noise_sampler):
    extra_args = {} if extra_args is None else extra_args
    if noise_sampler is None:
        noise_sampler = lambda sigma, sigma_next: _randn_like(x)  # noqa: E731
    s_in = x.new_ones([x.shape[0]])
    sig = [_f32(v) for v in _host_sigmas(sigmas)]
    x = _prep(x).clone()
    for i in range(len(sig) - 1):
        s0, s1 = sig[i], sig[i + 1]
        denoised = _prep(model(x, sigmas[i] * s_in, **extra_args).float())
        if callback is not None:
            callback({'x': x, 'i': i, 'sigma': sigmas[i], 'sigma_hat': sigmas[i], 'denoised': denoised})
        if float(s1) == 0:
            _lin(x, denoised, c_x=0.0, c_d=1.0)
            continue
        sigma_down, a, renoise = _rf_coeffs(s0, s1, eta)
        r = sigma_down / s0
        if eta > 0:   # x = a * (r x + (1 - r) D) + noise * s_noise * renoise
            noise = _prep(noise_sampler(sigmas[i], sigmas[i + 1]).float())
            _lin(x, denoised, c_x=a * r, c_d=a * (1 - r), noise=noise, c_noise=s_noise * renoise)
        else:
            _lin(x, denoised, c_x=r, c_d=1 - r)
    return x


def _sample_dpm_2_ancestral_rf(model, x, sigmas, extra_args, callback, eta, s_noise, noise_sampler):
    extra_args = {} if extra_args is None else extra_args
    if noise_sampler is None:
        noise_sampler = lambda sigma, sigma_next: _randn_like(x)  # noqa: E731
    s_in = x.new_ones([x.shape[0]])
    sig = [_f32(v) for v in _host_sigmas(sigmas)]
    x = _prep(x).clone()
    for i in range(len(sig) - 1):
        s0, s1 = sig[i], sig[i + 1]
        denoised = _prep(model(x, sigmas[i] * s_in, **extra_args).float())
        sigma_down, a, renoise = _rf_coeffs(s0, s1, eta)
        if callback is not None:
            callback({'x': x, 'i': i, 'sigma': sigmas[i], 'sigma_hat': sigmas[i], 'denoised': denoised})
        if float(sigma_down) == 0:
            ops.sampler_update(x, denoised, kind=ops.STEP_EULER, sigma=float(s0), dt=float(sigma_down - s0))
        else:
            _dpm2_stage(model, x, denoised, s0, sigma_down, sigmas[i], s_in, extra_args)
            noise = _prep(noise_sampler(sigmas[i], sigmas[i + 1]).float())
            _lin(x, noise, c_x=a, c_d=s_noise * renoise)       # x = (alpha_ip1 / alpha_down) x + noise * s_noise * renoise
    return x


# ---------------------------------------------------------------------------------------------------------------------
# Linear multistep samplers over the derivative history d_i = (x_i - denoised_i) / sigma_i (k_diffusion/sampling.py:771-978).
# Every update is x <- x + sum_k c_k d_{i-k}: the coefficients are host scalars, the history lives in device buffers, one
# 4-operand launch covers orders up to 3 and a second one adds the fourth history term.
reference_sample_heunpp2 = None
reference_sample_ipndm = None
reference_sample_ipndm_v = None
reference_sample_deis = None


def _to_d(x, denoised, sigma: float):
    d = x.clone()
    inv = 1.0 / float(sigma)
    _lin(d, denoised, c_x=inv, c_d=-inv)  # (x - denoised) / sigma
    return d


def _multistep_update(x, hist, cs):
    """x += sum_k cs[k] * hist[k]  (hist[0] = newest derivative); 1 launch for <= 3 terms, 2 for 4."""
    n = len(cs)
    _lin(x, hist[0], c_x=1.0, c_d=cs[0], old=hist[1] if n > 1 else None, c_old=cs[1] if n > 1 else 0.0,
         noise=hist[2] if n > 2 else None, c_noise=cs[2] if n > 2 else 0.0)
    if n > 3:
        _lin(x, hist[3], c_x=1.0, c_d=cs[3])


def _multistep(model, x, sigmas, extra_args, callback, max_order, coeffs_fn):
    """Shared loop of iPNDM / iPNDM-v / DEIS: coeffs_fn(i, order, t) -> [c_cur, c_prev1, ...] multiplying the derivatives."""
    extra_args = {} if extra_args is None else extra_args
    s_in = x.new_ones([x.shape[0]])
    t = _host_sigmas(sigmas)
    x = _prep(x).clone()
    hist = []
    for i in range(len(t) - 1):
        denoised = _prep(model(x, sigmas[i] * s_in, **extra_args).float())
        if callback is not None:
            callback({'x': x, 'i': i, 'sigma': sigmas[i], 'sigma_hat': sigmas[i], 'denoised': denoised})
        hist.insert(0, _to_d(x, denoised, t[i]))
        del hist[max_order:]
        order = min(max_order, i + 1)
        _multistep_update(x, hist, coeffs_fn(i, order, t))
    return x


@torch.no_grad()
def sample_ipndm(model, x, sigmas, extra_args=None, callback=None, disable=None, max_order=4):
    """k_diffusion/sampling.py:829-865: Adams-Bashforth weights on a uniform-step assumption."""
    if not _fusable(x):
        return _defer(reference_sample_ipndm, "sample_ipndm", model, x, sigmas, extra_args, callback, disable, max_order)
    ab = {1: [1.0], 2: [3 / 2, -1 / 2], 3: [23 / 12, -16 / 12, 5 / 12], 4: [55 / 24, -59 / 24, 37 / 24, -9 / 24]}

    def coeffs(i, order, t):
        h = _f32(t[i + 1]) - _f32(t[i])  # the reference forms (t_next - t_cur) in fp32
        return [float(h) * c for c in ab[order]]

    return _multistep(model, x, sigmas, extra_args, callback, max_order, coeffs)


@torch.no_grad()
def sample_ipndm_v(model, x, sigmas, extra_args=None, callback=None, disable=None, max_order=4):
    """k_diffusion/sampling.py:869-926: the variable-step Adams-Bashforth weights."""
    if not _fusable(x):
        return _defer(reference_sample_ipndm_v, "sample_ipndm_v", model, x, sigmas, extra_args, callback, disable, max_order)

    def coeffs(i, order, t):
        f = torch.float32
        tt = [torch.tensor(v, dtype=f) for v in t]
        h_n = tt[i + 1] - tt[i]
        if order == 1:
            return [float(h_n)]
        h_n_1 = tt[i] - tt[i - 1]
        if order == 2:
            c1 = (2 + (h_n / h_n_1)) / 2
            c2 = -(h_n / h_n_1) / 2
            cs = [c1, c2]
        elif order == 3:
            h_n_2 = tt[i - 1] - tt[i - 2]
            temp = (1 - h_n / (3 * (h_n + h_n_1)) * (h_n * (h_n + h_n_1)) / (h_n_1 * (h_n_1 + h_n_2))) / 2
            cs = [(2 + (h_n / h_n_1)) / 2 + temp, -(h_n / h_n_1) / 2 - (1 + h_n_1 / h_n_2) * temp, temp * h_n_1 / h_n_2]
        else:
            h_n_2 = tt[i - 1] - tt[i - 2]
            h_n_3 = tt[i - 2] - tt[i - 3]
            temp1 = (1 - h_n / (3 * (h_n + h_n_1)) * (h_n * (h_n + h_n_1)) / (h_n_1 * (h_n_1 + h_n_2))) / 2
            temp2 = ((1 - h_n / (3 * (h_n + h_n_1))) / 2 + (1 - h_n / (2 * (h_n + h_n_1))) * h_n / (6 * (h_n + h_n_1 + h_n_2))) \
                * (h_n * (h_n + h_n_1) * (h_n + h_n_1 + h_n_2)) / (h_n_1 * (h_n_1 + h_n_2) * (h_n_1 + h_n_2 + h_n_3))
            r12 = h_n_1 * (h_n_1 + h_n_2) / (h_n_2 * (h_n_2 + h_n_3))
            cs = [(2 + (h_n / h_n_1)) / 2 + temp1 + temp2,
                  -(h_n / h_n_1) / 2 - (1 + h_n_1 / h_n_2) * temp1 - (1 + (h_n_1 / h_n_2) + r12) * temp2,
                  temp1 * h_n_1 / h_n_2 + ((h_n_1 / h_n_2) + r12 * (1 + h_n_2 / h_n_3)) * temp2,
                  -temp2 * r12 * h_n_1 / h_n_2]
        return [float(h_n * c) for c in cs]

    return _multistep(model, x, sigmas, extra_args, callback, max_order, coeffs)


def deis_coeff_list(t_steps: torch.Tensor, max_order: int, N: int = 10000):
    """k_diffusion/deis.py:56-83 ('tab' mode): per step, the integrals of (Lagrange basis) x (-1/2 dlog(alpha)/dtau /
    sqrt(alpha (1 - alpha))) over [t_cur, t_next] in the VP time that edm2t maps the sigmas to (:13-20).  dlog(alpha)/dtau
    is written out (-(tau (beta_1 - beta_0) + beta_0)) instead of being taken by autograd; same N-point rectangle sum."""
    sig = t_steps.detach().float().cpu()
    eps_s, smin, smax = 1e-3, 0.002, 80.0
    beta_d = 2 * (torch.log(torch.tensor(smin) ** 2 + 1) / eps_s - torch.log(torch.tensor(smax) ** 2 + 1)) / (eps_s - 1)
    beta_min = torch.log(torch.tensor(smax) ** 2 + 1) - 0.5 * beta_d
    ts = ((beta_min ** 2 + 2 * beta_d * (sig ** 2 + 1).log()).sqrt() - beta_min) / beta_d
    beta_0, beta_1 = beta_min, beta_d + beta_min
    C = []
    for i in range(len(ts) - 1):
        order = min(i + 1, max_order)
        if order == 1:
            C.append([])
            continue
        t_cur, t_next = ts[i], ts[i + 1]
        taus = torch.linspace(float(t_cur), float(t_next), N)
        dtau = (t_next - t_cur) / N
        prev_t = ts[[i - k for k in range(order)]]
        alpha = torch.exp(-0.5 * taus ** 2 * (beta_1 - beta_0) - taus * beta_0)
        integrand = -0.5 * (-(taus * (beta_1 - beta_0) + beta_0)) / torch.sqrt(alpha * (1 - alpha))
        row = []
        for j in range(order):
            poly = 1
            for k in range(order):
                if k != j:
                    poly = poly * (taus - prev_t[k]) / (prev_t[j] - prev_t[k])
            row.append(float(torch.sum(integrand * poly) * dtau))
        C.append(row)
    return C


@torch.no_grad()
def sample_deis(model, x, sigmas, extra_args=None, callback=None, disable=None, max_order=3, deis_mode='tab'):
    """k_diffusion/sampling.py:933-978."""
    if deis_mode != 'tab' or not _fusable(x):
        return _defer(reference_sample_deis, "sample_deis", model, x, sigmas, extra_args, callback, disable, max_order, deis_mode)
    C = deis_coeff_list(sigmas, max_order)

    def coeffs(i, order, t):
        if t[i + 1] <= 0 or order == 1:
            return [float(_f32(t[i + 1]) - _f32(t[i]))]
        return C[i]

    return _multistep(model, x, sigmas, extra_args, callback, max_order, coeffs)


@torch.no_grad()
def sample_heunpp2(model, x, sigmas, extra_args=None, callback=None, disable=None, s_churn=0., s_tmin=0.,
                   s_tmax=float('inf'), s_noise=1.):
    """k_diffusion/sampling.py:771-824 with s_churn = 0: Euler on the last step, weighted Heun on the one before, else a
    three-evaluation step whose derivative is w1 d + w2 d_2 + w3 d_3 with w_k = sigma_k / (3 sigma_0)."""
    if s_churn > 0 or not _fusable(x):
        return _defer(reference_sample_heunpp2, "sample_heunpp2", model, x, sigmas, extra_args, callback, disable, s_churn, s_tmin, s_tmax, s_noise)
    extra_args = {} if extra_args is None else extra_args
    s_in = x.new_ones([x.shape[0]])
    sig = [_f32(v) for v in _host_sigmas(sigmas)]
    s_end = sig[-1]
    x = _prep(x).clone()
    for i in range(len(sig) - 1):
        _randn_like(x)  # the reference draws eps every step (unused when gamma == 0): keep the RNG stream aligned
        denoised = _prep(model(x, sigmas[i] * s_in, **extra_args).float())
        if callback is not None:
            callback({'x': x, 'i': i, 'sigma': sigmas[i], 'sigma_hat': sigmas[i], 'denoised': denoised})
        dt = sig[i + 1] - sig[i]
        if sig[i + 1] == s_end:
            ops.sampler_update(x, denoised, kind=ops.STEP_EULER, sigma=float(sig[i]), dt=float(dt))
            continue
        d = _to_d(x, denoised, float(sig[i]))
        x2 = x.clone()
        _lin(x2, d, c_x=1.0, c_d=float(dt))
        den2 = _prep(model(x2, sigmas[i + 1] * s_in, **extra_args).float())
        d2 = _to_d(x2, den2, float(sig[i + 1]))
        if sig[i + 2] == s_end:
            w = 2 * sig[0]
            w2 = sig[i + 1] / w
            w1 = 1 - w2
            _lin(x, d, c_x=1.0, c_d=float(w1 * dt), old=d2, c_old=float(w2 * dt))
        else:
            dt2 = sig[i + 2] - sig[i + 1]
            x3 = x2  # x_3 = x_2 + d_2 dt_2 (x_2 is not needed afterwards)
            _lin(x3, d2, c_x=1.0, c_d=float(dt2))
            den3 = _prep(model(x3, sigmas[i + 2] * s_in, **extra_args).float())
            d3 = _to_d(x3, den3, float(sig[i + 2]))
            w = 3 * sig[0]
            w2 = sig[i + 1] / w
            w3 = sig[i + 2] / w
            w1 = 1 - w2 - w3
            _lin(x, d, c_x=1.0, c_d=float(w1 * dt), old=d2, c_old=float(w2 * dt), noise=d3, c_noise=float(w3 * dt))
    return x


# ---------------------------------------------------------------------------------------------------------------------
# Forge's extra samplers (modules/sd_samplers_extra.py:7-74 Restart, modules/sd_samplers_lcm.py:68-82 LCM).  They live in
# Forge's `modules/` package and are registered by it; the fused versions keep the same call contracts.
reference_restart_sampler = None
reference_sample_lcm = None


@torch.no_grad()
def restart_sampler(model, x, sigmas, extra_args=None, callback=None, disable=None, s_noise=1., restart_list=None):
    """modules/sd_samplers_extra.py:7-74 — Restart sampling (Xu et al. 2023): Heun steps over a Karras schedule with
    re-noising jumps back to a higher sigma.  The schedule surgery is the reference's own (host scalars); every Heun step is
    two model evaluations + three launches, the re-noising one launch."""
    if not _fusable(x):
        return _defer(reference_restart_sampler, "restart_sampler", model, x, sigmas, extra_args, callback, disable, s_noise, restart_list)
    extra_args = {} if extra_args is None else extra_args
    s_in = x.new_ones([x.shape[0]])
    steps = sigmas.shape[0] - 1
    sig = sigmas.detach().float().cpu()
    if restart_list is None:
        if steps >= 20:
            restart_steps = 9
            restart_times = 1
            if steps >= 36:
                restart_steps = steps // 4
                restart_times = 2
            sig = sampling.get_sigmas_karras(steps - restart_steps * restart_times, sig[-2].item(), sig[0].item())
            restart_list = {0.1: [restart_steps + 1, restart_times, 2]}
        else:
            restart_list = {}
    restart_list = {int(torch.argmin(abs(sig - key), dim=0)): value for key, value in restart_list.items()}
    step_list = []
    for i in range(len(sig) - 1):
        step_list.append((sig[i], sig[i + 1]))
        if i + 1 in restart_list:
            restart_steps, restart_times, restart_max = restart_list[i + 1]
            min_idx = i + 1
            max_idx = int(torch.argmin(abs(sig - restart_max), dim=0))
            if max_idx < min_idx:
                sigma_restart = sampling.get_sigmas_karras(restart_steps, sig[min_idx].item(), sig[max_idx].item())[:-1]
                while restart_times > 0:
                    restart_times -= 1
                    step_list.extend(zip(sigma_restart[:-1], sigma_restart[1:]))
    x = _prep(x).clone()
    dev_sigma = lambda s: s.to(x.device) * s_in  # noqa: E731
    last_sigma = None
    step_id = 0
    for old_sigma, new_sigma in step_list:
        if last_sigma is None:
            last_sigma = old_sigma
        elif last_sigma < old_sigma:
            noise = _prep(_randn_like(x).float())
            _lin(x, noise, c_x=1.0, c_d=float(s_noise * (old_sigma ** 2 - last_sigma ** 2) ** 0.5))
        denoised = _prep(model(x, dev_sigma(old_sigma), **extra_args).float())
        if callback is not None:
            callback({'x': x, 'i': step_id, 'sigma': new_sigma, 'sigma_hat': old_sigma, 'denoised': denoised})
        dt = new_sigma - old_sigma
        if float(new_sigma) == 0:
            ops.sampler_update(x, denoised, kind=ops.STEP_EULER, sigma=float(old_sigma), dt=float(dt))
        else:
            d = _to_d(x, denoised, float(old_sigma))
            x2 = x.clone()
            _lin(x2, d, c_x=1.0, c_d=float(dt))
            den2 = _prep(model(x2, dev_sigma(new_sigma), **extra_args).float())
            d2 = _to_d(x2, den2, float(new_sigma))
            _lin(x, d, c_x=1.0, c_d=float(dt / 2), old=d2, c_old=float(dt / 2))   # x + (d + d_2) / 2 * dt
        step_id += 1
        last_sigma = new_sigma
    return x


@torch.no_grad()
def sample_lcm(model, x, sigmas, extra_args=None, callback=None, disable=None, noise_sampler=None):
    """modules/sd_samplers_lcm.py:68-82: x <- denoised (+ sigma_next * noise while sigma_next > 0)."""
    if not _fusable(x):
        return _defer(reference_sample_lcm, "sample_lcm", model, x, sigmas, extra_args, callback, disable, noise_sampler)
    extra_args = {} if extra_args is None else extra_args
    s_in = x.new_ones([x.shape[0]])
    sig = _host_sigmas(sigmas)
    x = _prep(x).clone()
    for i in range(len(sig) - 1):
        denoised = _prep(model(x, sigmas[i] * s_in, **extra_args).float())
        if callback is not None:
            callback({'x': x, 'i': i, 'sigma': sigmas[i], 'sigma_hat': sigmas[i], 'denoised': denoised})
        noise = None
        if sig[i + 1] > 0:
            noise = noise_sampler(sigmas[i], sigmas[i + 1]) if noise_sampler is not None else _randn_like(x)
        # x = denoised + sigma_next * noise as one launch: 0 * x + 1 * denoised (+ noise term)
        ops.sampler_update(x, denoised, kind=ops.STEP_LINEAR, sigma=1.0, c_x=0.0, c_d=1.0, noise=noise,
                           noise_scale=float(sig[i + 1]) if noise is not None else 0.0)
    return x
