"""GPU parity tests, model level: the fused channels-last UNet forward and the full denoise loop against
(a) golden vectors made by the imported reference on CPU fp32 (tests/golden, oracle/gen_golden.py) and
(b) the oracle evaluated in fp32 on the same device.

Stated tolerances (fp16 compute, fp32 accumulate): per-forward rel-RMS <= 3e-3 and max-abs <= 2e-2 of
the output RMS against the fp32 reference — the same order as the reference's own fp16-vs-fp32 gap, which the test
also measures with the oracle run in fp16; Euler / DPM++ 2M trajectories PSNR >= 40 dB on the final latent.
"""
import os

import pytest
import torch

from oracle import configs as CF
from oracle import sampling as S
from oracle import unet as OU
from tests.util import assert_close, err_stats

pytestmark = pytest.mark.gpu
DEV = "cuda"
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _gold(name):
    return torch.load(os.path.join(GOLD, name), weights_only=False)


def _engine(name, seed, dtype=torch.float16):
    from b200forge.unet_engine import UNetEngine
    cfg = CF.CONFIGS[name]
    sd = OU.random_state_dict(cfg, seed=seed)
    return UNetEngine(cfg, sd, dtype=dtype, device=DEV), cfg, sd


@pytest.mark.parametrize("name", ["tiny_xl", "tiny_15", "tiny_15h"])
def test_unet_forward_vs_reference_golden(name):
    g = _gold(f"unet_{name}.pt")
    eng, cfg, sd = _engine(name, g["weight_seed"])
    x = g["x"].to(DEV).half()
    y = None if g["y"] is None else g["y"].to(DEV).half()
    out = eng.forward(x, g["t"].to(DEV), g["context"].to(DEV).half(), y)
    torch.cuda.synchronize()
    assert_close(f"unet {name} fp16 engine vs reference fp32 golden", out, g["out"], max_rel=2e-2, rel_rms=3e-3)
    # how far the reference's own fp16 arithmetic is from fp32 on the same inputs (context for the tolerance)
    sd16 = {k: v.to(DEV).half() for k, v in sd.items()}
    ref16 = OU.unet_forward(sd16, cfg, x, g["t"].to(DEV), g["context"].to(DEV).half(), y)
    m, r = err_stats(ref16, g["out"])
    print(f"[parity] oracle-in-fp16 vs fp32 golden ({name}): max_abs={m:.3e} rel_rms={r:.3e}")


def test_unet_forward_bf16():
    g = _gold("unet_tiny_xl.pt")
    eng, cfg, sd = _engine("tiny_xl", g["weight_seed"], dtype=torch.bfloat16)
    out = eng.forward(g["x"].to(DEV).bfloat16(), g["t"].to(DEV), g["context"].to(DEV).bfloat16(), g["y"].to(DEV).bfloat16())
    torch.cuda.synchronize()
    assert_close("unet tiny_xl bf16 engine vs reference fp32 golden", out, g["out"], max_rel=1.6e-1, rel_rms=2.4e-2)  # 8x the fp16 bound (3 fewer mantissa bits)


@pytest.mark.parametrize("name,hw,batch", [("sdxl", 32, 2), ("sd15", 32, 2)])
def test_unet_full_width_vs_oracle_fp32(name, hw, batch):
    """Full-width SDXL / SD1.5 UNet (real channel counts and depths) on a small latent, against the oracle in fp32
    on the GPU (TF32 off) with the same fp16-rounded weights."""
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    cfg = CF.CONFIGS[name]
    sd = {k: v.half() for k, v in OU.random_state_dict(cfg, seed=11).items()}
    from b200forge.unet_engine import UNetEngine
    eng = UNetEngine(cfg, sd, dtype=torch.float16, device=DEV)
    g = torch.Generator().manual_seed(12)
    x = torch.randn(batch, 4, hw, hw, generator=g).half().to(DEV)
    ctx = torch.randn(batch, 77, cfg["context_dim"], generator=g).half().to(DEV)
    y = torch.randn(batch, cfg["adm_in_channels"], generator=g).half().to(DEV) if cfg["adm_in_channels"] else None
    t = torch.tensor([800.0, 100.0][:batch], device=DEV)
    out = eng.forward(x, t, ctx, y)  # sd15: head dims 40 / 80 (padded to 64 / 128) and 160 (GEMM-softmax-GEMM path)
    torch.cuda.synchronize()
    sd32 = {k: v.float().to(DEV) for k, v in sd.items()}
    with torch.no_grad():
        ref = OU.unet_forward(sd32, cfg, x.float(), t, ctx.float(), None if y is None else y.float())
    assert_close(f"unet {name} full width fp16 engine vs oracle fp32", out, ref, max_rel=2e-2, rel_rms=3e-3)


def _pipeline(g, use_graph):
    from b200forge.pipeline import Txt2ImgPipeline
    cfg = CF.CONFIGS[g["config"]]
    sd = OU.random_state_dict(cfg, seed=g["weight_seed"])
    return Txt2ImgPipeline(cfg, sd, dtype=torch.float16, device=DEV, use_graph=use_graph)


@pytest.mark.parametrize("use_graph", [False, True])
def test_first_step_denoised_vs_reference_golden(use_graph):
    """One Euler-a step: CFG-combined `denoised` from the fused step kernel vs the reference's
    sampling_function_inner output (golden)."""
    g = _gold("traj_tiny_xl.pt")
    pipe = _pipeline(g, use_graph)
    dens = []
    pipe.sample(g["cond"], g["uncond"], g["noise0"], sampler="euler_a", cfg_scale=g["cfg_scale"],
                sigmas=g["sigmas_auto"][:2], step_noise=g["euler_a_step_noise"][:1],
                callback=lambda i, x, d: dens.append(d.clone()))
    torch.cuda.synchronize()
    # denoised = x - sigma * eps with sigma_0 = 14.6 and CFG 7 amplify the UNet's relative error against x's scale
    assert_close("first-step denoised (CFG) vs reference golden", dens[0], g["euler_a_denoised0"], rel_rms=6e-3)


@pytest.mark.parametrize("sampler,key,sig", [("euler_a", "euler_a", "sigmas_auto"), ("euler", "euler", "sigmas_auto"),
                                             ("dpmpp_2m", "dpmpp_2m", "sigmas_karras")])
def test_trajectory_vs_reference_golden(sampler, key, sig):
    """Whole loops (6 steps, CFG 7, identical injected noise) vs the reference's k-diffusion loops on CPU fp32.
    The bound is on the final latent's PSNR relative to its own dynamic range (>= 40 dB)."""
    g = _gold("traj_tiny_xl.pt")
    pipe = _pipeline(g, True)
    x = pipe.sample(g["cond"], g["uncond"], g["noise0"], sampler=sampler, cfg_scale=g["cfg_scale"], sigmas=g[sig],
                    step_noise=g["euler_a_step_noise"] if sampler == "euler_a" else None)
    torch.cuda.synchronize()
    ref = g[key]
    mse = (x.float().cpu() - ref).pow(2).mean()
    peak = ref.abs().max()
    psnr = float(10 * torch.log10(peak ** 2 / mse))
    m, r = err_stats(x, ref)
    print(f"[parity] trajectory {sampler}: PSNR={psnr:.1f} dB max_abs={m:.3e} rel_rms={r:.3e}")
    assert psnr >= 40.0, psnr
    # the oracle itself in fp16-rounded weights/activations lands at a comparable distance
    assert torch.isfinite(x).all()


def test_sampler_loop_equals_oracle_loop_given_same_eps():
    """Isolates the fused CFG+update kernel from UNet rounding: feed the same synthetic eps sequence to the
    oracle loop and to run_sampler."""
    from b200forge import sampling as BS
    torch.manual_seed(0)
    B, C, H, W, steps = 2, 4, 16, 16, 8
    pred = S.EpsPrediction()
    sig = S.get_sigmas_uniform(pred, steps)
    eps_seq = [torch.randn(2 * B, H, W, 8).half() for _ in range(steps)]
    noise = torch.randn(steps, B, C, H, W)
    x0 = torch.randn(B, C, H, W) * float(sig[0])
    for sampler in ("euler", "euler_a", "dpmpp_2m"):
        sigs = S.get_sigmas_karras(steps, float(pred.sigma_min), float(pred.sigma_max)) if sampler == "dpmpp_2m" else sig
        it = iter(range(steps))

        def model(x, sigma):
            i = next(it)
            e = eps_seq[i][..., :C].permute(0, 3, 1, 2).float()
            return S.cfg_denoised_eps(x, e[:B], e[B:], float(sigma[0]), 7.0)
        if sampler == "euler":
            ref = S.sample_euler(model, x0.clone(), sigs)
        elif sampler == "euler_a":
            k = iter(range(steps))
            ref = S.sample_euler_ancestral(model, x0.clone(), sigs, lambda: noise[next(k)])
        else:
            ref = S.sample_dpmpp_2m(model, x0.clone(), sigs)
        builder = BS.SAMPLERS[sampler][0]
        plan = builder(sigs)
        xd = x0.to(DEV).clone()
        nd = noise.to(DEV)
        BS.run_sampler(lambda i: eps_seq[i].to(DEV), xd, plan, cfg_scale=7.0, has_uncond=True,
                       noise_fn=(lambda i: nd[i]) if sampler == "euler_a" else None)
        torch.cuda.synchronize()
        assert_close(f"loop {sampler} fused vs oracle", xd, ref, rel_rms=2e-6)


def test_img2img_vs_oracle():
    """img2img (sample_img2img, modules/sd_samplers_kdiffusion.py:136-194): pixels -> fused VAE encode + process_in ->
    noise * sigma_sched[0] + latent -> Euler over the last t_enc + 1 sigmas, vs the same steps through the oracle in fp32."""
    from b200forge.pipeline import Txt2ImgPipeline
    from oracle import vae as OV
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    cfg, vcfg = CF.CONFIGS["tiny_xl"], CF.VAE_CONFIGS["tiny"]
    sd = OU.random_state_dict(cfg, seed=1)
    esd = OV.random_encoder_state_dict(vcfg, seed=8)
    pipe = Txt2ImgPipeline(cfg, sd, vae_cfg=vcfg, vae_state_dict=OV.random_state_dict(vcfg, seed=3),
                           vae_encoder_state_dict=esd, dtype=torch.float16, vae_dtype=torch.float16, device=DEV)
    g = torch.Generator().manual_seed(5)
    B, steps, strength = 2, 10, 0.5
    px = torch.rand(B, 64, 64, 3, generator=g)
    vnoise = torch.randn(B, 4, 32, 32, generator=g)
    noise = torch.randn(B, 4, 32, 32, generator=g)
    cond = dict(crossattn=torch.randn(B, 77, cfg["context_dim"], generator=g), vector=torch.randn(B, cfg["adm_in_channels"], generator=g))
    uncond = dict(crossattn=torch.randn(B, 77, cfg["context_dim"], generator=g), vector=torch.randn(B, cfg["adm_in_channels"], generator=g))
    x = pipe.img2img(cond, uncond, px, noise, steps=steps, denoising_strength=strength, sampler="euler", cfg_scale=5.0, vae_noise=vnoise)
    torch.cuda.synchronize()
    pred = S.EpsPrediction()
    full = S.get_sigmas_uniform(pred, steps)
    t_enc = int(min(strength, 0.999) * steps)
    sched = full[steps - t_enc - 1:]
    assert len(sched) == t_enc + 2 and float(sched[-1]) == 0.0
    with torch.no_grad():
        latent = OV.encode_first_stage(esd, vcfg, px, vnoise)
        den = S.Denoiser(lambda xc, t, cx, yy: OU.unet_forward(sd, cfg, xc, t, cx, yy), pred, cond, uncond, 5.0)  # CPU fp32
        ref = S.sample_euler(den, noise * sched[0] + latent, sched)
    mse = (x.cpu() - ref).pow(2).mean()
    psnr = float(10 * torch.log10(ref.abs().max() ** 2 / mse))
    m, r = err_stats(x, ref)
    print(f"[parity] img2img euler: PSNR={psnr:.1f} dB max_abs={m:.3e} rel_rms={r:.3e}")
    assert psnr >= 40.0, psnr
