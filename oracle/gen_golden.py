"""ORACLE tooling — generates tests/golden/*.pt by running the UNMODIFIED reference (imported from
/root/reference through oracle/ref_import.py) on CPU fp32 with deterministic synthetic weights.

    python -m oracle.gen_golden            # in the build container (needs /root/reference)

The fixtures pin (a) the oracle restatement and (b) the CUDA path on the GPU box, where the reference tree
does not exist.  Weights are not stored: `oracle.unet.random_state_dict(cfg, seed)` regenerates them
bit-identically (CPU torch.Generator), and each fixture stores a checksum of the weights it was made with.
"""
from __future__ import annotations

import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import configs as CF  # noqa: E402
from oracle import ref_import  # noqa: E402
from oracle import unet as OU  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def sd_checksum(sd) -> float:
    return float(sum(v.double().abs().sum() for v in sd.values()))


def make_inputs(cfg, B, hw, seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, 4, hw, hw, generator=g)
    ctx = torch.randn(B, 77, cfg["context_dim"], generator=g)
    y = torch.randn(B, cfg["adm_in_channels"], generator=g) if cfg.get("adm_in_channels") else None
    return x, ctx, y


def gen_unet(name: str, hw: int = 16):
    from backend.nn.unet import IntegratedUNet2DConditionModel as RefUNet
    cfg = CF.CONFIGS[name]
    sd = OU.random_state_dict(cfg, seed=1)
    m = RefUNet(**cfg).eval()
    m.load_state_dict(sd, strict=True)
    x, ctx, y = make_inputs(cfg, 2, hw, seed=2)
    t = torch.tensor([981.0, 23.0])
    with torch.no_grad():
        out = m(x, t, context=ctx, y=y, transformer_options={})
    torch.save(dict(config=name, weight_seed=1, weight_checksum=sd_checksum(sd), x=x, t=t, context=ctx, y=y, out=out),
               os.path.join(GOLD, f"unet_{name}.pt"))
    print("unet", name, "out std", out.std().item())


def gen_v_trajectory(name: str = "tiny_21", hw: int = 16, steps: int = 5):
    """SD2.x-style run: 4-level linear-transformer UNet without label_emb, v-prediction, Euler, CFG 6 — the reference's
    KModel.apply_model -> sampling_function_inner -> k_diffusion.sample_euler on CPU fp32."""
    import k_diffusion.sampling as ks
    from backend.modules.k_model import KModel
    from backend.modules.k_prediction import Prediction
    from backend.nn.unet import IntegratedUNet2DConditionModel as RefUNet
    from backend.sampling.condition import compile_conditions
    from backend.sampling.sampling_function import sampling_function_inner
    from k_diffusion.external import ForgeScheduleLinker
    ks.to_d = lambda x, sigma, denoised: (x - denoised) / sigma
    cfg = CF.CONFIGS[name]
    sd = OU.random_state_dict(cfg, seed=1)
    unet = RefUNet(**cfg).eval()
    unet.load_state_dict(sd, strict=True)
    unet.storage_dtype = torch.float32
    unet.computation_dtype = torch.float32
    pred = Prediction(prediction_type="v_prediction")
    kmodel = KModel(unet, diffusers_scheduler=None, k_predictor=pred)
    B = 2
    g = torch.Generator().manual_seed(17)
    cond = dict(crossattn=torch.randn(B, 77, cfg["context_dim"], generator=g))
    uncond = dict(crossattn=torch.randn(B, 77, cfg["context_dim"], generator=g))
    # SD1.x / 2.x conditioning is a bare tensor (condition.py:95-103): no pooled vector
    cond_c, uncond_c = compile_conditions(cond["crossattn"]), compile_conditions(uncond["crossattn"])
    cfg_scale = 6.0

    class Wrap:
        class _Inner:
            predictor = pred
        inner_model = _Inner()

        def __call__(self, x, sigma, **kw):
            return sampling_function_inner(kmodel, x, sigma, uncond_c, cond_c, cfg_scale, {}, None)

    sig = ForgeScheduleLinker(pred).get_sigmas(steps)
    noise0 = torch.randn(B, 4, hw, hw, generator=g)
    with torch.no_grad():
        # modules/sd_samplers_kdiffusion.py:207 with opts.sgm_noise_multiplier at its default False (shared_options.py:410)
        x0 = pred.noise_scaling(sig[0], noise0.clone(), torch.zeros_like(noise0), max_denoise=False)
        x = make_inputs(cfg, B, hw, seed=2)[0]
        fwd = unet(x, torch.tensor([981.0, 23.0]), context=cond["crossattn"], y=None, transformer_options={})
        dens = []
        out = ks.sample_euler(Wrap(), x0.clone(), sig, extra_args={}, callback=lambda d: dens.append(d["denoised"].clone()), disable=True)
    torch.save(dict(config=name, weight_seed=1, weight_checksum=sd_checksum(sd), cond=cond, uncond=uncond, cfg_scale=cfg_scale,
                    sigmas=sig, noise0=noise0, x0=x0, euler=out, denoised0=dens[0], fwd_x=x, fwd_t=torch.tensor([981.0, 23.0]),
                    fwd_out=fwd), os.path.join(GOLD, f"traj_{name}_v.pt"))
    print("v-pred traj", name, float(out.std()), float(fwd.std()))


def gen_trajectories(name: str = "tiny_xl", hw: int = 16, steps: int = 6):
    """Reference denoise loop: KModel.apply_model -> sampling_function_inner (CFG) -> k_diffusion sample_*."""
    import k_diffusion.sampling as ks
    from backend.modules.k_model import KModel
    from backend.modules.k_prediction import Prediction
    from backend.nn.unet import IntegratedUNet2DConditionModel as RefUNet
    from backend.sampling.condition import compile_conditions
    from backend.sampling.sampling_function import sampling_function_inner
    from k_diffusion.external import ForgeScheduleLinker

    # modules/sd_schedulers.py:10-15 replaces k_diffusion.sampling.to_d at import time; `modules` cannot be
    # imported here (gradio etc. missing), so the same override is applied by hand.
    ks.to_d = lambda x, sigma, denoised: (x - denoised) / sigma

    cfg = CF.CONFIGS[name]
    sd = OU.random_state_dict(cfg, seed=1)
    unet = RefUNet(**cfg).eval()
    unet.load_state_dict(sd, strict=True)
    unet.storage_dtype = torch.float32
    unet.computation_dtype = torch.float32
    pred = Prediction(prediction_type="epsilon")
    kmodel = KModel(unet, diffusers_scheduler=None, k_predictor=pred)
    linker = ForgeScheduleLinker(pred)

    B = 2
    g = torch.Generator().manual_seed(7)
    cond = dict(crossattn=torch.randn(B, 77, cfg["context_dim"], generator=g),
                vector=torch.randn(B, cfg["adm_in_channels"], generator=g))
    uncond = dict(crossattn=torch.randn(B, 77, cfg["context_dim"], generator=g),
                  vector=torch.randn(B, cfg["adm_in_channels"], generator=g))
    cond_c, uncond_c = compile_conditions(cond), compile_conditions(uncond)
    cfg_scale = 7.0

    class Wrap:  # what k-diffusion sees as `model`: CFGDenoiser minus UI glue (sd_samplers_cfg_denoiser.py:199)
        class _Inner:
            predictor = pred
        inner_model = _Inner()

        def __call__(self, x, sigma, **kw):
            return sampling_function_inner(kmodel, x, sigma, uncond_c, cond_c, cfg_scale, {}, None)

    seeds = [1000, 1001]
    gens = [torch.Generator().manual_seed(s) for s in seeds]

    def draw():  # modules/rng.py:167-177 ImageRNG.next(): one randn per image generator, stacked
        return torch.stack([torch.randn((4, hw, hw), generator=gg) for gg in gens])

    sig_auto = linker.get_sigmas(steps)                       # Euler / Euler a "Automatic" schedule
    sig_karras = ks.get_sigmas_karras(steps, float(pred.sigma_min), float(pred.sigma_max))
    out = dict(config=name, weight_seed=1, weight_checksum=sd_checksum(sd), cond=cond, uncond=uncond,
               cfg_scale=cfg_scale, seeds=seeds, hw=hw, steps=steps, sigmas_auto=sig_auto, sigmas_karras=sig_karras)
    with torch.no_grad():
        noise0 = draw()
        # modules/sd_samplers_kdiffusion.py:207: max_denoise = opts.sgm_noise_multiplier, default False (shared_options.py:410)
        x0 = pred.noise_scaling(sig_auto[0], noise0.clone(), torch.zeros_like(noise0), max_denoise=False)
        out["noise0"] = noise0
        out["x0"] = x0
        step_noise = []

        class Hijack:  # TorchHijack (modules/sd_samplers_common.py:214-235): randn_like -> ImageRNG.next
            @staticmethod
            def randn_like(x):
                n = draw()
                step_noise.append(n)
                return n

            def __getattr__(self, item):
                return getattr(torch, item)

        ks.torch = Hijack()
        try:
            dens = []
            cb = lambda d: dens.append(d["denoised"].clone())  # noqa: E731
            out["euler_a"] = ks.sample_euler_ancestral(Wrap(), x0.clone(), sig_auto, extra_args={}, callback=cb, disable=True)
            out["euler_a_step_noise"] = torch.stack(step_noise)
            out["euler_a_denoised0"] = dens[0]
            out["euler_a_denoised_last"] = dens[-1]
            step_noise.clear()
            out["euler"] = ks.sample_euler(Wrap(), x0.clone(), sig_auto, extra_args={}, disable=True)
            x0k = pred.noise_scaling(sig_karras[0], noise0.clone(), torch.zeros_like(noise0), max_denoise=False)
            out["x0_karras"] = x0k
            out["dpmpp_2m"] = ks.sample_dpmpp_2m(Wrap(), x0k.clone(), sig_karras, extra_args={}, disable=True)
            # one run with the "SGM noise multiplier" option on (max_denoise=True)
            x0s = pred.noise_scaling(sig_auto[0], noise0.clone(), torch.zeros_like(noise0), max_denoise=True)
            out["x0_sgm"] = x0s
            out["euler_sgm"] = ks.sample_euler(Wrap(), x0s.clone(), sig_auto, extra_args={}, disable=True)
        finally:
            ks.torch = torch
    torch.save(out, os.path.join(GOLD, f"traj_{name}.pt"))
    print("traj", name, {k: float(v.std()) for k, v in out.items() if k in ("euler_a", "euler", "dpmpp_2m")})


def gen_schedules():
    import k_diffusion.sampling as ks
    from backend.modules.k_prediction import Prediction
    from k_diffusion.external import ForgeScheduleLinker
    pred = Prediction(prediction_type="epsilon")
    linker = ForgeScheduleLinker(pred)
    probe = torch.tensor([14.6146, 10.0, 3.3, 1.0, 0.5, 0.1, 0.0292])
    out = dict(sigmas=pred.sigmas.clone(), auto20=linker.get_sigmas(20), auto30=linker.get_sigmas(30),
               karras30=ks.get_sigmas_karras(30, float(pred.sigma_min), float(pred.sigma_max)),
               probe=probe, probe_timestep=pred.timestep(probe))
    torch.save(out, os.path.join(GOLD, "schedules.pt"))
    print("schedules sigma_min/max", float(pred.sigma_min), float(pred.sigma_max))


def gen_vae(name: str = "tiny", hw: int = 16):
    from backend.nn.vae import IntegratedAutoencoderKL
    from oracle import vae as OV
    cfg = CF.VAE_CONFIGS[name]
    sd = OV.random_state_dict(cfg, seed=3)
    m = IntegratedAutoencoderKL(**{k: v for k, v in cfg.items()}).eval()
    missing = m.load_state_dict(sd, strict=False)
    assert not missing.unexpected_keys, missing
    assert all(k.startswith(("encoder.", "quant_conv.")) for k in missing.missing_keys), missing.missing_keys
    g = torch.Generator().manual_seed(4)
    z = torch.randn(2, cfg["latent_channels"], hw, hw, generator=g)
    with torch.no_grad():
        out = m.decode(m.process_out(z))
    torch.save(dict(config=name, weight_seed=3, weight_checksum=sd_checksum(sd), z=z, out=out),
               os.path.join(GOLD, f"vae_{name}.pt"))
    print("vae", name, "out std", out.std().item())


def gen_vae_tiled(name: str = "tiny"):
    """The reference's tiled decode: its own `tiled_scale` (backend/patcher/vae.py:11-57) around its own
    IntegratedAutoencoderKL.decode, composed as VAE.decode_tiled_ composes them (:104-115; the VAE wrapper class itself needs
    the memory manager and a loaded model, the three-line composition is restated here)."""
    from backend.nn.vae import IntegratedAutoencoderKL
    from backend.patcher.vae import tiled_scale
    from oracle import vae as OV
    cfg = CF.VAE_CONFIGS[name]
    sd = OV.random_state_dict(cfg, seed=3)
    m = IntegratedAutoencoderKL(**{k: v for k, v in cfg.items()}).eval()
    m.load_state_dict(sd, strict=False)
    g = torch.Generator().manual_seed(6)
    z = torch.randn(2, cfg["latent_channels"], 20, 28, generator=g)
    tile_x = tile_y = 8
    overlap = 2
    up = 2 ** (len(cfg["block_out_channels"]) - 1)
    fn = lambda a: (m.decode(a) + 1.0).float()  # noqa: E731
    with torch.no_grad():
        zz = m.process_out(z)
        out = torch.clamp(((tiled_scale(zz, fn, tile_x // 2, tile_y * 2, overlap, upscale_amount=up) +
                            tiled_scale(zz, fn, tile_x * 2, tile_y // 2, overlap, upscale_amount=up) +
                            tiled_scale(zz, fn, tile_x, tile_y, overlap, upscale_amount=up)) / 3.0) / 2.0, min=0.0, max=1.0)
    torch.save(dict(config=name, weight_seed=3, weight_checksum=sd_checksum(sd), z=z, tile_x=tile_x, tile_y=tile_y, overlap=overlap,
                    out=out.movedim(1, -1)), os.path.join(GOLD, f"vae_tiled_{name}.pt"))
    print("vae tiled", name, "out std", out.std().item(), tuple(out.shape))


def gen_samplers(steps: int = 7):
    """The reference's own k-diffusion loops (k_diffusion/sampling.py) on CPU fp32 around oracle.sampling.toy_denoiser,
    with the to_d override of modules/sd_schedulers.py:10-15 and a recorded noise stream."""
    import k_diffusion.sampling as ks
    from backend.modules.k_prediction import Prediction
    from oracle import sampling as OS
    ks.to_d = lambda x, sigma, denoised: (x - denoised) / sigma
    pred = Prediction(prediction_type="epsilon")
    sig = ks.get_sigmas_karras(steps, float(pred.sigma_min), float(pred.sigma_max))
    g = torch.Generator().manual_seed(11)
    x0 = torch.randn(2, 4, 16, 16, generator=g) * sig[0]
    noise = torch.randn(4 * steps, 2, 4, 16, 16, generator=g)

    class Model:
        class _Inner:
            predictor = pred
        inner_model = _Inner()

        def __call__(self, x, sigma, **kw):
            return OS.toy_denoiser(x, sigma)

    out = dict(sigmas=sig, x0=x0, noise=noise)
    runs = [("sample_heun", {}), ("sample_dpm_2", {}), ("sample_dpm_2_ancestral", {}), ("sample_dpmpp_2s_ancestral", {}),
            ("sample_lms", {}), ("sample_dpmpp_sde", {}), ("sample_dpmpp_2m_sde", {}),
            ("sample_dpmpp_2m_sde", {"solver_type": "heun"}), ("sample_dpmpp_3m_sde", {}),
            ("sample_heunpp2", {}), ("sample_ipndm", {}), ("sample_ipndm_v", {}), ("sample_deis", {})]
    for name, extra in runs:
        k = iter(range(noise.shape[0]))
        kw = dict(extra)
        if "ancestral" in name or "sde" in name:
            kw["noise_sampler"] = lambda s, sn: noise[next(k)]
        key = name + ("_heun" if extra.get("solver_type") == "heun" else "")
        with torch.no_grad():
            out[key] = getattr(ks, name)(Model(), x0.clone(), sig, extra_args={}, disable=True, **kw)
        print(key, float(out[key].std()))
    # Restart (modules/sd_samplers_extra.py imports only torch / tqdm / k_diffusion, so the reference file itself is run):
    # 24 steps so that the automatic restart list is non-empty (steps >= 20); noise through k_diffusion.sampling.torch
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_sd_samplers_extra", os.path.join(ref_import.REF_ROOT, "modules", "sd_samplers_extra.py"))
    extra = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(extra)
    sig24 = ks.get_sigmas_karras(24, float(pred.sigma_min), float(pred.sigma_max))
    x24 = torch.randn(2, 4, 16, 16, generator=torch.Generator().manual_seed(13)) * sig24[0]
    rn = torch.randn(8, 2, 4, 16, 16, generator=torch.Generator().manual_seed(14))
    kk = iter(range(8))

    class Hijack:
        @staticmethod
        def randn_like(x):
            return rn[next(kk)]

        def __getattr__(self, item):
            return getattr(torch, item)

    ks.torch = Hijack()
    try:
        with torch.no_grad():
            out["restart_sampler"] = extra.restart_sampler(Model(), x24.clone(), sig24, extra_args={}, disable=True)
    finally:
        ks.torch = torch
    out["restart_sigmas"], out["restart_x0"], out["restart_noise"] = sig24, x24, rn
    print("restart_sampler", float(out["restart_sampler"].std()), "noise draws used", next(kk))
    # rectified-flow variants: the reference dispatches on isinstance(model.inner_model.predictor, PredictionFlux)
    from backend.modules.k_prediction import PredictionFlux
    from oracle import sampling as OS2
    fpred = PredictionFlux()
    fsig = OS2.simple_scheduler(steps, OS2.flux_sigma_table())

    class FluxModel(Model):
        class _Inner:
            predictor = fpred
        inner_model = _Inner()

    xf0 = torch.randn(2, 4, 16, 16, generator=torch.Generator().manual_seed(12)) * fsig[0]
    out["flux_sigmas"], out["flux_x0"] = fsig, xf0
    for name in ("sample_euler_ancestral", "sample_dpm_2_ancestral"):
        k = iter(range(noise.shape[0]))
        with torch.no_grad():
            out[name + "_rf"] = getattr(ks, name)(FluxModel(), xf0.clone(), fsig, extra_args={}, disable=True,
                                                  noise_sampler=lambda s, sn: noise[next(k)])
        print(name + "_rf", float(out[name + "_rf"].std()))
    torch.save(out, os.path.join(GOLD, "samplers_toy.pt"))


def gen_vae_encode(name: str = "tiny", hw: int = 64):
    """Reference IntegratedAutoencoderKL.encode (backend/nn/vae.py:293-303): moments via a `regulation` hook, and the
    default .sample() with the global CPU generator seeded (the reference draws torch.randn(mean.shape) there)."""
    from backend.nn.vae import IntegratedAutoencoderKL
    from oracle import vae as OV
    cfg = CF.VAE_CONFIGS[name]
    sd = OV.random_encoder_state_dict(cfg, seed=8)
    m = IntegratedAutoencoderKL(**{k: v for k, v in cfg.items()}).eval()
    missing = m.load_state_dict(sd, strict=False)
    assert not missing.unexpected_keys, missing
    assert all(k.startswith(("decoder.", "post_quant_conv.")) for k in missing.missing_keys), missing.missing_keys
    g = torch.Generator().manual_seed(9)
    pixels = torch.rand(2, hw, hw, 3, generator=g)                       # NHWC in [0, 1], what VAE.encode receives
    x = 2.0 * pixels.movedim(-1, 1) - 1.0                                # patcher/vae.py:177
    with torch.no_grad():
        mean, logvar = m.encode(x, regulation=lambda p: (p.mean, p.logvar))
        torch.manual_seed(1234)
        sample = m.encode(x)
        torch.manual_seed(1234)
        noise = torch.randn(mean.shape)
    latent = m.process_in(sample)
    torch.save(dict(config=name, weight_seed=8, weight_checksum=sd_checksum(sd), pixels=pixels, mean=mean, logvar=logvar,
                    noise=noise, sample=sample, latent=latent), os.path.join(GOLD, f"vae_enc_{name}.pt"))
    print("vae encode", name, "mean std", mean.std().item(), "logvar mean", logvar.mean().item())


def gen_unet_control(name: str = "tiny_xl", hw: int = 16):
    """Reference UNet forward with ControlNet-style residuals (backend/nn/unet.py:44-52, 714, 733, 739): one tensor per input
    block, one for the middle block, one per output skip, plus a None entry, consumed from the end of each list."""
    from backend.nn.unet import IntegratedUNet2DConditionModel as RefUNet
    cfg = CF.CONFIGS[name]
    sd = OU.random_state_dict(cfg, seed=1)
    m = RefUNet(**cfg).eval()
    m.load_state_dict(sd, strict=True)
    x, ctx, y = make_inputs(cfg, 2, hw, seed=2)
    t = torch.tensor([981.0, 23.0])
    # shapes of the activations the residuals are added to: run once with hooks
    shapes = {"input": [], "middle": [], "output": []}
    hooks = [blk.register_forward_hook(lambda mod, i, o, k="input": shapes[k].append(tuple(o.shape))) for blk in m.input_blocks]
    hooks.append(m.middle_block.register_forward_hook(lambda mod, i, o: shapes["middle"].append(tuple(o.shape))))
    with torch.no_grad():
        m(x, t, context=ctx, y=y, transformer_options={})
    for h in hooks:
        h.remove()
    g = torch.Generator().manual_seed(19)
    ins = [torch.randn(s, generator=g) * 0.3 for s in shapes["input"]]
    control = {"input": list(reversed(ins)),                                   # popped from the end: block 0 first
               "middle": [torch.randn(shapes["middle"][0], generator=g) * 0.3],
               "output": [None if i == 2 else torch.randn(s, generator=g) * 0.3 for i, s in enumerate(shapes["input"])]}
    with torch.no_grad():
        out = m(x, t, context=ctx, y=y, control={k: list(v) for k, v in control.items()}, transformer_options={})
    torch.save(dict(config=name, weight_seed=1, weight_checksum=sd_checksum(sd), x=x, t=t, context=ctx, y=y, control=control, out=out),
               os.path.join(GOLD, f"unet_{name}_control.pt"))
    print("unet control", name, "out std", out.std().item())


def gen_chroma(name: str = "tiny_chroma", hw: int = 16, txt_len: int = 64):
    """Reference Chroma transformer (backend/nn/chroma.py) on CPU fp32."""
    from backend.nn.chroma import IntegratedChromaTransformer2DModel
    from oracle import chroma as OC
    cfg = OC.CONFIGS[name]
    sd = OC.random_state_dict(cfg, seed=5)
    m = IntegratedChromaTransformer2DModel(**cfg).eval()
    m.load_state_dict(sd, strict=True)
    g = torch.Generator().manual_seed(6)
    x = torch.randn(2, cfg["in_channels"], hw, hw_w or hw, generator=g)
    ctx = torch.randn(2, txt_len, cfg["context_in_dim"], generator=g)
    t = torch.tensor([0.93, 0.12])
    with torch.no_grad():
        out = m(x, t, ctx)
    torch.save(dict(config=name, weight_seed=5, weight_checksum=sd_checksum(sd), x=x, t=t, context=ctx, out=out),
               os.path.join(GOLD, "chroma_tiny.pt"))
    print("chroma", name, "out std", out.std().item())


def gen_flux(name: str = "tiny_flux", hw: int = 16, txt_len: int = 128, fname: str = "flux_tiny.pt", hw_w: int = None):
    """Reference Flux transformer (backend/nn/flux.py) on CPU fp32, distilled-guidance input included."""
    from backend.nn.flux import IntegratedFluxTransformer2DModel
    from oracle import flux as OF
    cfg = OF.CONFIGS[name]
    sd = OF.random_state_dict(cfg, seed=5)
    m = IntegratedFluxTransformer2DModel(**cfg).eval()
    m.load_state_dict(sd, strict=True)
    g = torch.Generator().manual_seed(6)
    x = torch.randn(2, cfg["in_channels"], hw, hw_w or hw, generator=g)
    ctx = torch.randn(2, txt_len, cfg["context_in_dim"], generator=g)
    y = torch.randn(2, cfg["vec_in_dim"], generator=g)
    t = torch.tensor([0.93, 0.12])
    guidance = torch.tensor([4.0, 4.0])  # 4000 is exact in bf16 (3.5 * 1000 rounds to 3504 in the reference's bf16 run)
    with torch.no_grad():
        out = m(x, t, ctx, y, guidance)
    torch.save(dict(config=name, weight_seed=5, weight_checksum=sd_checksum(sd), x=x, t=t, context=ctx, y=y,
                    guidance=guidance, out=out), os.path.join(GOLD, fname))
    print("flux", name, "out std", out.std().item())


if __name__ == "__main__":
    os.makedirs(GOLD, exist_ok=True)
    ref_import.load()
    which = sys.argv[1:] or ["unet", "traj", "vtraj", "sched", "samplers", "vae", "vae_tiled", "vae_enc", "control", "chroma", "flux", "flux_odd"]
    if "unet" in which:
        gen_unet("tiny_xl")
        gen_unet("tiny_15")
        gen_unet("tiny_15h")
    if "traj" in which:
        gen_trajectories("tiny_xl")
    if "vtraj" in which:
        gen_v_trajectory("tiny_21")
    if "sched" in which:
        gen_schedules()
    if "vae" in which:
        gen_vae("tiny")
    if "vae_tiled" in which:
        gen_vae_tiled("tiny")
    if "samplers" in which:
        gen_samplers()
    if "vae_enc" in which:
        gen_vae_encode("tiny")
    if "control" in which:
        gen_unet_control("tiny_xl")
    if "chroma" in which:
        gen_chroma()
    if "flux" in which:
        gen_flux()                                                  # 64 img + 128 txt tokens: per-stream GEMM launches
        gen_flux(hw=32, txt_len=256, fname="flux_tiny_seg.pt")      # 256 + 256 tokens: two-segment GEMM path
    if "flux_odd" in which:
        gen_flux(hw=15, hw_w=18, txt_len=64, fname="flux_tiny_odd.pt")  # odd height: circular pad to the patch size + crop
