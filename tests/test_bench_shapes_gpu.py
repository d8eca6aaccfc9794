"""GPU parity at the BENCHMARKED shapes (BASELINE.json configs[1..4]) — the shapes bench.py times, not reduced ones:

  SDXL UNet     [16, 4, 128, 128]  (batch 8 x [uncond | cond], 1024x1024)      fp16 engine vs oracle fp32 on the GPU
  SD1.5 UNet    [16, 4, 64, 64]    (512x512)                                   fp16 engine vs oracle fp32
  VAE decode    [8, 4, 128, 128] -> 8 x 1024x1024x3                            bf16 engine vs oracle fp32 (per image)
  Flux.1-dev    19 + 38 blocks, batch 1, 4096 + 256 tokens                      bf16 engine vs oracle fp32
  Euler loop    SDXL batch 2 x CFG at 128x128, 6 steps                         latent PSNR vs the oracle-port loop in fp16

Stated tolerances: fp16 path per forward rel-RMS <= 3e-3 and max-abs <= 2e-2 of the output's own scale against
the fp32 reference arithmetic on the same fp16-rounded weights; Euler (non-ancestral) final latent PSNR >= 40 dB against the
reference's fp16 arithmetic.  The bf16 paths (VAE, Flux: the reference's own dtypes) are bounded by the distance at which the
reference's own bf16 arithmetic (the oracle port run in bf16 on the same GPU) sits from fp32: ours must not be further
than 1.5x that.  Tiles above row 65 536, all 2048 M-tiles and the 4096-token attention are only exercised here.
"""
import pytest
import torch

from oracle import flux as OF
from oracle import sampling as S
from oracle import unet as OU
from oracle import vae as OV
from tests.util import assert_close, err_stats

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _no_tf32():
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False


def _scaled_max_abs(got, ref):
    """max |got - ref| relative to the reference's RMS (the outputs are O(1) but not exactly unit scale)."""
    got, ref = got.float(), ref.float()
    return ((got - ref).abs().max() / ref.pow(2).mean().sqrt()).item()


@pytest.mark.parametrize("name,hw", [("sdxl", 128), ("sd15", 64)])
def test_unet_benchmark_shape_vs_oracle_fp32(name, hw):
    from b200forge import synthetic
    from b200forge.unet_engine import UNetEngine
    _no_tf32()
    cfg = synthetic.SDXL if name == "sdxl" else synthetic.SD15
    n = 16
    sd = synthetic.random_unet_state_dict(cfg, device=DEV, dtype=torch.float16, seed=0)
    eng = UNetEngine(cfg, sd, dtype=torch.float16, device=DEV)
    g = torch.Generator().manual_seed(33)
    x = torch.randn(n, 4, hw, hw, generator=g).half().to(DEV)
    ctx = torch.randn(n, 77, cfg["context_dim"], generator=g).half().to(DEV)
    y = torch.randn(n, cfg["adm_in_channels"], generator=g).half().to(DEV) if cfg["adm_in_channels"] else None
    t = torch.linspace(999.0, 1.0, n).to(DEV)
    out = eng.forward(x, t, ctx, y)
    out2 = eng.forward(x, t, ctx, y)
    torch.cuda.synchronize()
    assert torch.isfinite(out).all()
    assert torch.equal(out, out2), "the fused forward must be bit-reproducible run to run"
    del eng
    with torch.no_grad():
        ref16 = OU.unet_forward(sd, cfg, x, t, ctx, y)           # the reference's own fp16 arithmetic (ATen / SDPA)
        sd32 = {k: v.float() for k, v in sd.items()}
        ref = torch.empty((n, 4, hw, hw), dtype=torch.float32, device=DEV)
        for i in range(0, n, 4):                                   # fp32 oracle in chunks of 4 samples (memory)
            ref[i:i + 4] = OU.unet_forward(sd32, cfg, x[i:i + 4].float(), t[i:i + 4], ctx[i:i + 4].float(),
                                           None if y is None else y[i:i + 4].float())
    m16, r16 = err_stats(ref16, ref)
    print(f"[parity] oracle-in-fp16 vs oracle fp32 ({name} b16 @{hw}): max_abs={m16:.3e} rel_rms={r16:.3e}")
    sm = _scaled_max_abs(out, ref)
    print(f"[parity] {name} b16 @{hw}: max_abs/ref_rms={sm:.3e}")
    assert_close(f"unet {name} [16,4,{hw},{hw}] fp16 engine vs oracle fp32", out, ref, rel_rms=3e-3)
    assert sm <= 2e-2, f"max-abs {sm:.3e} of the output RMS > 2e-2"
    # per-sample: a wrong tile in one image must not hide in the batch RMS
    for i in range(n):
        _, r = err_stats(out[i], ref[i])
        assert r <= 4e-3, f"sample {i}: rel_rms {r:.3e}"


def test_vae_decode_benchmark_shape_vs_oracle_fp32():
    from b200forge import synthetic
    from b200forge.vae_engine import VAEDecoderEngine
    _no_tf32()
    cfg = synthetic.VAE_SDXL
    sd = synthetic.random_vae_decoder_state_dict(cfg, device=DEV, dtype=torch.bfloat16, seed=1)
    eng = VAEDecoderEngine(cfg, sd, dtype=torch.bfloat16, device=DEV)
    g = torch.Generator().manual_seed(34)
    z = (torch.randn(8, 4, 128, 128, generator=g) * cfg["scaling_factor"]).to(DEV)
    img = eng.decode(z)
    img2 = eng.decode(z)
    torch.cuda.synchronize()
    assert img.shape == (8, 1024, 1024, 3) and torch.isfinite(img).all()
    assert torch.equal(img, img2), "VAE decode must be bit-reproducible run to run"
    del eng
    sd32 = {k: v.float() for k, v in sd.items()}
    worst_ours, worst_ref = 0.0, 0.0
    with torch.no_grad():
        for i in range(8):
            ref = OV.decode_first_stage(sd32, cfg, z[i:i + 1])
            if i < 2:  # the reference's own bf16 arithmetic on the same input (context for the bound)
                ref_bf = OV.decode_first_stage(sd, cfg, z[i:i + 1].bfloat16()).float()
                worst_ref = max(worst_ref, err_stats(ref_bf, ref)[1])
            m, r = err_stats(img[i:i + 1], ref)
            worst_ours = max(worst_ours, r)
            mse = (img[i:i + 1] - ref).pow(2).mean().item()
            psnr = 10 * torch.log10(torch.tensor(1.0 / max(mse, 1e-20))).item()
            print(f"[parity] vae 1024^2 image {i}: max_abs={m:.3e} rel_rms={r:.3e} PSNR={psnr:.1f} dB")
            assert psnr >= 40.0, psnr
    print(f"[parity] vae 1024^2: ours rel_rms {worst_ours:.3e}; oracle-in-bf16 rel_rms {worst_ref:.3e}")
    assert worst_ours <= max(1.5 * worst_ref, 5e-3)


def test_flux_full_depth_vs_oracle_fp32():
    """Flux.1-dev at full depth (19 double + 38 single blocks), batch 1, the benchmark's 128x128 latent
    (4096 image tokens) + 256 text tokens."""
    from b200forge import synthetic
    from b200forge.flux_engine import FluxEngine
    _no_tf32()
    cfg = synthetic.FLUX_DEV
    sd = synthetic.random_flux_state_dict(cfg, device=DEV, dtype=torch.bfloat16, seed=2)
    eng = FluxEngine(cfg, sd, dtype=torch.bfloat16, device=DEV)
    g = torch.Generator().manual_seed(35)
    x = torch.randn(1, 16, 128, 128, generator=g).to(DEV)
    ctx = torch.randn(1, 256, cfg["context_in_dim"], generator=g).bfloat16().to(DEV)
    yv = torch.randn(1, cfg["vec_in_dim"], generator=g).bfloat16().to(DEV)
    t = torch.tensor([0.5], device=DEV)    # 500 and 4000 are exact in bf16 (the reference's bf16 run rounds t*1000 / g*1000)
    gd = torch.tensor([4.0], device=DEV)
    out = eng.forward_nhwc(x, t, ctx, yv, gd).float().permute(0, 3, 1, 2).contiguous()
    torch.cuda.synchronize()
    assert torch.isfinite(out).all()
    del eng
    torch.cuda.empty_cache()
    with torch.no_grad():
        ref_bf = OF.flux_forward(sd, cfg, x.bfloat16(), t, ctx, yv, gd).float()
        sd32 = {k: v.float() for k, v in sd.items()}
        del sd
        ref = OF.flux_forward(sd32, cfg, x, t, ctx.float(), yv.float(), gd)
    m_ref, r_ref = err_stats(ref_bf, ref)
    m, r = err_stats(out, ref)
    print(f"[parity] flux full depth: ours max_abs={m:.3e} rel_rms={r:.3e}; oracle-in-bf16 max_abs={m_ref:.3e} rel_rms={r_ref:.3e}")
    assert r <= max(1.5 * r_ref, 1e-2), (r, r_ref)


def test_sdxl_euler_trajectory_psnr_at_benchmark_latent():
    """6 Euler steps (non-ancestral, CFG 7) of full-width SDXL at the 128x128 latent, batch 2 images (UNet batch 4), through the
    public pipeline with its CUDA graph, against the oracle-port loop in the reference's fp16 arithmetic on the same GPU."""
    from b200forge import synthetic
    from b200forge.pipeline import Txt2ImgPipeline
    cfg = synthetic.SDXL
    sd = synthetic.random_unet_state_dict(cfg, device=DEV, dtype=torch.float16, seed=0)
    pipe = Txt2ImgPipeline(cfg, sd, dtype=torch.float16, device=DEV)
    g = torch.Generator().manual_seed(36)
    B, hw, steps = 2, 128, 6
    cond = dict(crossattn=torch.randn(B, 77, 2048, generator=g), vector=torch.randn(B, 2816, generator=g))
    uncond = dict(crossattn=torch.randn(B, 77, 2048, generator=g), vector=torch.randn(B, 2816, generator=g))
    noise = torch.randn(B, 4, hw, hw, generator=g)
    x = pipe.sample(cond, uncond, noise, steps=steps, sampler="euler", cfg_scale=7.0)
    torch.cuda.synchronize()
    assert torch.isfinite(x).all()
    pred = S.EpsPrediction()
    sig = S.get_sigmas_uniform(pred, steps)
    to16 = lambda d: {k: v.to(DEV).half() for k, v in d.items()}  # noqa: E731
    c16, u16 = to16(cond), to16(uncond)

    def unet16(xc, t, c, yy):
        return OU.unet_forward(sd, cfg, xc.half(), t, c, yy).float()

    den = S.Denoiser(unet16, pred, c16, u16, 7.0, compute_dtype=torch.float16)
    with torch.no_grad():
        ref = S.sample_euler(den, noise.to(DEV) * float(sig[0]), sig.to(DEV))  # sgm_noise_multiplier off (Forge default)
    mse = (x - ref).pow(2).mean()
    psnr = float(10 * torch.log10(ref.abs().max() ** 2 / mse))
    m, r = err_stats(x, ref)
    print(f"[parity] SDXL euler 6 steps @128x128: PSNR={psnr:.1f} dB max_abs={m:.3e} rel_rms={r:.3e}")
    assert psnr >= 40.0, psnr
