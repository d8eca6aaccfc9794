"""GPU parity tests for the VAE decode path and the VAE encode path (§8f rank 1) against the reference golden (tiny config,
made on CPU fp32 by the imported reference) and the oracle in fp32 at SDXL width."""
import os

import pytest
import torch

from oracle import configs as CF
from oracle import vae as OV
from tests.util import assert_close

pytestmark = pytest.mark.gpu
DEV = "cuda"
GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("dtype,tol", [(torch.float16, dict(max_abs=2e-2, rel_rms=4e-3)),
                                       (torch.bfloat16, dict(max_abs=1.5e-1, rel_rms=3e-2))])
def test_vae_decode_vs_reference_golden(dtype, tol):
    from b200forge.vae_engine import VAEDecoderEngine
    g = torch.load(os.path.join(GOLD, "vae_tiny.pt"), weights_only=False)
    cfg = CF.VAE_CONFIGS[g["config"]]
    sd = OV.random_state_dict(cfg, seed=g["weight_seed"])
    eng = VAEDecoderEngine(cfg, sd, dtype=dtype, device=DEV)
    img = eng.decode(g["z"].to(DEV))
    torch.cuda.synchronize()
    ref = torch.clamp((g["out"] + 1.0) / 2.0, 0.0, 1.0).movedim(1, -1)  # patcher/vae.py:142,147
    assert img.shape == ref.shape and img.dtype == torch.float32
    assert_close(f"vae tiny {dtype} vs reference golden", img, ref, **tol)


def test_vae_decode_sdxl_width_vs_oracle_fp32():
    from b200forge.vae_engine import VAEDecoderEngine
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    cfg = CF.VAE_CONFIGS["sdxl"]
    sd = {k: v.bfloat16() for k, v in OV.random_state_dict(cfg, seed=21).items()}
    eng = VAEDecoderEngine(cfg, sd, dtype=torch.bfloat16, device=DEV)
    g = torch.Generator().manual_seed(22)
    z = (torch.randn(2, 4, 32, 32, generator=g) * cfg["scaling_factor"]).to(DEV)
    img = eng.decode(z)
    torch.cuda.synchronize()
    sd32 = {k: v.float().to(DEV) for k, v in sd.items()}
    with torch.no_grad():
        ref = OV.decode_first_stage(sd32, cfg, z)
    assert_close("vae sdxl-width bf16 vs oracle fp32", img, ref, max_abs=1.5e-1, rel_rms=3e-2)


@pytest.mark.parametrize("dtype,tol", [(torch.float16, dict(max_abs=2e-2, rel_rms=5e-3)),
                                       (torch.bfloat16, dict(max_abs=1.5e-1, rel_rms=3e-2))])
def test_vae_encode_vs_reference_golden(dtype, tol):
    """Encoder + quant_conv + DiagonalGaussianDistribution vs the imported reference (tests/golden/vae_enc_tiny.pt):
    moments, the sample with the reference's own noise draw, and process_in."""
    from b200forge.vae_engine import VAEEncoderEngine
    g = torch.load(os.path.join(GOLD, "vae_enc_tiny.pt"), weights_only=False)
    cfg = CF.VAE_CONFIGS[g["config"]]
    sd = OV.random_encoder_state_dict(cfg, seed=g["weight_seed"])
    eng = VAEEncoderEngine(cfg, sd, dtype=dtype, device=DEV)
    px = g["pixels"].to(DEV)
    mom = eng.moments(px)
    torch.cuda.synchronize()
    zc = cfg["latent_channels"]
    assert_close(f"vae encode mean {dtype}", mom[..., :zc].movedim(-1, 1), g["mean"], **tol)
    assert_close(f"vae encode logvar {dtype}", mom[..., zc:2 * zc].movedim(-1, 1).float().clamp(-30, 20), g["logvar"], **tol)
    z = eng.encode(px, g["noise"])
    zmode = eng.encode(px, mode=True)
    lat = eng.encode(px, g["noise"], process_in=True)
    torch.cuda.synchronize()
    assert z.dtype == torch.float32 and z.shape == g["sample"].shape
    assert_close(f"vae encode sample {dtype}", z, g["sample"], **tol)
    assert_close(f"vae encode mode {dtype}", zmode, g["mean"], **tol)
    assert_close(f"vae encode latent (process_in) {dtype}", lat, g["latent"], max_abs=tol["max_abs"], rel_rms=tol["rel_rms"])
    # the default draw is the reference's: torch.randn on the CPU default generator
    torch.manual_seed(1234)
    z2 = eng.encode(px)
    torch.cuda.synchronize()
    # same noise as the golden's; not bit-equal run to run because GroupNorm statistics are accumulated with float atomics
    assert_close(f"vae encode default noise draw {dtype}", z2, z, rel_rms=2e-3 if dtype == torch.float16 else 1.5e-2)


def test_vae_encode_sdxl_width_vs_oracle_fp32():
    from b200forge.vae_engine import VAEEncoderEngine
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    cfg = CF.VAE_CONFIGS["sdxl"]
    sd = {k: v.bfloat16() for k, v in OV.random_encoder_state_dict(cfg, seed=23).items()}
    eng = VAEEncoderEngine(cfg, sd, dtype=torch.bfloat16, device=DEV)
    g = torch.Generator().manual_seed(24)
    px = torch.rand(2, 256, 256, 3, generator=g).to(DEV)
    noise = torch.randn(2, 4, 32, 32, generator=g).to(DEV)
    lat = eng.encode(px, noise, process_in=True)
    torch.cuda.synchronize()
    sd32 = {k: v.float().to(DEV) for k, v in sd.items()}
    with torch.no_grad():
        ref = OV.encode_first_stage(sd32, cfg, px, noise)
    assert_close("vae encode sdxl-width bf16 vs oracle fp32", lat, ref, rel_rms=3e-2)


def test_vae_decode_tiled_vs_reference_golden_and_oracle():
    """Tiled decode (backend/patcher/vae.py:104-115): tiny config against the reference's own tiled_scale run (golden), and the
    SDXL-width decoder at a 96x64 latent with the reference's default 64 / 16 tiling against the oracle restatement in fp32."""
    from b200forge import synthetic
    from b200forge.vae_engine import VAEDecoderEngine
    g = torch.load(os.path.join(GOLD, "vae_tiled_tiny.pt"), weights_only=False)
    cfg = CF.VAE_CONFIGS[g["config"]]
    sd = OV.random_state_dict(cfg, seed=g["weight_seed"])
    eng = VAEDecoderEngine(cfg, sd, dtype=torch.float16, device=DEV)
    img = eng.decode_tiled(g["z"].to(DEV), tile_x=g["tile_x"], tile_y=g["tile_y"], overlap=g["overlap"])
    torch.cuda.synchronize()
    assert_close("tiled vae decode (tiny, fp16) vs reference golden", img, g["out"], max_abs=2e-2, rel_rms=4e-3)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    cfg = synthetic.VAE_SDXL
    sd = synthetic.random_vae_decoder_state_dict(cfg, device=DEV, dtype=torch.bfloat16, seed=1)
    eng = VAEDecoderEngine(cfg, sd, dtype=torch.bfloat16, device=DEV)
    z = (torch.randn(1, 4, 64, 96, generator=torch.Generator().manual_seed(50)) * cfg["scaling_factor"]).to(DEV)
    img = eng.decode_tiled(z)
    torch.cuda.synchronize()
    # the oracle's tile loop accumulates on the CPU; run its decoder calls on the GPU through a device-moving closure
    import oracle.vae as OVm
    sdg = {k: v.float() for k, v in sd.items()}
    up = 8
    fn = lambda a: (OVm.decode(sdg, cfg, a.to(DEV)) + 1.0).float().cpu()  # noqa: E731
    zz = (z / cfg["scaling_factor"]).cpu()
    with torch.no_grad():
        out = (OVm.tiled_scale(zz, fn, 128, 32, 16, up) + OVm.tiled_scale(zz, fn, 32, 128, 16, up) + OVm.tiled_scale(zz, fn, 64, 64, 16, up))
    ref = torch.clamp(out / 3.0 / 2.0, 0.0, 1.0).movedim(1, -1)
    mse = (img.cpu() - ref).pow(2).mean().item()
    psnr = 10 * torch.log10(torch.tensor(1.0 / max(mse, 1e-20))).item()
    print(f"[parity] tiled vae decode SDXL width 768x512: PSNR {psnr:.1f} dB")
    assert psnr >= 40.0, psnr


def test_flux_vae_decode_16_channels_vs_oracle_fp32():
    """The 16-channel Flux / SD3 VAE (scaling 0.3611, shift 0.1159, no post-quant conv) at full width on a 32x32 latent."""
    from b200forge import synthetic
    from b200forge.vae_engine import VAEDecoderEngine
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    cfg = synthetic.VAE_FLUX
    sd = synthetic.random_vae_decoder_state_dict(cfg, device=DEV, dtype=torch.bfloat16, seed=3)
    assert "post_quant_conv.weight" not in sd
    eng = VAEDecoderEngine(cfg, sd, dtype=torch.bfloat16, device=DEV)
    z = ((torch.randn(2, 16, 32, 32, generator=torch.Generator().manual_seed(51)) - cfg["shift_factor"]) * cfg["scaling_factor"]).to(DEV)
    img = eng.decode(z)
    torch.cuda.synchronize()
    with torch.no_grad():
        ref = OV.decode_first_stage({k: v.float() for k, v in sd.items()}, cfg, z)
        ref_bf = OV.decode_first_stage(sd, cfg, z.bfloat16()).float()
    from tests.util import err_stats
    r_ref = err_stats(ref_bf, ref)[1]
    m, r = err_stats(img, ref)
    print(f"[parity] flux vae: ours rel_rms={r:.3e} max_abs={m:.3e}; oracle-in-bf16 rel_rms={r_ref:.3e}")
    assert r <= max(1.5 * r_ref, 5e-3)
