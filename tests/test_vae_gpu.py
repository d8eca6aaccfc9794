"""GPU parity tests for the VAE decode path against the reference golden (tiny config,
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
