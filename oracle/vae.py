"""ORACLE — test infrastructure only (see oracle/ops.py header).

Functional restatement of the reference VAE decode path: IntegratedAutoencoderKL.decode
(backend/nn/vae.py:305-310) -> Decoder.forward (:248-271) with ResnetBlock (:99-115), AttnBlock (:127-137),
Upsample (:43-57), plus the driver arithmetic of VAE.decode_inner (backend/patcher/vae.py:128-148) and
process_out (backend/nn/vae.py:315-316).  State-dict keys are the reference's parameter names.
"""
from __future__ import annotations

from typing import Dict

import torch

from . import ops as O

SD = Dict[str, torch.Tensor]


def _conv(sd, p, x, padding=1):
    return O.conv2d(x, sd[p + ".weight"], sd[p + ".bias"], padding=padding)


def _gn(sd, p, x):
    return O.group_norm(x, 32, sd[p + ".weight"], sd[p + ".bias"], 1e-6)  # Normalize(), vae.py:12-13


def resnet_block(sd: SD, p: str, x):
    """backend/nn/vae.py:99-115 with temb=None, dropout=0."""
    h = _conv(sd, p + ".conv1", O.silu(_gn(sd, p + ".norm1", x)))
    h = _conv(sd, p + ".conv2", O.silu(_gn(sd, p + ".norm2", h)))
    if (p + ".nin_shortcut.weight") in sd:
        x = _conv(sd, p + ".nin_shortcut", x, padding=0)
    return x + h


def attn_block(sd: SD, p: str, x):
    """backend/nn/vae.py:127-137 + pytorch_attention_single_head_spatial (backend/attention.py:412-427):
    single-head attention over the h*w positions with C-dim queries/keys/values."""
    h_ = _gn(sd, p + ".norm", x)
    q = _conv(sd, p + ".q", h_, padding=0)
    k = _conv(sd, p + ".k", h_, padding=0)
    v = _conv(sd, p + ".v", h_, padding=0)
    b, c, hh, ww = q.shape
    qt, kt, vt = (t.reshape(b, c, hh * ww).transpose(1, 2) for t in (q, k, v))  # [b, L, C]
    out = O.attention(qt, kt, vt, heads=1)
    out = out.transpose(1, 2).reshape(b, c, hh, ww)
    return x + _conv(sd, p + ".proj_out", out, padding=0)


def decoder_structure(cfg: dict):
    boc = list(cfg["block_out_channels"])
    ch = boc[0]
    ch_mult = [c // ch for c in boc]
    nres = len(ch_mult)
    return ch, ch_mult, nres, cfg["layers_per_block"]


def decode(sd: SD, cfg: dict, z: torch.Tensor) -> torch.Tensor:
    """IntegratedAutoencoderKL.decode: post_quant_conv then Decoder.forward.  z is the *processed-out* latent."""
    ch, ch_mult, nres, nrb = decoder_structure(cfg)
    if "post_quant_conv.weight" in sd:
        z = _conv(sd, "post_quant_conv", z, padding=0)
    h = _conv(sd, "decoder.conv_in", z)
    h = resnet_block(sd, "decoder.mid.block_1", h)
    h = attn_block(sd, "decoder.mid.attn_1", h)
    h = resnet_block(sd, "decoder.mid.block_2", h)
    for lvl in reversed(range(nres)):
        for j in range(nrb + 1):
            h = resnet_block(sd, f"decoder.up.{lvl}.block.{j}", h)
        if lvl != 0:
            h = _conv(sd, f"decoder.up.{lvl}.upsample.conv", O.upsample_nearest2x(h))
    h = O.silu(_gn(sd, "decoder.norm_out", h))
    return _conv(sd, "decoder.conv_out", h)


def decode_first_stage(sd: SD, cfg: dict, latent: torch.Tensor) -> torch.Tensor:
    """process_out (vae.py:315-316) -> decode -> clamp((x+1)/2, 0, 1) -> NHWC  (patcher/vae.py:142,147).
    Returns [B, H, W, 3] fp32 in [0, 1]."""
    z = latent / cfg["scaling_factor"] + cfg.get("shift_factor", 0.0)
    x = decode(sd, cfg, z)
    return torch.clamp((x.float() + 1.0) / 2.0, min=0.0, max=1.0).movedim(1, -1)


def tiled_scale(samples, function, tile_y, tile_x, overlap, upscale):
    """backend/patcher/vae.py:11-49 (tiled_scale_multidim, 2-D): per sample, tiles at stride (tile - overlap) clamped to the
    image, each tile's result multiplied by a mask whose first / last `overlap * upscale` rows and columns ramp linearly, the
    masked results and the masks accumulated, output = quotient."""
    out_all = []
    for b in range(samples.shape[0]):
        s = samples[b:b + 1]
        H, W = round(s.shape[2] * upscale), round(s.shape[3] * upscale)
        out = torch.zeros((1, 3, H, W))
        div = torch.zeros((1, 3, H, W))
        for y in range(0, s.shape[2], tile_y - overlap):
            for x in range(0, s.shape[3], tile_x - overlap):
                py = max(0, min(s.shape[2] - overlap, y))
                px = max(0, min(s.shape[3] - overlap, x))
                ly, lx = min(tile_y, s.shape[2] - py), min(tile_x, s.shape[3] - px)
                ps = function(s[:, :, py:py + ly, px:px + lx])
                mask = torch.ones_like(ps)
                feather = round(overlap * upscale)
                for t in range(feather):
                    for d in (2, 3):
                        mask.narrow(d, t, 1).mul_((1.0 / feather) * (t + 1))
                        mask.narrow(d, mask.shape[d] - 1 - t, 1).mul_((1.0 / feather) * (t + 1))
                uy, ux = round(py * upscale), round(px * upscale)
                out[:, :, uy:uy + ps.shape[2], ux:ux + ps.shape[3]] += ps * mask
                div[:, :, uy:uy + ps.shape[2], ux:ux + ps.shape[3]] += mask
        out_all.append(out / div)
    return torch.cat(out_all)


def decode_tiled(sd: SD, cfg: dict, latent: torch.Tensor, tile_x: int = 64, tile_y: int = 64, overlap: int = 16) -> torch.Tensor:
    """VAE.decode_tiled_ (backend/patcher/vae.py:104-115) after process_out: three tilings averaged, tiles of decode + 1,
    clamp(sum / 3 / 2, 0, 1); returns NHWC [B, H, W, 3] like decode_first_stage."""
    z = latent / cfg["scaling_factor"] + cfg.get("shift_factor", 0.0)
    up = 2 ** (len(cfg["block_out_channels"]) - 1)
    fn = lambda a: (decode(sd, cfg, a) + 1.0).float()  # noqa: E731
    out = (tiled_scale(z, fn, tile_y * 2, tile_x // 2, overlap, up) + tiled_scale(z, fn, tile_y // 2, tile_x * 2, overlap, up) +
           tiled_scale(z, fn, tile_y, tile_x, overlap, up))
    return torch.clamp(out / 3.0 / 2.0, min=0.0, max=1.0).movedim(1, -1)


def random_state_dict(cfg: dict, seed: int = 0, dtype=torch.float32) -> SD:
    """Synthetic decoder weights (+ post_quant_conv) with the reference's names; the encoder half is not on
    the txt2img path and is left out."""
    g = torch.Generator().manual_seed(seed)
    sd: SD = {}

    def conv(p, cin, cout, k):
        sd[p + ".weight"] = (torch.randn(cout, cin, k, k, generator=g) * (cin * k * k) ** -0.5).to(dtype)
        sd[p + ".bias"] = (torch.randn(cout, generator=g) * 0.05).to(dtype)

    def norm(p, c):
        sd[p + ".weight"] = (1.0 + 0.1 * torch.randn(c, generator=g)).to(dtype)
        sd[p + ".bias"] = (0.05 * torch.randn(c, generator=g)).to(dtype)

    def res(p, cin, cout):
        norm(p + ".norm1", cin)
        conv(p + ".conv1", cin, cout, 3)
        norm(p + ".norm2", cout)
        conv(p + ".conv2", cout, cout, 3)
        if cin != cout:
            conv(p + ".nin_shortcut", cin, cout, 1)

    ch, ch_mult, nres, nrb = decoder_structure(cfg)
    zc = cfg["latent_channels"]
    if cfg.get("use_post_quant_conv", True):  # backend/nn/vae.py:285
        conv("post_quant_conv", zc, zc, 1)
    block_in = ch * ch_mult[-1]
    conv("decoder.conv_in", zc, block_in, 3)
    res("decoder.mid.block_1", block_in, block_in)
    norm("decoder.mid.attn_1.norm", block_in)
    for n in ("q", "k", "v", "proj_out"):
        conv(f"decoder.mid.attn_1.{n}", block_in, block_in, 1)
    res("decoder.mid.block_2", block_in, block_in)
    for lvl in reversed(range(nres)):
        block_out = ch * ch_mult[lvl]
        for j in range(nrb + 1):
            res(f"decoder.up.{lvl}.block.{j}", block_in, block_out)
            block_in = block_out
        if lvl != 0:
            conv(f"decoder.up.{lvl}.upsample.conv", block_in, block_in, 3)
    norm("decoder.norm_out", block_in)
    conv("decoder.conv_out", block_in, cfg["out_channels"], 3)
    return sd


# ------------------------------------------------------------------------------------------------- encoder
def encode_moments(sd: SD, cfg: dict, x: torch.Tensor) -> torch.Tensor:
    """Encoder.forward (backend/nn/vae.py:180-200) + quant_conv (:293-297): x NCHW in [-1, 1] -> moments [B, 2*zc, h, w].
    Downsample = pad (0,1,0,1) then conv3x3 stride 2 pad 0 (:61-73)."""
    ch, ch_mult, nres, nrb = decoder_structure(cfg)
    h = _conv(sd, "encoder.conv_in", x)
    for lvl in range(nres):
        for j in range(nrb):
            h = resnet_block(sd, f"encoder.down.{lvl}.block.{j}", h)
        if lvl != nres - 1:
            p = f"encoder.down.{lvl}.downsample.conv"
            h = torch.nn.functional.conv2d(torch.nn.functional.pad(h, (0, 1, 0, 1)), sd[p + ".weight"], sd[p + ".bias"], stride=2)
    h = resnet_block(sd, "encoder.mid.block_1", h)
    h = attn_block(sd, "encoder.mid.attn_1", h)
    h = resnet_block(sd, "encoder.mid.block_2", h)
    h = _conv(sd, "encoder.conv_out", O.silu(_gn(sd, "encoder.norm_out", h)))
    if "quant_conv.weight" in sd:
        h = _conv(sd, "quant_conv", h, padding=0)
    return h


def posterior(moments: torch.Tensor, noise=None):
    """DiagonalGaussianDistribution (vae.py:16-32): returns mean + std * noise (noise None -> mode)."""
    mean, logvar = torch.chunk(moments, 2, dim=1)
    logvar = torch.clamp(logvar, -30.0, 20.0)
    return mean if noise is None else mean + torch.exp(0.5 * logvar) * noise


def encode_first_stage(sd: SD, cfg: dict, pixels_nhwc: torch.Tensor, noise=None) -> torch.Tensor:
    """VAE.encode_inner (backend/patcher/vae.py:162-184): pixels NHWC in [0, 1] -> 2x-1 -> encode -> sample, float;
    then process_in (backend/nn/vae.py:312-313) as the diffusion engines apply it (diffusion_engine/sdxl.py:128-132)."""
    x = 2.0 * pixels_nhwc.movedim(-1, 1) - 1.0
    z = posterior(encode_moments(sd, cfg, x), noise).float()
    return (z - cfg.get("shift_factor", 0.0)) * cfg["scaling_factor"]


def random_encoder_state_dict(cfg: dict, seed: int = 0, dtype=torch.float32) -> SD:
    """Synthetic encoder weights (+ quant_conv) with the reference's names (a separate generator so the decoder fixtures
    keep their weights)."""
    g = torch.Generator().manual_seed(seed)
    sd: SD = {}

    def conv(p, cin, cout, k):
        sd[p + ".weight"] = (torch.randn(cout, cin, k, k, generator=g) * (cin * k * k) ** -0.5).to(dtype)
        sd[p + ".bias"] = (torch.randn(cout, generator=g) * 0.05).to(dtype)

    def norm(p, c):
        sd[p + ".weight"] = (1.0 + 0.1 * torch.randn(c, generator=g)).to(dtype)
        sd[p + ".bias"] = (0.05 * torch.randn(c, generator=g)).to(dtype)

    def res(p, cin, cout):
        norm(p + ".norm1", cin)
        conv(p + ".conv1", cin, cout, 3)
        norm(p + ".norm2", cout)
        conv(p + ".conv2", cout, cout, 3)
        if cin != cout:
            conv(p + ".nin_shortcut", cin, cout, 1)

    ch, ch_mult, nres, nrb = decoder_structure(cfg)
    zc = cfg["latent_channels"]
    conv("encoder.conv_in", cfg["in_channels"], ch, 3)
    block_in = ch
    for lvl in range(nres):
        block_out = ch * ch_mult[lvl]
        for j in range(nrb):
            res(f"encoder.down.{lvl}.block.{j}", block_in, block_out)
            block_in = block_out
        if lvl != nres - 1:
            conv(f"encoder.down.{lvl}.downsample.conv", block_in, block_in, 3)
    res("encoder.mid.block_1", block_in, block_in)
    norm("encoder.mid.attn_1.norm", block_in)
    for n in ("q", "k", "v", "proj_out"):
        conv(f"encoder.mid.attn_1.{n}", block_in, block_in, 1)
    res("encoder.mid.block_2", block_in, block_in)
    norm("encoder.norm_out", block_in)
    conv("encoder.conv_out", block_in, 2 * zc, 3)
    conv("quant_conv", 2 * zc, 2 * zc, 1)
    return sd
