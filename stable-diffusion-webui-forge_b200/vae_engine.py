"""Fused channels-last VAE decoder — the B200 replacement for the decode path
`VAE.decode -> IntegratedAutoencoderKL.decode -> Decoder.forward` (reference backend/patcher/vae.py:128-155,
backend/nn/vae.py:305-310, 248-271) in bf16 (the reference's VAE dtype on sm_80+, memory_management.py:193-199).

Everything is the same kernel set as the UNet: GroupNorm stats/apply+SiLU, tiled implicit-GEMM conv3x3 with the
residual add in its epilogue, GEMMs for the 1x1 convs.  The single-head Dh=C attention of the mid block
(backend/nn/vae.py:118-137) runs as S = Q K^T (GEMM) -> row softmax -> O = S V (GEMM with V^T produced directly by
the V projection with swapped operands), one image at a time.
"""
from __future__ import annotations

from typing import Dict

import torch

from . import ops

SD = Dict[str, torch.Tensor]


class VAEDecoderEngine:
    def __init__(self, cfg: dict, state_dict: SD, dtype: torch.dtype = torch.bfloat16, device="cuda"):
        self.cfg = dict(cfg)
        self.dtype = dtype
        self.device = torch.device(device)
        boc = list(cfg["block_out_channels"])
        self.ch = boc[0]
        self.ch_mult = [c // self.ch for c in boc]
        self.nres = len(boc)
        self.nrb = cfg["layers_per_block"]
        self.zc = cfg["latent_channels"]
        self.scaling = float(cfg["scaling_factor"])
        self.shift = float(cfg.get("shift_factor", 0.0) or 0.0)
        self.w: Dict[str, torch.Tensor] = {}
        if self.device.type == "cuda":
            ops.gn_workspace(self.device)  # created (zeroed) outside any graph capture
        self._pack(state_dict)

    def _t(self, t):
        return t.detach().to(device=self.device, dtype=self.dtype).contiguous()

    def repack(self, state_dict: SD) -> None:
        """Re-pack after the module's parameters changed: same buffers, new contents."""
        old, self.w = self.w, {}
        self._pack(state_dict)
        self.w = ops.refresh_packed(old, self.w)

    def _pack(self, sd: SD) -> None:
        w = self.w
        g = lambda k: self._t(sd[k])  # noqa: E731
        zc = self.zc
        assert zc <= 16, "latent channels: 4 (SD / SDXL) or 16 (Flux, SD3)"
        zp = self.zp = 8 if zc <= 8 else 16  # latent channels padded to a 16-byte row
        # post_quant_conv 1x1 on the padded latent (identity when the VAE has none: Flux / SD3).  process_out is
        # z / scaling + shift (backend/nn/vae.py:315-316): the scale rides on the layout kernel, the shift is folded into this
        # GEMM's bias (W (z/s + shift) + b = W z/s + (W shift + b)), in fp32 before the cast.
        pq = torch.zeros((zp, zp), dtype=torch.float32, device=self.device)
        pqb = torch.zeros((zp,), dtype=torch.float32, device=self.device)
        if "post_quant_conv.weight" in sd:
            pq[:zc, :zc] = sd["post_quant_conv.weight"].detach().to(self.device).float().reshape(zc, zc)
            pqb[:zc] = sd["post_quant_conv.bias"].detach().to(self.device).float()
        else:
            pq[:zc, :zc] = torch.eye(zc, dtype=torch.float32, device=self.device)
        w["pq.b0"] = pqb.to(self.dtype)                       # for latents that already went through process_out
        pqb = pqb.clone()
        pqb[:zc] += pq[:zc, :zc].sum(dim=1) * self.shift
        w["pq.w"], w["pq.b"] = pq.to(self.dtype), pqb.to(self.dtype)
        # conv_in on the padded input via im2col: k = tap*zp + c
        ci = g("decoder.conv_in.weight")  # [Cb, zc, 3, 3]
        cb = ci.shape[0]
        cip = torch.zeros((cb, 3, 3, zp), dtype=self.dtype, device=self.device)
        cip[..., :zc] = ci.permute(0, 2, 3, 1)
        w["conv_in.w"], w["conv_in.b"] = cip.reshape(cb, 9 * zp).contiguous(), g("decoder.conv_in.bias")

        def res(p):
            for n in ("norm1", "norm2"):
                w[f"{p}.{n}.g"], w[f"{p}.{n}.b"] = g(f"{p}.{n}.weight"), g(f"{p}.{n}.bias")
            for n in ("conv1", "conv2"):
                w[f"{p}.{n}.w"], w[f"{p}.{n}.b"] = ops.pack_conv3x3(g(f"{p}.{n}.weight")), g(f"{p}.{n}.bias")
            if f"{p}.nin_shortcut.weight" in sd:
                sw = g(f"{p}.nin_shortcut.weight")
                w[f"{p}.skip.w"], w[f"{p}.skip.b"] = sw.reshape(sw.shape[0], sw.shape[1]).contiguous(), g(f"{p}.nin_shortcut.bias")

        res("decoder.mid.block_1")
        res("decoder.mid.block_2")
        p = "decoder.mid.attn_1"
        w[p + ".norm.g"], w[p + ".norm.b"] = g(p + ".norm.weight"), g(p + ".norm.bias")
        for n in ("q", "k", "v", "proj_out"):
            cw = g(f"{p}.{n}.weight")
            w[f"{p}.{n}.w"], w[f"{p}.{n}.b"] = cw.reshape(cw.shape[0], cw.shape[1]).contiguous(), g(f"{p}.{n}.bias")
        for lvl in range(self.nres):
            for j in range(self.nrb + 1):
                res(f"decoder.up.{lvl}.block.{j}")
            if lvl != 0:
                q = f"decoder.up.{lvl}.upsample.conv"
                uw = g(q + ".weight")
                if uw.shape[1] % 64 == 0 and ops.upconv_folded():  # upsample folded into the convolution (ops.conv3x3_up2x)
                    w[q + ".w4"] = ops.pack_conv3x3_up2x(uw)
                else:
                    w[q + ".w"] = ops.pack_conv3x3(uw)
                w[q + ".b"] = g(q + ".bias")
        w["norm_out.g"], w["norm_out.b"] = g("decoder.norm_out.weight"), g("decoder.norm_out.bias")
        co = ops.pack_conv3x3(g("decoder.conv_out.weight"))  # [3, 9*ch] -> pad to 8 rows
        cop = torch.zeros((8, co.shape[1]), dtype=self.dtype, device=self.device)
        cop[: co.shape[0]] = co
        cob = torch.zeros((8,), dtype=self.dtype, device=self.device)
        cob[: co.shape[0]] = g("decoder.conv_out.bias")
        w["conv_out.w"], w["conv_out.b"] = cop, cob

    # ------------------------------------------------------------------------------------------ blocks
    def _res(self, p: str, x: torch.Tensor) -> torch.Tensor:
        w = self.w
        n, hh, ww, cin = x.shape
        h = ops.groupnorm(x, w[p + ".norm1.g"], w[p + ".norm1.b"], eps=1e-6, silu=True)
        h = ops.conv3x3_any(h, w[p + ".conv1.w"], w[p + ".conv1.b"])
        h = ops.groupnorm(h, w[p + ".norm2.g"], w[p + ".norm2.b"], eps=1e-6, silu=True)
        if (p + ".skip.w") in w:
            cout = w[p + ".skip.w"].shape[0]
            skip = ops.gemm(x.view(-1, cin), w[p + ".skip.w"], w[p + ".skip.b"]).view(n, hh, ww, cout)
        else:
            skip = x
        return ops.conv3x3_any(h, w[p + ".conv2.w"], w[p + ".conv2.b"], residual=skip)

    def _attn(self, p: str, x: torch.Tensor) -> torch.Tensor:
        w = self.w
        n, hh, ww, c = x.shape
        L = hh * ww
        x2d = x.view(n * L, c)
        hn = ops.groupnorm(x, w[p + ".norm.g"], w[p + ".norm.b"], eps=1e-6, silu=False).view(n, L, c)
        q = ops.gemm(hn.view(n * L, c), w[p + ".q.w"], w[p + ".q.b"]).view(n, L, c)
        k = ops.gemm(hn.view(n * L, c), w[p + ".k.w"], w[p + ".k.b"]).view(n, L, c)
        o = torch.empty((n, L, c), dtype=self.dtype, device=self.device)
        if L % 8 == 0:
            s = torch.empty((L, L), dtype=self.dtype, device=self.device)
            vt = torch.empty((c, L), dtype=self.dtype, device=self.device)
            for b in range(n):
                ops.gemm(w[p + ".v.w"], hn[b], w[p + ".v.b"], bias_along_m=True, out=vt)  # V^T [C, L]
                ops.gemm(q[b], k[b], out=s, alpha=c ** -0.5)                              # S = Q K^T / sqrt(C) [L, L]
                ops.softmax_rows_(s, 1.0)
                ops.gemm(s, vt, out=o[b])                                                 # O = P V [L, C]
        else:
            # token counts that are not a multiple of 8 (e.g. the 4x3 edge tiles of the tiled decode): keys padded to L8 rows —
            # the GEMM N / K granularity — and masked by the softmax (valid_cols), so the padded rows contribute exactly 0
            L8 = (L + 7) // 8 * 8
            s = torch.empty((L, L8), dtype=self.dtype, device=self.device)
            vt = torch.empty((c, L8), dtype=self.dtype, device=self.device)
            hn8 = torch.empty((L8, c), dtype=self.dtype, device=self.device)
            k8 = torch.empty((L8, c), dtype=self.dtype, device=self.device)
            ops.zero_(hn8)
            ops.zero_(k8)
            for b in range(n):
                hn8[:L].copy_(hn[b])
                k8[:L].copy_(k[b])
                ops.gemm(w[p + ".v.w"], hn8, w[p + ".v.b"], bias_along_m=True, out=vt)    # V^T [C, L8]
                ops.gemm(q[b], k8, out=s, alpha=c ** -0.5)                                # [L, L8]
                ops.softmax_rows_(s, 1.0, L)
                ops.gemm(s, vt, out=o[b])
        out = ops.gemm(o.view(n * L, c), w[p + ".proj_out.w"], w[p + ".proj_out.b"], residual=x2d)
        return out.view(n, hh, ww, c)

    def supports_latent(self, hh: int, ww: int) -> bool:
        """Every decoder resolution (latent size x 1, 2, 4, ...) must tile on the TMA convolution path."""
        for _ in range(self.nres):
            if not (ops.conv3x3_supported(hh, ww) or ops.any_size_enabled()):
                return False
            hh, ww = hh * 2, ww * 2
        return True

    # ------------------------------------------------------------------------------------------ decode
    @torch.no_grad()
    def decode(self, latent: torch.Tensor, processed_out: bool = False) -> torch.Tensor:
        """latent fp32 NCHW [B, zc, h, w] (sampler output) -> fp32 NHWC [B, 8h, 8w, 3] in [0, 1].
        processed_out: the latent already is z / scaling + shift (what Forge hands VAE.decode, diffusion_engine/sdxl.py:134-138)."""
        return ops.vae_postprocess(self._decode_raw(latent, processed_out))

    @torch.no_grad()
    def decode_tiled(self, latent: torch.Tensor, tile_x: int = 64, tile_y: int = 64, overlap: int = 16,
                     processed_out: bool = False) -> torch.Tensor:
        """VAE.decode_tiled_ (backend/patcher/vae.py:104-115 over tiled_scale_multidim, :11-49) — the reference's low-memory /
        always-tiled decode: three passes with tiles (2*tile_y, tile_x/2), (tile_y/2, 2*tile_x), (tile_y, tile_x), each tile
        decoded on its own (its own GroupNorm statistics, as in the reference), feather-blended over `overlap` latent pixels,
        the three passes averaged: clamp((A + B + C) / 3 / 2, 0, 1) with tiles of (decode + 1).  Same input / output contract
        as `decode`.  Every tile runs the fused decoder; the blending is two small kernels."""
        assert latent.dtype == torch.float32 and latent.dim() == 4
        n, zc, hh, ww = latent.shape
        up = 2 ** (self.nres - 1)
        out = torch.empty((n, hh * up, ww * up, 3), dtype=torch.float32, device=latent.device)
        acc = torch.empty((hh * up, ww * up, 4), dtype=torch.float32, device=latent.device)
        passes = ((tile_y * 2, tile_x // 2), (tile_y // 2, tile_x * 2), (tile_y, tile_x))
        for b in range(n):
            for pi, (ty, tx) in enumerate(passes):
                ops.zero_(acc)
                for y in range(0, hh, ty - overlap):
                    py = max(0, min(hh - overlap, y))
                    ly = min(ty, hh - py)
                    for x in range(0, ww, tx - overlap):
                        px = max(0, min(ww - overlap, x))
                        lx = min(tx, ww - px)
                        raw = self._decode_raw(latent[b:b + 1, :, py:py + ly, px:px + lx].contiguous(), processed_out)
                        ops.tile_blend_(acc, raw, py * up, px * up, feather=overlap * up, bias=1.0)
                ops.tile_resolve_(acc, out[b], accumulate=pi > 0, finalize=pi == len(passes) - 1, final_scale=1.0 / 6.0)
        return out

    def _decode_raw(self, latent: torch.Tensor, processed_out: bool = False) -> torch.Tensor:
        """Decoder output NHWC [B, 8h, 8w, ld] in the VAE dtype, before the (x + 1) / 2 clamp."""
        w = self.w
        assert latent.dtype == torch.float32
        latent = latent.contiguous()
        n, zc, hh, ww = latent.shape
        zp = self.zp
        z = ops.nchw_to_nhwc(latent, self.dtype, ldy=zp, scale=1.0 if processed_out else 1.0 / self.scaling)  # process_out (vae.py:315-316)
        z = ops.gemm(z.view(-1, zp), w["pq.w"], w["pq.b0" if processed_out else "pq.b"]).view(n, hh, ww, zp)
        cols = ops.im2col3x3(z, ldo=9 * zp)
        h = ops.gemm(cols, w["conv_in.w"], w["conv_in.b"]).view(n, hh, ww, -1)
        h = self._res("decoder.mid.block_1", h)
        h = self._attn("decoder.mid.attn_1", h)
        h = self._res("decoder.mid.block_2", h)
        for lvl in reversed(range(self.nres)):
            for j in range(self.nrb + 1):
                h = self._res(f"decoder.up.{lvl}.block.{j}", h)
            if lvl != 0:
                q = f"decoder.up.{lvl}.upsample.conv"
                if q + ".w4" in w:
                    h = ops.conv3x3_up2x(h, w[q + ".w4"], w[q + ".b"])
                else:
                    h = ops.conv3x3_any(ops.upsample2x(h), w[q + ".w"], w[q + ".b"])
        h = ops.groupnorm(h, w["norm_out.g"], w["norm_out.b"], eps=1e-6, silu=True)
        return ops.conv3x3_any(h, w["conv_out.w"], w["conv_out.b"])


class VAEEncoderEngine:
    """Fused channels-last VAE encoder — `VAE.encode -> IntegratedAutoencoderKL.encode -> Encoder.forward ->
    DiagonalGaussianDistribution.sample` (reference backend/patcher/vae.py:162-191, backend/nn/vae.py:16-32, 140-200,
    293-303) on the same kernels as the decoder: GroupNorm stats/apply+SiLU, implicit-GEMM conv3x3 with the residual in
    its epilogue, 1x1 convs as GEMMs, the stride-2 downsample (pad (0,1,0,1), 3x3, stride 2) as im2col + GEMM, the
    single-head mid attention as GEMM -> row softmax -> GEMM.  img2img / hires-fix entry."""

    def __init__(self, cfg: dict, state_dict: SD, dtype: torch.dtype = torch.bfloat16, device="cuda"):
        self.cfg = dict(cfg)
        self.dtype = dtype
        self.device = torch.device(device)
        boc = list(cfg["block_out_channels"])
        self.ch = boc[0]
        self.ch_mult = [c // self.ch for c in boc]
        self.nres = len(boc)
        self.nrb = cfg["layers_per_block"]
        self.zc = cfg["latent_channels"]
        self.scaling = float(cfg["scaling_factor"])
        self.shift = float(cfg.get("shift_factor", 0.0) or 0.0)
        self.w: Dict[str, torch.Tensor] = {}
        if self.device.type == "cuda":
            ops.gn_workspace(self.device)  # created (zeroed) outside any graph capture
        self._pack(state_dict)

    _t = VAEDecoderEngine._t
    _res = VAEDecoderEngine._res
    _attn = VAEDecoderEngine._attn

    def repack(self, state_dict: SD) -> None:
        """Re-pack after the module's parameters changed: same buffers, new contents."""
        old, self.w = self.w, {}
        self._pack(state_dict)
        self.w = ops.refresh_packed(old, self.w)

    def _pack(self, sd: SD) -> None:
        w = self.w
        g = lambda k: self._t(sd[k])  # noqa: E731
        zc = self.zc
        assert 2 * zc <= 8 and cfg_in(self.cfg) == 3
        ci = g("encoder.conv_in.weight")  # [ch, 3, 3, 3] -> im2col on the 8-channel padded input: k = tap*8 + c
        cip = torch.zeros((ci.shape[0], 3, 3, 8), dtype=self.dtype, device=self.device)
        cip[..., :3] = ci.permute(0, 2, 3, 1)
        w["conv_in.w"], w["conv_in.b"] = cip.reshape(ci.shape[0], 72).contiguous(), g("encoder.conv_in.bias")

        def res(p):
            for n in ("norm1", "norm2"):
                w[f"{p}.{n}.g"], w[f"{p}.{n}.b"] = g(f"{p}.{n}.weight"), g(f"{p}.{n}.bias")
            for n in ("conv1", "conv2"):
                w[f"{p}.{n}.w"], w[f"{p}.{n}.b"] = ops.pack_conv3x3(g(f"{p}.{n}.weight")), g(f"{p}.{n}.bias")
            if f"{p}.nin_shortcut.weight" in sd:
                sw = g(f"{p}.nin_shortcut.weight")
                w[f"{p}.skip.w"], w[f"{p}.skip.b"] = sw.reshape(sw.shape[0], sw.shape[1]).contiguous(), g(f"{p}.nin_shortcut.bias")

        for lvl in range(self.nres):
            for j in range(self.nrb):
                res(f"encoder.down.{lvl}.block.{j}")
            if lvl != self.nres - 1:
                q = f"encoder.down.{lvl}.downsample.conv"
                w[q + ".w"], w[q + ".b"] = ops.pack_conv3x3(g(q + ".weight")), g(q + ".bias")
        res("encoder.mid.block_1")
        res("encoder.mid.block_2")
        p = "encoder.mid.attn_1"
        w[p + ".norm.g"], w[p + ".norm.b"] = g(p + ".norm.weight"), g(p + ".norm.bias")
        for n in ("q", "k", "v", "proj_out"):
            cw = g(f"{p}.{n}.weight")
            w[f"{p}.{n}.w"], w[f"{p}.{n}.b"] = cw.reshape(cw.shape[0], cw.shape[1]).contiguous(), g(f"{p}.{n}.bias")
        w["norm_out.g"], w["norm_out.b"] = g("encoder.norm_out.weight"), g("encoder.norm_out.bias")
        co = ops.pack_conv3x3(g("encoder.conv_out.weight"))  # [2zc, 9*C]
        cop = torch.zeros((8, co.shape[1]), dtype=self.dtype, device=self.device)
        cop[: co.shape[0]] = co
        cob = torch.zeros((8,), dtype=self.dtype, device=self.device)
        cob[: co.shape[0]] = g("encoder.conv_out.bias")
        w["conv_out.w"], w["conv_out.b"] = cop, cob
        qw = torch.zeros((8, 8), dtype=self.dtype, device=self.device)
        qb = torch.zeros((8,), dtype=self.dtype, device=self.device)
        if "quant_conv.weight" in sd:
            qw[: 2 * zc, : 2 * zc] = g("quant_conv.weight").reshape(2 * zc, 2 * zc)
            qb[: 2 * zc] = g("quant_conv.bias")
        else:
            qw[: 2 * zc, : 2 * zc] = torch.eye(2 * zc, dtype=self.dtype, device=self.device)
        w["quant.w"], w["quant.b"] = qw, qb

    def supports_image(self, H: int, W: int) -> bool:
        """Every encoder resolution (image size / 1, 2, 4, ...) must tile on the TMA convolution path."""
        for lvl in range(self.nres):
            if not (ops.conv3x3_supported(H, W) or ops.any_size_enabled()):
                return False
            if lvl != self.nres - 1:
                if H % 2 or W % 2:
                    return False
                H, W = H // 2, W // 2
        return True

    @torch.no_grad()
    def moments(self, pixels: torch.Tensor) -> torch.Tensor:
        """pixels fp32 NHWC [B, H, W, 3] in [0, 1] -> moments NHWC [B, H/8.., W/8.., 8] (mean | logvar) in the VAE dtype."""
        w = self.w
        n, H, W, _ = pixels.shape
        x = ops.vae_preprocess(pixels.contiguous(), self.dtype)                      # 2x - 1, 8-channel padded
        h = ops.gemm(ops.im2col3x3(x, ldo=72), w["conv_in.w"], w["conv_in.b"]).view(n, H, W, -1)
        for lvl in range(self.nres):
            for j in range(self.nrb):
                h = self._res(f"encoder.down.{lvl}.block.{j}", h)
            if lvl != self.nres - 1:
                q = f"encoder.down.{lvl}.downsample.conv"
                nn_, hh, ww, c = h.shape
                cols = ops.im2col3x3(h, stride=2, pad_lo=0, pad_hi=1)              # F.pad(x, (0,1,0,1)) + stride-2 conv
                h = ops.gemm(cols, w[q + ".w"], w[q + ".b"]).view(nn_, hh // 2, ww // 2, c)
        h = self._res("encoder.mid.block_1", h)
        h = self._attn("encoder.mid.attn_1", h)
        h = self._res("encoder.mid.block_2", h)
        h = ops.groupnorm(h, w["norm_out.g"], w["norm_out.b"], eps=1e-6, silu=True)
        h = ops.conv3x3_any(h, w["conv_out.w"], w["conv_out.b"])                          # [B, h, w, 8]
        n2, hh, ww, _ = h.shape
        return ops.gemm(h.view(-1, 8), w["quant.w"], w["quant.b"]).view(n2, hh, ww, 8)

    @torch.no_grad()
    def encode(self, pixels: torch.Tensor, noise: torch.Tensor = None, *, mode: bool = False,
               process_in: bool = False) -> torch.Tensor:
        """VAE.encode: latent NCHW fp32 [B, zc, h, w] = mean + std * noise.  `noise` [B, zc, h, w] fp32; None draws it
        the way the reference does (torch.randn on the CPU default generator, vae.py:28) unless `mode`.
        process_in=True also applies (z - shift) * scaling_factor (backend/nn/vae.py:312-313)."""
        mom = self.moments(pixels)
        n, hh, ww, _ = mom.shape
        if noise is None and not mode:
            noise = torch.randn((n, self.zc, hh, ww)).to(self.device)
        if noise is not None:
            noise = noise.to(device=self.device, dtype=torch.float32).contiguous()
        assert not process_in or self.shift == 0.0
        return ops.vae_posterior(mom, self.zc, noise, scale=self.scaling if process_in else 1.0)


def cfg_in(cfg: dict) -> int:
    return int(cfg.get("in_channels", 3))
