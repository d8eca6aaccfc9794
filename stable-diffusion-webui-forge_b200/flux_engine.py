"""Flux (DiT) transformer forward on the sm_100a kernels — the replacement for
`IntegratedFluxTransformer2DModel.forward` (backend/nn/flux.py:326-422) behind plug point P3.

Layout.  One joint token activation `xj` [B, Lt + Li, hidden] holds every sample's txt rows followed by its img
rows for the whole forward — the order the reference concatenates them in for attention (flux.py:238-243) and for
the single-stream blocks (flux.py:370) — so no concat / split copies exist.  Double-stream blocks run ONE GEMM per
projection over all rows with two weight sets selected per 256-row tile (`ops.gemm(seg=...)`); when Lt or Li is
not a multiple of 256 the same projections run per sample and per stream on row views.

Per block:  adaLN (LayerNorm + per-sample shift/scale, one pass)  ->  QKV GEMM (bias)  ->  RMSNorm(q,k) + RoPE in
place  ->  flash attention reading q/k/v as column slices of the QKV buffer  ->  projection GEMM whose epilogue
applies the modulation gate and the residual add in place.  MLP: GEMM + tanh-GELU epilogue, GEMM + gate + residual.
Single-stream blocks: linear1 = [qkv | mlp] in one GEMM with GELU on the mlp columns only; linear2 reads
[attention | gelu(mlp)] as a two-source A operand (no concat).
All 19*12 + 38*3 + 2 modulation vectors come from one stacked GEMM on silu(vec) per forward (they depend only on
the timestep / guidance / pooled-text vector), instead of one small Linear per block per stream.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch

from . import ops
from .lib import EPI_GELU_TANH, EPI_NONE, EPI_SILU


def rope_tables(h_len: int, w_len: int, txt_len: int, axes_dim, theta: float, device) -> tuple:
    """EmbedND over ids = [txt zeros | (0, row, col)] (flux.py:21-42, 75-89, 402-409): fp64 angles -> fp32 cos/sin
    tables [txt_len + h_len*w_len, sum(axes_dim)/2]."""
    ids = torch.zeros(h_len, w_len, 3, dtype=torch.float64)
    ids[..., 1] += torch.arange(h_len, dtype=torch.float64)[:, None]
    ids[..., 2] += torch.arange(w_len, dtype=torch.float64)[None, :]
    ids = torch.cat([torch.zeros(txt_len, 3, dtype=torch.float64), ids.reshape(-1, 3)], dim=0)
    cs, sn = [], []
    for i, d in enumerate(axes_dim):
        omega = 1.0 / (theta ** (torch.arange(0, d, 2, dtype=torch.float64) / d))
        ang = ids[:, i:i + 1] * omega[None, :]
        cs.append(torch.cos(ang).float())
        sn.append(torch.sin(ang).float())
    return torch.cat(cs, 1).contiguous().to(device), torch.cat(sn, 1).contiguous().to(device)


class FluxEngine:
    def __init__(self, cfg: dict, sd: Dict[str, torch.Tensor], dtype=torch.bfloat16, device="cuda"):
        self.cfg = dict(cfg)
        self.dtype, self.device = dtype, torch.device(device)
        self.hidden = hs = cfg["hidden_size"]
        self.heads = cfg["num_heads"]
        if hs // self.heads != 128:
            raise ops.B200Error(-2, f"Flux head dim {hs // self.heads} not supported by the fused path (128 only)")
        self.mlp = int(hs * cfg["mlp_ratio"])
        self.depth, self.depth_single = cfg["depth"], cfg["depth_single_blocks"]
        self.in_ch = cfg["in_channels"]
        self.guidance_embed = cfg["guidance_embed"]
        self.w = self._pack(sd)
        self._plans: dict = {}

    # ------------------------------------------------------------------ weights
    def repack(self, sd) -> None:
        """Re-pack after the module's parameters changed (Forge merged or removed a LoRA): same buffers, new contents."""
        self.w = ops.refresh_packed(self.w, self._pack(sd))

    def _pack(self, sd):
        dt, dev = self.dtype, self.device

        def t(k):
            return sd[k].to(device=dev, dtype=dt).contiguous()

        w = {}
        for k in sd:
            if ".mod.lin." in k or ".modulation.lin." in k or "adaLN_modulation" in k:
                continue
            w[k] = t(k)
        # every Modulation / adaLN linear stacked into one [sum, hidden] weight (flux.py:169-178, 315)
        names = []
        for i in range(self.depth):
            names += [f"double_blocks.{i}.img_mod.lin", f"double_blocks.{i}.txt_mod.lin"]
        names += [f"single_blocks.{i}.modulation.lin" for i in range(self.depth_single)]
        names += ["final_layer.adaLN_modulation.1"]
        self.mod_off = {}
        off = 0
        for n in names:
            self.mod_off[n] = off
            off += sd[n + ".weight"].shape[0]
        w["mod.weight"] = torch.cat([sd[n + ".weight"].to(device=dev, dtype=dt) for n in names], 0).contiguous()
        w["mod.bias"] = torch.cat([sd[n + ".bias"].to(device=dev, dtype=dt) for n in names], 0).contiguous()
        self.mod_total = off
        return w

    # ------------------------------------------------------------------ per-shape buffers
    def _plan(self, B: int, H: int, W: int, Lt: int):
        key = (B, H, W, Lt)
        p = self._plans.get(key)
        if p is not None:
            return p
        hs, dt, dev = self.hidden, self.dtype, self.device
        Li = (H // 2) * (W // 2)
        L = Lt + Li
        rows = B * L

        def buf(*shape, dtype=dt):
            return torch.empty(shape, dtype=dtype, device=dev)

        p = dict(Li=Li, L=L, rows=rows, grouped=(Lt % 256 == 0 and Li % 256 == 0),
                 patches=buf(B * Li, 4 * self.in_ch), temb=buf(B, 256), gemb=buf(B, 256), h1=buf(B, hs), vec=buf(B, hs),
                 svec=buf(B, hs), mod=buf(B, self.mod_total), xj=buf(rows, hs), xm=buf(rows, hs), qkv=buf(rows, 3 * hs),
                 attn=buf(rows, hs), hmlp=buf(rows, self.mlp), y1=buf(rows, 3 * hs + self.mlp),
                 tok_out=buf(B * Li, 4 * self.in_ch), t1000=buf(B, dtype=torch.float32), g1000=buf(B, dtype=torch.float32))
        p["cos"], p["sin"] = rope_tables(H // 2, W // 2, Lt, self.cfg["axes_dim"], self.cfg["theta"], dev)
        self._plans[key] = p
        return p

    # ------------------------------------------------------------------ helpers
    def _mod(self, p, name: str, k: int) -> torch.Tensor:
        """k-th hidden-wide chunk of a Modulation output: a [B, hidden] view of the stacked modulation buffer."""
        o = self.mod_off[name] + k * self.hidden
        return p["mod"][:, o:o + self.hidden]

    def _seg_gemm(self, p, a, wt, wi, bt, bi, out, *, B, Lt, epilogue=EPI_NONE, gate_t=None, gate_i=None, residual=None):
        """out = [residual +] [gate *] epi(a @ W_seg^T + b_seg) with txt rows using (wt, bt, gate_t), img rows (wi, bi, gate_i)."""
        L = p["L"]
        if p["grouped"]:
            return ops.gemm(a, wt, bt, epilogue=epilogue, rowvec=gate_t, rows_per_vec=L, rowvec_mul=gate_t is not None,
                            residual=residual, out=out, seg=(L, Lt, wi, bi, gate_i))
        for b in range(B):
            for (lo, hi, w_, b_, g_) in ((b * L, b * L + Lt, wt, bt, gate_t), (b * L + Lt, (b + 1) * L, wi, bi, gate_i)):
                ops.gemm(a[lo:hi], w_, b_, epilogue=epilogue, rowvec=None if g_ is None else g_[b:b + 1],
                         rows_per_vec=hi - lo, rowvec_mul=g_ is not None,
                         residual=None if residual is None else residual[lo:hi], out=out[lo:hi])
        return out

    def _attention(self, p, src, B, out):
        hs, L = self.hidden, p["L"]
        v3 = src.view(B, L, src.shape[1])
        ops.attention(v3[:, :, 0:hs], v3[:, :, hs:2 * hs], v3[:, :, 2 * hs:3 * hs], self.heads, out=out.view(B, L, hs))
        return out

    # ------------------------------------------------------------------ modulation vectors
    def _compute_modulation(self, p, timestep, y, guidance) -> None:
        """Fills p["mod"] [B, mod_total]: every block's shift / scale / gate vectors for this forward."""
        w, hs = self.w, self.hidden
        # vec = time_in(temb(1000 t)) [+ guidance_in(temb(1000 g))] + vector_in(y)      (flux.py:355-361, 52-72)
        torch.mul(timestep, 1000.0, out=p["t1000"])
        ops.timestep_embedding(p["t1000"], 256, self.dtype, out=p["temb"])
        ops.gemm(p["temb"], w["time_in.in_layer.weight"], w["time_in.in_layer.bias"], epilogue=EPI_SILU, out=p["h1"])
        ops.gemm(p["h1"], w["time_in.out_layer.weight"], w["time_in.out_layer.bias"], out=p["vec"])
        if self.guidance_embed:
            if guidance is None:
                raise ValueError("Didn't get guidance strength for guidance distilled model.")
            # KModel casts every extra cond to the computation dtype (k_model.py:37-42) and flux.py:53 scales it there:
            # in bf16, 3.5 * 1000 is 3504 — reproduced so the embedding matches the reference's bf16 run
            p["g1000"].copy_(guidance.to(self.dtype) * 1000.0)
            ops.timestep_embedding(p["g1000"], 256, self.dtype, out=p["gemb"])
            ops.gemm(p["gemb"], w["guidance_in.in_layer.weight"], w["guidance_in.in_layer.bias"], epilogue=EPI_SILU, out=p["h1"])
            ops.gemm(p["h1"], w["guidance_in.out_layer.weight"], w["guidance_in.out_layer.bias"], residual=p["vec"], out=p["vec"])
        ops.gemm(y, w["vector_in.in_layer.weight"], w["vector_in.in_layer.bias"], epilogue=EPI_SILU, out=p["h1"])
        ops.gemm(p["h1"], w["vector_in.out_layer.weight"], w["vector_in.out_layer.bias"], residual=p["vec"], out=p["vec"])
        ops.silu(p["vec"], out=p["svec"])
        ops.gemm(p["svec"], w["mod.weight"], w["mod.bias"], out=p["mod"])

    # ------------------------------------------------------------------ forward
    def forward_tokens(self, x: torch.Tensor, timestep: torch.Tensor, context: torch.Tensor, y: torch.Tensor,
                       guidance: Optional[torch.Tensor]) -> tuple:
        """Runs the transformer; returns (token output [B*Li, 4*C], plan).  x NCHW fp32 or dtype; timestep / guidance
        fp32 [B] (sigma in [0, 1] and the distilled guidance scale, k_model.py:25-46); context [B, Lt, ctx]; y [B, vec]."""
        w, hs, Hh = self.w, self.hidden, self.heads
        B, C, H, W = x.shape
        Lt = context.shape[1]
        if (H | W) & 1:
            raise ops.B200Error(-2, "odd latent size (circular padding branch, flux.py:396-397) is not on the fused path")
        p = self._plan(B, H, W, Lt)
        L, rows = p["L"], p["rows"]
        ops.flux_patchify(x, self.dtype, out=p["patches"])
        self._compute_modulation(p, timestep, y, guidance)
        # token streams into the joint activation
        xj = p["xj"]
        ctx2d = context.reshape(B * Lt, context.shape[2])
        for b in range(B):
            ops.gemm(ctx2d[b * Lt:(b + 1) * Lt], w["txt_in.weight"], w["txt_in.bias"], out=xj[b * L:b * L + Lt])
            ops.gemm(p["patches"][b * p["Li"]:(b + 1) * p["Li"]], w["img_in.weight"], w["img_in.bias"], out=xj[b * L + Lt:(b + 1) * L])
        cos, sin = p["cos"], p["sin"]
        for i in range(self.depth):
            q = f"double_blocks.{i}"
            im, tm = q + ".img_mod.lin", q + ".txt_mod.lin"
            ops.adaln(xj, self._mod(p, tm, 0), self._mod(p, tm, 1), shift1=self._mod(p, im, 0), scale1=self._mod(p, im, 1),
                      seg_period=L, seg_split=Lt, out=p["xm"])
            self._seg_gemm(p, p["xm"], w[q + ".txt_attn.qkv.weight"], w[q + ".img_attn.qkv.weight"],
                           w.get(q + ".txt_attn.qkv.bias"), w.get(q + ".img_attn.qkv.bias"), p["qkv"], B=B, Lt=Lt)
            ops.qk_norm_rope_(p["qkv"], Hh, w[q + ".txt_attn.norm.query_norm.scale"], w[q + ".txt_attn.norm.key_norm.scale"], cos, sin,
                              q_scale1=w[q + ".img_attn.norm.query_norm.scale"], k_scale1=w[q + ".img_attn.norm.key_norm.scale"],
                              seg_split=Lt)
            self._attention(p, p["qkv"], B, p["attn"])
            self._seg_gemm(p, p["attn"], w[q + ".txt_attn.proj.weight"], w[q + ".img_attn.proj.weight"],
                           w[q + ".txt_attn.proj.bias"], w[q + ".img_attn.proj.bias"], xj, B=B, Lt=Lt,
                           gate_t=self._mod(p, tm, 2), gate_i=self._mod(p, im, 2), residual=xj)
            ops.adaln(xj, self._mod(p, tm, 3), self._mod(p, tm, 4), shift1=self._mod(p, im, 3), scale1=self._mod(p, im, 4),
                      seg_period=L, seg_split=Lt, out=p["xm"])
            self._seg_gemm(p, p["xm"], w[q + ".txt_mlp.0.weight"], w[q + ".img_mlp.0.weight"], w[q + ".txt_mlp.0.bias"],
                           w[q + ".img_mlp.0.bias"], p["hmlp"], B=B, Lt=Lt, epilogue=EPI_GELU_TANH)
            self._seg_gemm(p, p["hmlp"], w[q + ".txt_mlp.2.weight"], w[q + ".img_mlp.2.weight"], w[q + ".txt_mlp.2.bias"],
                           w[q + ".img_mlp.2.bias"], xj, B=B, Lt=Lt, gate_t=self._mod(p, tm, 5), gate_i=self._mod(p, im, 5),
                           residual=xj)
        y1 = p["y1"]
        for i in range(self.depth_single):
            q = f"single_blocks.{i}"
            m = q + ".modulation.lin"
            ops.adaln(xj, self._mod(p, m, 0), self._mod(p, m, 1), seg_period=L, seg_split=L, out=p["xm"])
            ops.gemm(p["xm"], w[q + ".linear1.weight"], w[q + ".linear1.bias"], epilogue=EPI_GELU_TANH, act_col0=3 * hs, out=y1)
            ops.qk_norm_rope_(y1, Hh, w[q + ".norm.query_norm.scale"], w[q + ".norm.key_norm.scale"], cos, sin)
            self._attention(p, y1, B, p["attn"])
            ops.gemm(p["attn"], w[q + ".linear2.weight"], w[q + ".linear2.bias"], a2=y1[:, 3 * hs:], rowvec=self._mod(p, m, 2),
                     rows_per_vec=L, rowvec_mul=True, residual=xj, out=xj)
        f = "final_layer.adaLN_modulation.1"
        ops.adaln(xj, self._mod(p, f, 0), self._mod(p, f, 1), seg_period=L, seg_split=L, out=p["xm"])  # chunk order: shift, scale (flux.py:318)
        for b in range(B):
            ops.gemm(p["xm"][b * L + Lt:(b + 1) * L], w["final_layer.linear.weight"], w["final_layer.linear.bias"],
                     out=p["tok_out"][b * p["Li"]:(b + 1) * p["Li"]])
        return p["tok_out"], p

    def forward(self, x, timestep, context, y, guidance=None) -> torch.Tensor:
        """Reference signature and output: NCHW fp32 [B, C, H, W] (flux.py:389-422; KModel casts to float, k_model.py:44).
        Odd latent sizes take the reference's route: circular padding to the patch size on the way in (:394-397, a tensor copy
        of the 16-channel latent), crop on the way out (:412)."""
        B, C, H, W = x.shape
        if (H | W) & 1:
            xp = torch.nn.functional.pad(x, (0, W & 1, 0, H & 1), mode="circular")
            return self.forward(xp, timestep, context, y, guidance)[:, :, :H, :W].contiguous()
        tok, _ = self.forward_tokens(x.contiguous(), timestep.float().contiguous(), context.to(self.dtype).contiguous(),
                                     y.to(self.dtype).contiguous(), None if guidance is None else guidance.float().contiguous())
        return ops.flux_unpatchify(tok, B, C, H, W, nchw_f32=True)

    def forward_nhwc(self, x, timestep, context, y, guidance=None, out=None) -> torch.Tensor:
        """Same forward, output as channels-last [B, H, W, C] in the compute dtype — what the fused sampler step reads."""
        B, C, H, W = x.shape
        tok, _ = self.forward_tokens(x, timestep, context, y, guidance)
        return ops.flux_unpatchify(tok, B, C, H, W, nchw_f32=False, out=out)


class ChromaEngine(FluxEngine):
    """Chroma (backend/nn/chroma.py:138-307): Flux's blocks with the modulation vectors of all blocks produced by one
    Approximator MLP (chroma.py:14-28) from [timestep embedding | zero-guidance embedding | modulation-index embedding]
    — no time / vector / guidance embedders, no per-block Modulation linears, no pooled-text input.  Only the source of
    p["mod"] and the vector order (distribute_modulations, chroma.py:181-243) differ from FluxEngine; the block launch
    sequence is inherited.

    Status: the launch sequence is pinned on the CPU against the imported reference (tests/test_engines_emulated.py);
    `b200_rmsnorm_rows`, the one kernel this engine adds, has not run on hardware yet."""

    def __init__(self, cfg: dict, sd: Dict[str, torch.Tensor], dtype=torch.bfloat16, device="cuda"):
        self.g_layers = cfg["guidance_n_layers"]
        self.g_hidden = cfg["guidance_hidden_dim"]
        if cfg["guidance_out_dim"] != cfg["hidden_size"]:
            raise ValueError("Chroma: guidance_out_dim must equal hidden_size (the vectors modulate hidden-wide rows)")
        super().__init__(dict(cfg, guidance_embed=False), sd, dtype=dtype, device=device)

    def _pack(self, sd):
        dt, dev = self.dtype, self.device
        w = {k: v.to(device=dev, dtype=dt).contiguous() for k, v in sd.items()}
        # vector order of distribute_modulations: single blocks (3 each), img_mod of every double block (6 each), txt_mod
        # of every double block (6 each), final layer (2: shift, scale)
        self.mod_off, idx = {}, 0
        for i in range(self.depth_single):
            self.mod_off[f"single_blocks.{i}.modulation.lin"] = idx * self.hidden
            idx += 3
        for s in ("img", "txt"):
            for i in range(self.depth):
                self.mod_off[f"double_blocks.{i}.{s}_mod.lin"] = idx * self.hidden
                idx += 6
        self.mod_off["final_layer.adaLN_modulation.1"] = idx * self.hidden
        self.n_vec = idx + 2
        self.mod_total = self.n_vec * self.hidden
        # modulation-index embedding: timestep_embedding(arange(n_vec), 32) (chroma.py:257), a constant of the model
        half = 16
        freqs = torch.exp(-torch.log(torch.tensor(10000.0)) * torch.arange(half, dtype=torch.float32) / half)
        args = (1000.0 * torch.arange(self.n_vec, dtype=torch.float32))[:, None] * freqs[None]
        self.idx_emb = torch.cat([torch.cos(args), torch.sin(args)], -1).to(device=dev, dtype=dt)
        return w

    def _plan(self, B, H, W, Lt):
        p = super()._plan(B, H, W, Lt)
        if "g_in" not in p:
            dt, dev, rows = self.dtype, self.device, B * self.n_vec
            p["g_in"] = torch.zeros((rows, 64), dtype=dt, device=dev)
            p["g_in"].view(B, self.n_vec, 64)[:, :, 32:] = self.idx_emb                      # modulation index part
            p["g_in"].view(B, self.n_vec, 64)[:, :, 16:24] = 1.0                             # timestep_embedding(0, 16) = [cos 0 | sin 0]
            p["t16"] = torch.empty((B, 16), dtype=dt, device=dev)
            p["g_x"] = torch.empty((rows, self.g_hidden), dtype=dt, device=dev)
            p["g_n"] = torch.empty((rows, self.g_hidden), dtype=dt, device=dev)
            p["g_h"] = torch.empty((rows, self.g_hidden), dtype=dt, device=dev)
        return p

    def _compute_modulation(self, p, timestep, y, guidance) -> None:
        w = self.w
        B = timestep.shape[0]
        torch.mul(timestep, 1000.0, out=p["t1000"])
        ops.timestep_embedding(p["t1000"], 16, self.dtype, out=p["t16"])                     # chroma.py:255
        p["g_in"].view(B, self.n_vec, 64)[:, :, :16] = p["t16"][:, None, :]
        q = "distilled_guidance_layer"
        ops.gemm(p["g_in"], w[q + ".in_proj.weight"], w[q + ".in_proj.bias"], out=p["g_x"])
        for i in range(self.g_layers):                                                       # x = x + MLP(RMSNorm(x))
            ops.rmsnorm_rows(p["g_x"], w[f"{q}.norms.{i}.scale"], out=p["g_n"])
            ops.gemm(p["g_n"], w[f"{q}.layers.{i}.in_layer.weight"], w[f"{q}.layers.{i}.in_layer.bias"], epilogue=EPI_SILU, out=p["g_h"])
            ops.gemm(p["g_h"], w[f"{q}.layers.{i}.out_layer.weight"], w[f"{q}.layers.{i}.out_layer.bias"], residual=p["g_x"], out=p["g_x"])
        ops.gemm(p["g_x"], w[q + ".out_proj.weight"], w[q + ".out_proj.bias"], out=p["mod"].view(B * self.n_vec, self.hidden))

    def forward(self, x, timestep, context, y=None, guidance=None) -> torch.Tensor:
        """Reference signature: forward(x, timestep, context, **kwargs) (chroma.py:285)."""
        B, C, H, W = x.shape
        yy = torch.zeros((B, 8), dtype=self.dtype, device=self.device)  # unused placeholder (no pooled-text path)
        tok, _ = self.forward_tokens(x.contiguous(), timestep.float().contiguous(), context.to(self.dtype).contiguous(), yy, None)
        return ops.flux_unpatchify(tok, B, C, H, W, nchw_f32=True)
