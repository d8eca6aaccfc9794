"""ORACLE — test infrastructure only (see oracle/ops.py header).

Functional restatement of the reference's Chroma transformer forward (backend/nn/chroma.py:138-307): Flux's double /
single stream blocks (oracle/flux.py) whose modulation vectors come from one Approximator MLP evaluated on
[timestep embedding | zero-guidance embedding | modulation index embedding] instead of per-block Modulation linears.
Pinned by tests/golden/chroma_tiny.pt (imported reference, oracle/gen_golden.py).
"""
from __future__ import annotations

import torch

from . import flux as OF
from . import ops as O

SD = OF.SD

TINY_CHROMA = dict(in_channels=16, vec_in_dim=32, context_in_dim=64, hidden_size=256, mlp_ratio=4.0, num_heads=2, depth=2,
                   depth_single_blocks=2, axes_dim=[16, 56, 56], theta=10000, qkv_bias=True, guidance_out_dim=256,
                   guidance_hidden_dim=320, guidance_n_layers=3)
CONFIGS = {"tiny_chroma": TINY_CHROMA}


def n_mod_vectors(cfg: dict) -> int:
    return cfg["depth"] * 12 + cfg["depth_single_blocks"] * 3 + 2


def approximator(sd: SD, cfg: dict, x: torch.Tensor) -> torch.Tensor:
    """Approximator.forward (chroma.py:14-28): in_proj, n x (x + MLPEmbedder(RMSNorm(x))), out_proj."""
    p = "distilled_guidance_layer"
    x = O.linear(x, sd[p + ".in_proj.weight"], sd[p + ".in_proj.bias"])
    for i in range(cfg["guidance_n_layers"]):
        h = OF.rms_norm(x, sd[f"{p}.norms.{i}.scale"])
        h = O.linear(O.silu(O.linear(h, sd[f"{p}.layers.{i}.in_layer.weight"], sd[f"{p}.layers.{i}.in_layer.bias"])),
                     sd[f"{p}.layers.{i}.out_layer.weight"], sd[f"{p}.layers.{i}.out_layer.bias"])
        x = x + h
    return O.linear(x, sd[p + ".out_proj.weight"], sd[p + ".out_proj.bias"])


def modulation_vectors(sd: SD, cfg: dict, timesteps: torch.Tensor, dtype) -> torch.Tensor:
    """chroma.py:255-263: [B, n_vec, hidden]."""
    n = n_mod_vectors(cfg)
    B = timesteps.shape[0]
    t16 = OF.timestep_embedding(timesteps, 16).to(dtype)
    g16 = OF.timestep_embedding(torch.zeros_like(timesteps), 16).to(dtype)
    # integer input in the reference: 1000 * i is exact in int64 and in fp32 alike, and the embedding stays fp32 until this cast
    idx = OF.timestep_embedding(torch.arange(n).float(), 32).to(dtype)
    tg = torch.cat([t16, g16], 1).unsqueeze(1).repeat(1, n, 1)
    return approximator(sd, cfg, torch.cat([tg, idx.unsqueeze(0).repeat(B, 1, 1)], -1))


def distribute(cfg: dict):
    """distribute_modulations (chroma.py:181-243): vector index of the first vector of every block, in the reference's
    order — all single blocks (3 each), then img_mod of every double block (6 each), then txt_mod (6 each), then the final
    layer (2)."""
    off, idx = {}, 0
    for i in range(cfg["depth_single_blocks"]):
        off[f"single_blocks.{i}.modulation.lin"] = idx
        idx += 3
    for s in ("img", "txt"):
        for i in range(cfg["depth"]):
            off[f"double_blocks.{i}.{s}_mod.lin"] = idx
            idx += 6
    off["final_layer.adaLN_modulation.1"] = idx
    return off


def chroma_forward(sd: SD, cfg: dict, x: torch.Tensor, timestep: torch.Tensor, context: torch.Tensor) -> torch.Tensor:
    """IntegratedChromaTransformer2DModel.forward (chroma.py:246-307) for even latent sizes."""
    B, C, Hh, Ww = x.shape
    H, hidden = cfg["num_heads"], cfg["hidden_size"]
    img = O.linear(OF.patchify(x), sd["img_in.weight"], sd["img_in.bias"])
    mod = modulation_vectors(sd, cfg, timestep, img.dtype)
    off = distribute(cfg)

    def vec(name, k):
        return mod[:, off[name] + k:off[name] + k + 1, :]

    txt = O.linear(context, sd["txt_in.weight"], sd["txt_in.bias"])
    Lt = txt.shape[1]
    cos, sin = OF.rope_tables(OF.position_ids(Hh // 2, Ww // 2, Lt), cfg["axes_dim"], cfg["theta"])
    for i in range(cfg["depth"]):
        p = f"double_blocks.{i}"
        im, tm = p + ".img_mod.lin", p + ".txt_mod.lin"
        iq, ik, iv = OF._split_heads(O.linear((1 + vec(im, 1)) * OF._ln(img) + vec(im, 0), sd[p + ".img_attn.qkv.weight"], sd.get(p + ".img_attn.qkv.bias")), H)
        iq, ik = OF.rms_norm(iq, sd[p + ".img_attn.norm.query_norm.scale"]), OF.rms_norm(ik, sd[p + ".img_attn.norm.key_norm.scale"])
        tq, tk, tv = OF._split_heads(O.linear((1 + vec(tm, 1)) * OF._ln(txt) + vec(tm, 0), sd[p + ".txt_attn.qkv.weight"], sd.get(p + ".txt_attn.qkv.bias")), H)
        tq, tk = OF.rms_norm(tq, sd[p + ".txt_attn.norm.query_norm.scale"]), OF.rms_norm(tk, sd[p + ".txt_attn.norm.key_norm.scale"])
        attn = OF._attention(torch.cat((tq, iq), 2), torch.cat((tk, ik), 2), torch.cat((tv, iv), 2), cos, sin)
        t_attn, i_attn = attn[:, :Lt], attn[:, Lt:]

        def mlp(q, xx):
            return O.linear(OF.gelu_tanh(O.linear(xx, sd[q + ".0.weight"], sd[q + ".0.bias"])), sd[q + ".2.weight"], sd[q + ".2.bias"])

        img = img + vec(im, 2) * O.linear(i_attn, sd[p + ".img_attn.proj.weight"], sd[p + ".img_attn.proj.bias"])
        img = img + vec(im, 5) * mlp(p + ".img_mlp", (1 + vec(im, 4)) * OF._ln(img) + vec(im, 3))
        txt = txt + vec(tm, 2) * O.linear(t_attn, sd[p + ".txt_attn.proj.weight"], sd[p + ".txt_attn.proj.bias"])
        txt = txt + vec(tm, 5) * mlp(p + ".txt_mlp", (1 + vec(tm, 4)) * OF._ln(txt) + vec(tm, 3))
    xx = torch.cat((txt, img), 1)
    for i in range(cfg["depth_single_blocks"]):
        p = f"single_blocks.{i}"
        m = p + ".modulation.lin"
        yv = O.linear((1 + vec(m, 1)) * OF._ln(xx) + vec(m, 0), sd[p + ".linear1.weight"], sd[p + ".linear1.bias"])
        q, k, v = OF._split_heads(yv[..., :3 * hidden], H)
        q, k = OF.rms_norm(q, sd[p + ".norm.query_norm.scale"]), OF.rms_norm(k, sd[p + ".norm.key_norm.scale"])
        attn = OF._attention(q, k, v, cos, sin)
        xx = xx + vec(m, 2) * O.linear(torch.cat((attn, OF.gelu_tanh(yv[..., 3 * hidden:])), 2), sd[p + ".linear2.weight"], sd[p + ".linear2.bias"])
    img = xx[:, Lt:]
    f = "final_layer.adaLN_modulation.1"
    img = (1 + vec(f, 1)) * OF._ln(img) + vec(f, 0)
    out = O.linear(img, sd["final_layer.linear.weight"], sd["final_layer.linear.bias"])
    return OF.unpatchify(out, C, Hh, Ww)


def random_state_dict(cfg: dict, seed: int = 0, dtype=torch.float32) -> SD:
    """Flux block weights (oracle.flux conventions) + the Approximator; the per-block Modulation linears, time / vector /
    guidance embedders and the adaLN linear of Flux do not exist in Chroma."""
    flux_cfg = dict(cfg, guidance_embed=False)
    sd = {k: v for k, v in OF.random_state_dict(flux_cfg, seed=seed, dtype=dtype).items()
          if not any(t in k for t in ("_mod.lin", "modulation.lin", "adaLN_modulation", "time_in.", "vector_in.", "guidance_in."))}
    g = torch.Generator().manual_seed(seed + 1000)
    hd, od = cfg["guidance_hidden_dim"], cfg["guidance_out_dim"]

    def lin(p, cin, cout, wscale=1.0, bmean=0.0):
        sd[p + ".weight"] = (torch.randn(cout, cin, generator=g) * (wscale * cin ** -0.5)).to(dtype)
        sd[p + ".bias"] = (bmean + torch.randn(cout, generator=g) * 0.05).to(dtype)

    p = "distilled_guidance_layer"
    lin(p + ".in_proj", 64, hd)
    for i in range(cfg["guidance_n_layers"]):
        lin(f"{p}.layers.{i}.in_layer", hd, hd)
        lin(f"{p}.layers.{i}.out_layer", hd, hd, wscale=0.5)
        sd[f"{p}.norms.{i}.scale"] = (1.0 + 0.1 * torch.randn(hd, generator=g)).to(dtype)
    lin(p + ".out_proj", hd, od, wscale=0.3, bmean=0.1)
    return sd
