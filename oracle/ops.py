"""ORACLE — test infrastructure only (imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference leg; never by the product package).

Plain-PyTorch fp32 restatement of the reference's operator-level arithmetic on the hot path.  Each
function cites the reference file:line it follows (paths relative to the reference tree).  The reference
delegates the actual arithmetic of these ops to ATen (F.linear / F.conv2d / F.group_norm / F.layer_norm /
SDPA); the restatement spells the math out with elementary tensor ops so that it is an independent
statement of the algorithm, and `tests/test_oracle_vs_reference.py` pins it against the imported
reference modules (when /root/reference is present) and against committed golden vectors.

Layout: the oracle keeps the reference's layouts (NCHW activations, [b, L, C] tokens).
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

# The reference's call sites bottom out in ATen (F.linear / F.conv2d / F.group_norm / F.layer_norm / SDPA).
# With USE_ATEN the oracle calls those same entry points (this is what is timed as the CPU baseline);
# with USE_ATEN = False every op runs the explicit elementary-op restatement below.  The two are checked
# against each other in tests/test_oracle.py.
USE_ATEN = True


def linear(x: torch.Tensor, w: torch.Tensor, b: torch.Tensor | None = None) -> torch.Tensor:
    """backend/operations.py:149-156 (ForgeOperations.Linear.forward -> F.linear): y = x W^T + b."""
    if USE_ATEN:
        return F.linear(x, w, b)
    y = torch.matmul(x, w.transpose(-1, -2))
    return y if b is None else y + b


def conv2d(x: torch.Tensor, w: torch.Tensor, b: torch.Tensor | None = None, stride: int = 1, padding: int = 1):
    """backend/operations.py:169-176 (ForgeOperations.Conv2d.forward -> _conv_forward), NCHW.
    Restated as unfold (im2col) + matmul: y[n, co, p] = sum_k w[co, k] * patch[n, k, p]."""
    if USE_ATEN:
        return F.conv2d(x, w, b, stride=stride, padding=padding)
    n, c, h, ww = x.shape
    co, ci, kh, kw = w.shape
    cols = torch.nn.functional.unfold(x, (kh, kw), padding=padding, stride=stride)  # [n, ci*kh*kw, P]
    y = torch.matmul(w.reshape(co, -1), cols)
    ho = (h + 2 * padding - kh) // stride + 1
    wo = (ww + 2 * padding - kw) // stride + 1
    y = y.reshape(n, co, ho, wo)
    return y if b is None else y + b.view(1, -1, 1, 1)


def group_norm(x: torch.Tensor, groups: int, gamma: torch.Tensor, beta: torch.Tensor, eps: float) -> torch.Tensor:
    """backend/operations.py:304-310 (F.group_norm), NCHW; biased variance over (C/G, H, W)."""
    if USE_ATEN:
        return F.group_norm(x, groups, gamma, beta, eps)
    n, c = x.shape[:2]
    xg = x.reshape(n, groups, -1)
    mean = xg.mean(dim=2, keepdim=True)
    var = ((xg - mean) ** 2).mean(dim=2, keepdim=True)
    y = ((xg - mean) / torch.sqrt(var + eps)).reshape(x.shape)
    shape = (1, c) + (1,) * (x.dim() - 2)
    return y * gamma.view(shape) + beta.view(shape)


def layer_norm(x: torch.Tensor, gamma: torch.Tensor | None, beta: torch.Tensor | None, eps: float) -> torch.Tensor:
    """backend/operations.py:323-329 (F.layer_norm) over the last dimension."""
    if USE_ATEN:
        return F.layer_norm(x, (x.shape[-1],), gamma, beta, eps)
    mean = x.mean(dim=-1, keepdim=True)
    var = ((x - mean) ** 2).mean(dim=-1, keepdim=True)
    y = (x - mean) / torch.sqrt(var + eps)
    if gamma is not None:
        y = y * gamma
    if beta is not None:
        y = y + beta
    return y


def silu(x: torch.Tensor) -> torch.Tensor:
    """nn.SiLU (backend/nn/unet.py:396,419): x * sigmoid(x)."""
    if USE_ATEN:
        return F.silu(x)
    return x * torch.sigmoid(x)


def gelu_erf(x: torch.Tensor) -> torch.Tensor:
    """F.gelu default (exact erf form), used by GEGLU backend/nn/unet.py:111."""
    if USE_ATEN:
        return F.gelu(x)
    return 0.5 * x * (1.0 + torch.erf(x / math.sqrt(2.0)))


def geglu(x: torch.Tensor, w: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """backend/nn/unet.py:104-111: x, gate = proj(x).chunk(2, -1); x * gelu(gate)."""
    h = linear(x, w, b)
    a, gate = h.chunk(2, dim=-1)
    return a * gelu_erf(gate)


def attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, heads: int) -> torch.Tensor:
    """backend/attention.py:37-93 (attention_basic, the reference's own pure-PyTorch definition of what
    attention_xformers / attention_pytorch compute): softmax(q k^T * Dh^-0.5) v per head, no mask.
    q [b, Lq, H*Dh], k/v [b, Lk, H*Dh] -> [b, Lq, H*Dh]."""
    b, lq, hd = q.shape
    dh = hd // heads
    if USE_ATEN:
        # backend/attention.py:324-339 (attention_pytorch)
        qh, kh, vh = (t.view(b, -1, heads, dh).transpose(1, 2) for t in (q, k, v))
        out = F.scaled_dot_product_attention(qh, kh, vh, attn_mask=None, dropout_p=0.0, is_causal=False)
        return out.transpose(1, 2).reshape(b, -1, heads * dh)
    scale = dh ** -0.5

    def split(t):
        return t.reshape(b, -1, heads, dh).permute(0, 2, 1, 3).reshape(b * heads, -1, dh)

    qh, kh, vh = split(q), split(k), split(v)
    sim = torch.einsum("bid,bjd->bij", qh, kh) * scale
    sim = sim.softmax(dim=-1)
    out = torch.einsum("bij,bjd->bid", sim, vh)
    return out.reshape(b, heads, lq, dh).permute(0, 2, 1, 3).reshape(b, lq, hd)


def timestep_embedding(timesteps: torch.Tensor, dim: int, max_period: float = 10000.0) -> torch.Tensor:
    """backend/nn/unet.py:55-67: [cos | sin], freqs = exp(-ln(max_period) * arange(half) / half), fp32."""
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(start=0, end=half, dtype=torch.float32) / half).to(
        timesteps.device)
    args = timesteps[:, None].float() * freqs[None]
    emb = torch.cat([torch.cos(args), torch.sin(args)], dim=-1)
    if dim % 2:
        emb = torch.cat([emb, torch.zeros_like(emb[:, :1])], dim=-1)
    return emb


def upsample_nearest2x(x: torch.Tensor) -> torch.Tensor:
    """backend/nn/unet.py:352 F.interpolate(mode='nearest') with doubled size, NCHW."""
    if USE_ATEN:
        return F.interpolate(x, size=[x.shape[2] * 2, x.shape[3] * 2], mode="nearest")
    return x.repeat_interleave(2, dim=2).repeat_interleave(2, dim=3)
