"""Plug point P1 — drop-in for `backend.attention.attention_function` (reference backend/attention.py:280-339,
430-441) and `attention_function_single_head_spatial` (:342-427, 443-451).

Same signature, argument meaning and result layout as the reference:
    attention_function(q, k, v, heads, mask=None, attn_precision=None, skip_reshape=False) -> [b, Lq, heads*Dh]
q/k/v are [b, L, heads*Dh], or [b, heads, L, Dh] when skip_reshape (Flux).  The fused kernel covers the cases the
reference's txt2img path produces: no mask, fp16/bf16 CUDA tensors, Dh in {64, 128}.  Anything else is *not*
silently emulated: the call is handed to `fallback` (the reference's own function, recorded by plugin.install) or,
when there is none (standalone use), raises `B200Error(B200_EUNSUPPORTED)`.
"""
from __future__ import annotations

from typing import Callable, Optional

import torch

from . import lib as _l
from . import ops

SUPPORTED_HEAD_DIMS = (64, 128)

# the reference implementation to defer to for shapes outside the fast path (set by plugin.install_attention)
fallback: Optional[Callable] = None
fallback_single_head: Optional[Callable] = None


def supports(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, heads: int, mask, skip_reshape: bool) -> bool:
    if mask is not None or not q.is_cuda:
        return False
    # The flash kernels keep logits, softmax and accumulators in fp32 (what attn_precision=fp32 asks of attention_basic,
    # backend/attention.py:64-67), so an fp32 request on fp16/bf16 inputs is honoured by construction.
    if q.dtype not in (torch.float16, torch.bfloat16) or k.dtype != q.dtype or v.dtype != q.dtype:
        return False
    dh = q.shape[-1] if skip_reshape else q.shape[-1] // heads
    return dh in SUPPORTED_HEAD_DIMS


def _unsupported(name, q, heads, mask, skip_reshape):
    dh = q.shape[-1] if skip_reshape else q.shape[-1] // heads
    raise _l.B200Error(_l.E_UNSUPPORTED,
                       f"{name}: no fused path for dtype={q.dtype} device={q.device} head_dim={dh} mask={'yes' if mask is not None else 'no'}")


def attention_function(q, k, v, heads, mask=None, attn_precision=None, skip_reshape=False):
    if not supports(q, k, v, heads, mask, skip_reshape):
        if fallback is not None:
            return fallback(q, k, v, heads, mask, attn_precision, skip_reshape)
        _unsupported("attention_function", q, heads, mask, skip_reshape)
    if skip_reshape:
        # [b, H, L, Dh] -> the kernel's [b, L, H*Dh] addressing needs unit stride on Dh and H*Dh-contiguous rows;
        # Flux produces q/k/v as permuted views of a [b, L, 3, H, Dh] projection, so this is usually a free view.
        b, h, lq, dh = q.shape
        q, k, v = (t.permute(0, 2, 1, 3).reshape(t.shape[0], t.shape[2], h * dh) for t in (q, k, v))
    q, k, v = (t if t.stride(-1) == 1 else t.contiguous() for t in (q, k, v))
    return ops.attention(q, k, v, heads)


def attention_function_single_head_spatial(q, k, v):
    """VAE AttnBlock attention (reference backend/attention.py:412-427): q/k/v NCHW [B, C, H, W], one head of dim C.
    Runs as S = Q K^T -> row softmax -> O = S V on the tcgen05 GEMM, one image at a time."""
    if not (q.is_cuda and q.dtype in (torch.float16, torch.bfloat16)):
        if fallback_single_head is not None:
            return fallback_single_head(q, k, v)
        raise _l.B200Error(_l.E_UNSUPPORTED, f"single-head attention: no fused path for {q.dtype} on {q.device}")
    b, c, hh, ww = q.shape
    L = hh * ww
    qt = ops.nchw_to_nhwc(q.contiguous(), q.dtype).view(b, L, c)
    kt = ops.nchw_to_nhwc(k.contiguous(), q.dtype).view(b, L, c)
    vv = v.contiguous().view(b, c, L)  # NCHW is already V^T [C, L]: the K-major B operand of O = P V
    o = torch.empty((b, L, c), dtype=q.dtype, device=q.device)
    s = torch.empty((L, L), dtype=q.dtype, device=q.device)
    for i in range(b):
        ops.gemm(qt[i], kt[i], out=s, alpha=c ** -0.5)  # scaled logits: unscaled ones can leave the fp16 range at C = 512
        ops.softmax_rows_(s, 1.0)
        ops.gemm(s, vv[i], out=o[i])
    return ops.nhwc_to_nchw(o.view(b, hh, ww, c))
