"""Plug point P2 — operator classes for `backend.operations.using_forge_operations(operations=...)`
(reference backend/operations.py:441-467).  `B200Operations` exposes the attributes the reference's context manager
swaps into `torch.nn` (`Linear, Conv2d, GroupNorm, LayerNorm` + the untouched torch classes for the rest); each is an
`nn.Module` with the torch constructor signature, `weight`/`bias` as `nn.Parameter` (so LoRA patching and
`load_state_dict` keep working) and a `forward` that runs the sm_100a kernels.

These entry points accept the reference's layouts (NCHW activations, [..., C] tokens) and convert at the op boundary,
which is what a module-by-module drop-in must do; the fused engines (unet_engine / vae_engine) stay channels-last end
to end and are the fast path.  Inputs the kernels do not cover (fp32 modules, CPU tensors, grouped / dilated convs,
odd kernel sizes) go to the torch implementation of the parent class — the same thing ForgeOperations does
(backend/operations.py:149-156, 169-176, 304-310, 323-329).
"""
from __future__ import annotations

import torch
from torch import nn

from . import ops

import os

from . import lib as _l

_FAST_DTYPES = (torch.float16, torch.bfloat16)
# B200_STRICT=1: inputs outside the fused path raise instead of deferring to the stock torch module
STRICT = os.environ.get("B200_STRICT", "0") == "1"
DEFERRED = 0  # calls that were handed to the stock torch implementation (observable by tests / the plug-in)


def _defer(mod, x):
    global DEFERRED
    if STRICT:
        raise _l.B200Error(_l.E_UNSUPPORTED, f"{type(mod).__name__}: no fused path for input {tuple(x.shape)} {x.dtype} on {x.device}")
    DEFERRED += 1


def _fast(x: torch.Tensor, w: torch.Tensor) -> bool:
    return x.is_cuda and x.dtype in _FAST_DTYPES and w.dtype == x.dtype and w.is_cuda


class _Cache:
    """Packed-weight cache keyed on (data_ptr, _version): LoRA refresh replaces/rewrites parameters
    (backend/patcher/lora.py), which must invalidate the packed copy."""

    def __init__(self):
        self.key = None
        self.val = None

    def get(self, p: torch.Tensor, fn):
        key = (p.data_ptr(), p._version, p.dtype)
        if key != self.key:
            self.val = fn(p)
            self.key = key
        return self.val


class Linear(nn.Linear):
    def forward(self, x):
        if not _fast(x, self.weight) or self.in_features % 8 or self.out_features % 8:
            _defer(self, x)
            return super().forward(x)
        x2 = x.reshape(-1, self.in_features)
        if x2.stride(-1) != 1 or x2.stride(0) % 8:
            x2 = x2.contiguous()
        y = ops.gemm(x2, self.weight.detach(), None if self.bias is None else self.bias.detach())
        return y.view(*x.shape[:-1], self.out_features)


class Conv2d(nn.Conv2d):
    def __init__(self, *a, **kw):
        super().__init__(*a, **kw)
        self._packed = _Cache()

    def _supported(self, x):
        return (_fast(x, self.weight) and self.groups == 1 and self.dilation == (1, 1) and
                self.padding_mode == "zeros" and self.kernel_size in ((1, 1), (3, 3)) and self.out_channels % 8 == 0 and
                self.in_channels % 8 == 0 and isinstance(self.padding, tuple))

    def forward(self, x):
        if x.dim() != 4 or not self._supported(x):
            _defer(self, x)
            return super().forward(x)
        n, c, h, w = x.shape
        bias = None if self.bias is None else self.bias.detach()
        xn = ops.nchw_to_nhwc(x.contiguous(), x.dtype)
        if self.kernel_size == (1, 1):
            if self.stride != (1, 1) or self.padding != (0, 0):
                _defer(self, x)
                return super().forward(x)
            wp = self._packed.get(self.weight, lambda p: p.detach().reshape(self.out_channels, c).contiguous())
            y = ops.gemm(xn.view(-1, c), wp, bias).view(n, h, w, self.out_channels)
            return ops.nhwc_to_nchw(y)
        wp = self._packed.get(self.weight, lambda p: ops.pack_conv3x3(p.detach()))
        if self.stride == (1, 1) and self.padding == (1, 1) and c % 64 == 0 and (128 % min(w, 128) == 0) and w % min(w, 128) == 0:
            try:
                return ops.nhwc_to_nchw(ops.conv3x3(xn, wp, bias))
            except ops.B200Error as e:  # tiling not expressible -> im2col route below
                if e.code not in (-1, -2):
                    raise
        if self.stride[0] != self.stride[1] or self.padding[0] != self.padding[1] or self.padding[0] > 1:
            _defer(self, x)
            return super().forward(x)
        cols = ops.im2col3x3(xn, stride=self.stride[0], pad_lo=self.padding[0], pad_hi=self.padding[0])
        ho = (h + 2 * self.padding[0] - 3) // self.stride[0] + 1
        wo = (w + 2 * self.padding[0] - 3) // self.stride[0] + 1
        y = ops.gemm(cols, wp, bias).view(n, ho, wo, self.out_channels)
        return ops.nhwc_to_nchw(y)


class GroupNorm(nn.GroupNorm):
    def forward(self, x):
        if (x.dim() != 4 or not self.affine or not _fast(x, self.weight) or self.num_channels % 8):
            _defer(self, x)
            return super().forward(x)
        xn = ops.nchw_to_nhwc(x.contiguous(), x.dtype)
        y = ops.groupnorm(xn, self.weight.detach(), self.bias.detach(), groups=self.num_groups, eps=self.eps, silu=False)
        return ops.nhwc_to_nchw(y)


class LayerNorm(nn.LayerNorm):
    def forward(self, x):
        if (len(self.normalized_shape) != 1 or not x.is_cuda or x.dtype not in _FAST_DTYPES or
                self.normalized_shape[0] % 8 or self.normalized_shape[0] > 4096 or
                (self.elementwise_affine and self.weight.dtype != x.dtype)):
            _defer(self, x)
            return super().forward(x)
        xc = x.contiguous()
        g = self.weight.detach() if self.elementwise_affine else None
        b = self.bias.detach() if (self.elementwise_affine and self.bias is not None) else None
        return ops.layernorm(xc, g, b, self.eps)


class B200Operations:
    """Attribute set expected by using_forge_operations (backend/operations.py:455): the four hot-path ops are ours,
    the rest are the stock torch modules."""
    Linear = Linear
    Conv1d = nn.Conv1d
    Conv2d = Conv2d
    Conv3d = nn.Conv3d
    ConvTranspose1d = nn.ConvTranspose1d
    ConvTranspose2d = nn.ConvTranspose2d
    ConvTranspose3d = nn.ConvTranspose3d
    GroupNorm = GroupNorm
    LayerNorm = LayerNorm
    Embedding = nn.Embedding
