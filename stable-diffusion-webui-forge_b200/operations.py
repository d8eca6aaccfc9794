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

import weakref

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
    """Packed-weight cache keyed on the Parameter OBJECT (weak reference) and its `_version`: a LoRA refresh
    (backend/patcher/lora.py:352-446) installs fresh Parameters whose storage the caching allocator may place at the freed
    address of the previous one, so (data_ptr, _version) alone could serve a stale packed weight."""

    def __init__(self):
        self.ref = None
        self.key = None
        self.val = None

    def get(self, p: torch.Tensor, fn):
        key = (p._version, p.data_ptr(), p.dtype)
        if self.ref is None or self.ref() is not p or key != self.key:
            self.val = fn(p)
            self.key = key
            self.ref = weakref.ref(p)
        return self.val

    def clear(self):
        self.ref = self.key = self.val = None


def _plain(mod) -> bool:
    """The module holds ordinary resident parameters: no manual cast (storage dtype != computation dtype, fp8 / offloaded
    weights: backend/operations.py:57-97), no on-the-fly LoRA (`forge_online_loras`, :14-54), no fp8 scale, weights loaded."""
    return (not getattr(mod, "parameters_manual_cast", False) and not hasattr(mod, "forge_online_loras")
            and getattr(mod, "scale_weight", None) is None and getattr(mod, "weight", None) is not None)


class TorchOperations:
    """The stock torch modules as an operator set (what `make_operations` builds on outside Forge)."""
    Linear = nn.Linear
    Conv1d = nn.Conv1d
    Conv2d = nn.Conv2d
    Conv3d = nn.Conv3d
    ConvTranspose1d = nn.ConvTranspose1d
    ConvTranspose2d = nn.ConvTranspose2d
    ConvTranspose3d = nn.ConvTranspose3d
    GroupNorm = nn.GroupNorm
    LayerNorm = nn.LayerNorm
    Embedding = nn.Embedding


def make_operations(base):
    """Operator set whose Linear / Conv2d / GroupNorm / LayerNorm SUBCLASS `base`'s own classes: inside Forge `base` is
    `backend.operations.ForgeOperations`, so lazy weight creation (`dummy`, `_load_from_state_dict`), `parameters_manual_cast`
    / `weights_manual_cast`, fp8 `scale_weight` and `forge_online_loras` keep their reference behaviour — every call the fused
    kernels do not cover goes to the parent's `forward` (the reference's own code), not to a bare torch module."""

    class Linear(base.Linear):
        def forward(self, x):
            if (not _plain(self) or not _fast(x, self.weight) or self.in_features % 8 or self.out_features % 8):
                _defer(self, x)
                return super().forward(x)
            x2 = x.reshape(-1, self.in_features)
            if x2.stride(-1) != 1 or x2.stride(0) % 8:
                x2 = x2.contiguous()
            y = ops.gemm(x2, self.weight.detach(), None if self.bias is None else self.bias.detach())
            return y.view(*x.shape[:-1], self.out_features)

    class Conv2d(base.Conv2d):
        def __init__(self, *a, **kw):
            super().__init__(*a, **kw)
            self._packed = _Cache()

        def _load_from_state_dict(self, *a, **kw):
            self._packed.clear()
            return super()._load_from_state_dict(*a, **kw)

        def _supported(self, x):
            return (_plain(self) and _fast(x, self.weight) and self.groups == 1 and self.dilation == (1, 1) and
                    self.padding_mode == "zeros" and self.kernel_size in ((1, 1), (3, 3)) and self.out_channels % 8 == 0 and
                    self.in_channels % 8 == 0 and isinstance(self.padding, tuple))

        def forward(self, x):
            if x.dim() != 4 or not self._supported(x):
                _defer(self, x)
                return super().forward(x)
            n, c, h, w = x.shape
            bias = None if self.bias is None else self.bias.detach()
            if self.kernel_size == (1, 1):
                if self.stride != (1, 1) or self.padding != (0, 0):
                    _defer(self, x)
                    return super().forward(x)
                xn = ops.nchw_to_nhwc(x.contiguous(), x.dtype)
                wp = self._packed.get(self.weight, lambda p: p.detach().reshape(self.out_channels, c).contiguous())
                y = ops.gemm(xn.view(-1, c), wp, bias).view(n, h, w, self.out_channels)
                return ops.nhwc_to_nchw(y)
            if self.stride[0] != self.stride[1] or self.padding[0] != self.padding[1] or self.padding[0] > 1:
                _defer(self, x)
                return super().forward(x)
            xn = ops.nchw_to_nhwc(x.contiguous(), x.dtype)
            wp = self._packed.get(self.weight, lambda p: ops.pack_conv3x3(p.detach()))
            if self.stride == (1, 1) and self.padding == (1, 1):
                return ops.nhwc_to_nchw(ops.conv3x3_any(xn, wp, bias))
            cols = ops.im2col3x3(xn, stride=self.stride[0], pad_lo=self.padding[0], pad_hi=self.padding[0])
            ho = (h + 2 * self.padding[0] - 3) // self.stride[0] + 1
            wo = (w + 2 * self.padding[0] - 3) // self.stride[0] + 1
            y = ops.gemm(cols, wp, bias).view(n, ho, wo, self.out_channels)
            return ops.nhwc_to_nchw(y)

    class GroupNorm(base.GroupNorm):
        def forward(self, x):
            if (x.dim() != 4 or not self.affine or not _plain(self) or not _fast(x, self.weight) or self.num_channels % 8):
                _defer(self, x)
                return super().forward(x)
            xn = ops.nchw_to_nhwc(x.contiguous(), x.dtype)
            y = ops.groupnorm(xn, self.weight.detach(), self.bias.detach(), groups=self.num_groups, eps=self.eps, silu=False)
            return ops.nhwc_to_nchw(y)

    class LayerNorm(base.LayerNorm):
        def forward(self, x):
            affine = self.elementwise_affine
            if (len(self.normalized_shape) != 1 or not x.is_cuda or x.dtype not in _FAST_DTYPES or
                    self.normalized_shape[0] % 8 or self.normalized_shape[0] > 4096 or
                    getattr(self, "parameters_manual_cast", False) or (affine and self.weight.dtype != x.dtype)):
                _defer(self, x)
                return super().forward(x)
            xc = x.contiguous()
            g = self.weight.detach() if affine else None
            b = self.bias.detach() if (affine and self.bias is not None) else None
            return ops.layernorm(xc, g, b, self.eps)

    return type("B200Operations", (base,), dict(Linear=Linear, Conv2d=Conv2d, GroupNorm=GroupNorm, LayerNorm=LayerNorm,
                                                __doc__="Attribute set expected by using_forge_operations "
                                                        "(backend/operations.py:455): the four hot-path ops run the sm_100a kernels, "
                                                        "everything else (and every call they do not cover) is the base set's."))


# outside Forge (tests, standalone use): the same classes over the stock torch modules
B200Operations = make_operations(TorchOperations)
Linear, Conv2d, GroupNorm, LayerNorm = (B200Operations.Linear, B200Operations.Conv2d, B200Operations.GroupNorm,
                                        B200Operations.LayerNorm)
