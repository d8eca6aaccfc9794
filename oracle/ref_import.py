"""ORACLE tooling — imports the UNMODIFIED reference (`/root/reference`) in the build container so that
(a) the oracle restatement can be pinned against the reference's own modules and (b) golden vectors can
be generated (oracle/gen_golden.py).  `/root/reference` does not exist on the GPU box: nothing that runs
there imports this module (tests that need it skip when the tree is absent).

The reference's backend/ and k_diffusion/ import on CPU with four stub packages for dependencies that
are not installed here: diffusers (ConfigMixin/register_to_config used as a base class +
decorator; two one-line Flux schedule helpers), torchsde and torchdiffeq (SDE/adaptive samplers only).
No reference source is copied; the stubs below contain no reference code.
"""
from __future__ import annotations

import os
import sys
import types

REF_ROOT = os.environ.get("B200FORGE_REFERENCE", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "backend"))


def _install_stubs() -> None:
    if "diffusers" not in sys.modules:
        diffusers = types.ModuleType("diffusers")
        cfg = types.ModuleType("diffusers.configuration_utils")

        class ConfigMixin:  # only used as a base class by the reference's nn modules
            config_name = None

        def register_to_config(init):
            return init

        cfg.ConfigMixin = ConfigMixin
        cfg.register_to_config = register_to_config

        class FlowMatchEulerDiscreteScheduler:
            def time_shift(self, mu, sigma, t):  # called unbound with self=None (k_prediction.py:300)
                import math
                return math.exp(mu) / (math.exp(mu) + (1 / t - 1) ** sigma)

        pipelines = types.ModuleType("diffusers.pipelines")
        flux = types.ModuleType("diffusers.pipelines.flux")
        pflux = types.ModuleType("diffusers.pipelines.flux.pipeline_flux")

        def calculate_shift(image_seq_len, base_seq_len=256, max_seq_len=4096, base_shift=0.5, max_shift=1.16):
            m = (max_shift - base_shift) / (max_seq_len - base_seq_len)
            b = base_shift - m * base_seq_len
            return image_seq_len * m + b

        pflux.calculate_shift = calculate_shift
        diffusers.configuration_utils = cfg
        diffusers.FlowMatchEulerDiscreteScheduler = FlowMatchEulerDiscreteScheduler
        diffusers.pipelines = pipelines
        pipelines.flux = flux
        flux.pipeline_flux = pflux
        sys.modules.update({
            "diffusers": diffusers, "diffusers.configuration_utils": cfg, "diffusers.pipelines": pipelines,
            "diffusers.pipelines.flux": flux, "diffusers.pipelines.flux.pipeline_flux": pflux,
        })
    for name in ("torchsde", "torchdiffeq"):
        if name not in sys.modules:
            m = types.ModuleType(name)
            if name == "torchdiffeq":
                m.odeint = None
            sys.modules[name] = m


_loaded = False


def load():
    """Make `import backend...` / `import k_diffusion...` resolve to the reference tree (CPU, SDPA attention)."""
    global _loaded
    if _loaded:
        return
    if not available():
        raise RuntimeError(f"reference tree not found at {REF_ROOT}")
    _install_stubs()
    saved_argv = sys.argv
    sys.argv = [saved_argv[0] if saved_argv else "oracle", "--always-cpu", "--attention-pytorch"]
    sys.path.insert(0, REF_ROOT)
    sys.path.insert(0, os.path.join(REF_ROOT, "packages_3rdparty"))
    try:
        import backend.args  # noqa: F401  (flags are parsed at import time, backend/args.py:61)
        import backend.attention  # noqa: F401
        import backend.nn.unet  # noqa: F401
        import backend.nn.vae  # noqa: F401
        import backend.modules.k_prediction  # noqa: F401
        import k_diffusion.sampling  # noqa: F401
    finally:
        sys.argv = saved_argv
    _loaded = True
