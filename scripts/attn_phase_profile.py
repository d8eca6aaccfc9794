"""Runs the instrumented Dh=64 attention kernel once (libb200forge_prof.so, built with -DB200_ATTN_PROFILE:
`scripts/build_profile_lib.sh`) and lets its device-side printf report the clock64 phase breakdown per key block.

    B200FORGE_LIB=stable-diffusion-webui-forge_b200/libb200forge_prof.so python scripts/attn_phase_profile.py
"""
import sys

import torch

sys.path.insert(0, ".")
from b200forge import ops  # noqa: E402

for (B, H, L) in ((16, 10, 4096), (16, 20, 1024)):
    q, k, v = (torch.randn(B, L, H * 64, device="cuda", dtype=torch.float16) for _ in range(3))
    ops.attention(q, k, v, H)
    torch.cuda.synchronize()
    print(f"-- B={B} H={H} L={L}", flush=True)
    ops.attention(q, k, v, H)
    torch.cuda.synchronize()
