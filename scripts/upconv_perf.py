"""Upsample + 3x3 convolution: the folded kernel (b200_conv3x3_up2x, four 2x2 parity filters on the low-res image) beside the
unfolded route (b200_upsample2x + b200_conv3x3) at the UNet / VAE shapes of the SDXL benchmark.  CUDA events, 20 iterations."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from b200forge import ops  # noqa: E402

dev = torch.device("cuda")
shapes = [("unet 32->64", 16, 32, 32, 1280, torch.float16), ("unet 64->128", 16, 64, 64, 640, torch.float16),
          ("vae 128->256", 8, 128, 128, 512, torch.bfloat16), ("vae 256->512", 8, 256, 256, 512, torch.bfloat16),
          ("vae 512->1024", 8, 512, 512, 256, torch.bfloat16)]


def timed(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters


for name, n, h, w, c, dt in shapes:
    g = torch.Generator().manual_seed(1)
    x = torch.randn(n, h, w, c, generator=g).to(dt).to(dev)
    wt = (torch.randn(c, c, 3, 3, generator=g) * (9 * c) ** -0.5).to(dt).to(dev)
    b = torch.randn(c, generator=g).to(dt).to(dev)
    w9, w4 = ops.pack_conv3x3(wt), ops.pack_conv3x3_up2x(wt)
    out = torch.empty((n, 2 * h, 2 * w, c), dtype=dt, device=dev)
    up = torch.empty((n, 2 * h, 2 * w, c), dtype=dt, device=dev)
    t_new = timed(lambda: ops.conv3x3_up2x(x, w4, b, out=out))
    y_new = out.float().clone()
    t_old = timed(lambda: ops.conv3x3(ops.upsample2x(x, out=up), w9, b, out=out))
    diff = (out.float() - y_new)
    rel = (diff.pow(2).mean().sqrt() / y_new.pow(2).mean().sqrt()).item()
    fl_ref = 2.0 * n * 4 * h * w * c * 9 * c
    print(f"{name:14s} {str(dt)[6:]:8s} folded {t_new * 1e3:8.1f} us ({fl_ref * 4 / 9 / t_new / 1e9:6.0f} TF/s executed)   "
          f"upsample+conv {t_old * 1e3:8.1f} us ({fl_ref / t_old / 1e9:6.0f} TF/s)   x{t_old / t_new:.2f}   rel-RMS between routes {rel:.2e}")
