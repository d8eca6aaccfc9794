"""Time the CUDA-graphed UNet forward at the benchmark shape (SDXL batch 16 @128x128, or SD1.5 batch 16 @64x64) and check that the
graph replay is bit-identical to the eager forward.  Usage: python scripts/unet_step_time.py [sdxl|sd15] [iters]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from b200forge import ops, synthetic  # noqa: E402
from b200forge.pipeline import GraphedUNet  # noqa: E402
from b200forge.unet_engine import UNetEngine  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "sdxl"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 20
cfg, hw = (synthetic.SDXL, 128) if wl == "sdxl" else (synthetic.SD15, 64)
dev = torch.device("cuda")
eng = UNetEngine(cfg, synthetic.random_unet_state_dict(cfg, device=dev, dtype=torch.float16, seed=0), dtype=torch.float16, device=dev)
gu = GraphedUNet(eng, 8, 2, hw, hw, 77)
g = torch.Generator().manual_seed(0)
gu.x.copy_(torch.randn(8, 4, hw, hw, generator=g))
gu.sigma.fill_(3.0)
gu.timesteps.fill_(500.0)
ctx = torch.randn(16, 77, cfg["context_dim"], generator=g).half().to(dev)
y = torch.randn(16, cfg["adm_in_channels"], generator=g).half().to(dev) if cfg["adm_in_channels"] else None
gu.set_context(ctx, y)
a = gu().clone()
b = gu._eager().clone()
c = gu().clone()
torch.cuda.synchronize()
print("graph == eager:", torch.equal(a, b), " graph replay reproducible:", torch.equal(a, c), " finite:", bool(torch.isfinite(a).all()))
for _ in range(3):
    gu()
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(iters):
    gu()
e.record()
torch.cuda.synchronize()
print(f"{wl} unet graph: {s.elapsed_time(e) / iters:.3f} ms/step, {gu.launches_per_forward} launches, B200_PDL={os.environ.get('B200_PDL', '1')}")
