"""Per-shape time table of one instrumented eager UNet forward (SDXL batch 16 @128x128 by default) [+ VAE decode with `vae`]:
which shapes the step's time goes to, with their achieved TFLOP/s / GB/s.  Events bracket every library call, so launch gaps
are excluded; compare the sum with the graph-replay time.  Usage: python scripts/shape_table.py [sdxl|sd15] [vae]"""
import collections
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from b200forge import ops, synthetic  # noqa: E402
from b200forge.pipeline import GraphedUNet  # noqa: E402
from b200forge.unet_engine import UNetEngine  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "sdxl"
cfg, hw = (synthetic.SDXL, 128) if wl == "sdxl" else (synthetic.SD15, 64)
dev = torch.device("cuda")
eng = UNetEngine(cfg, synthetic.random_unet_state_dict(cfg, device=dev, dtype=torch.float16, seed=0), dtype=torch.float16, device=dev)
gu = GraphedUNet(eng, 8, 2, hw, hw, 77, use_graph=False)
gu.x.normal_()
gu.sigma.fill_(3.0)
gu.timesteps.fill_(500.0)
gu.set_context(torch.randn(16, 77, cfg["context_dim"], device=dev).half(),
               torch.randn(16, cfg["adm_in_channels"], device=dev).half() if cfg["adm_in_channels"] else None)
for _ in range(3):  # warm-up + sustained clocks
    gu._eager()
torch.cuda.synchronize()
ops.PROFILE = []
for _ in range(3):
    gu._eager()
if "vae" in sys.argv:
    from b200forge.vae_engine import VAEDecoderEngine
    vae = VAEDecoderEngine(synthetic.VAE_SDXL, synthetic.random_vae_decoder_state_dict(synthetic.VAE_SDXL, device=dev), device=dev)
    vae.decode(torch.randn(8, 4, hw, hw, device=dev) * 0.13)
torch.cuda.synchronize()
prof, ops.PROFILE = ops.PROFILE, None
agg = collections.OrderedDict()
for fam, fl, by, s, e, label in prof:
    d = agg.setdefault(label or fam, [0, 0.0, 0.0, 0.0])
    d[0] += 1
    d[1] += fl
    d[2] += by
    d[3] += s.elapsed_time(e)
tot = sum(d[3] for d in agg.values()) / 3
print(f"{wl}: sum of bracketed calls {tot:.2f} ms per forward")
for label, (n, fl, by, ms) in sorted(agg.items(), key=lambda kv: -kv[1][3]):
    tf = fl / (ms * 1e-3) / 1e12 if fl else 0.0
    gb = by / (ms * 1e-3) / 1e9
    print(f"{ms / 3:8.3f} ms {100 * ms / 3 / tot:5.1f}%  x{n // 3:<3d} {tf:7.1f} TF/s {gb:7.0f} GB/s  {label}")
