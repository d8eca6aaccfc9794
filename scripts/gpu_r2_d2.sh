#!/bin/bash
# round 2, GPU call D2: small-CTA Dh=64 attention kernel (VER 2) after the warpgroup fix
mkdir -p gpurun_out
B200_ATTN64_VER=2 timeout 120 python - <<'PY' 2>&1 | grep -v Warn | tail -12 | tee gpurun_out/d2_first.log
import torch, sys
sys.path.insert(0, ".")
from b200forge import ops
from oracle import ops as O
for (B, H, Lq, Lk) in ((1, 1, 128, 64), (1, 2, 128, 256), (2, 4, 200, 333), (2, 10, 4096, 4096)):
    g = torch.Generator().manual_seed(1)
    q, k, v = (torch.randn(B, L, H * 64, generator=g).half().cuda() for L in (Lq, Lk, Lk))
    o = ops.attention(q, k, v, H) if Lk > 128 else None
    torch.cuda.synchronize()
    if o is not None:
        ref = O.attention(q.float(), k.float(), v.float(), H)
        print(B, H, Lq, Lk, "max_abs", (o.float() - ref).abs().max().item(), "finite", bool(torch.isfinite(o).all()), flush=True)
PY
echo "== attention tests VER=2"
B200_ATTN64_VER=2 B200_ATTN64_CROSS=s timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "attention" -p no:cacheprovider 2>&1 | tail -6 | tee gpurun_out/d2_pytest_attn2.log
echo "== perf"
echo "-- VER=2" | tee -a gpurun_out/d2_attn.log
B200_ATTN64_VER=2 timeout 200 python scripts/kernel_perf.py attention 2>&1 | grep "attention B" | tee -a gpurun_out/d2_attn.log
echo "-- VER=2 cross=s" | tee -a gpurun_out/d2_attn.log
B200_ATTN64_VER=2 B200_ATTN64_CROSS=s timeout 200 python scripts/kernel_perf.py attention 2>&1 | grep "attention B" | tee -a gpurun_out/d2_attn.log
echo "== step time"
B200_ATTN64_VER=2 timeout 300 python scripts/unet_step_time.py sdxl 2>&1 | grep -v Warn | tail -2 | sed "s/^/VER=2 /" | tee gpurun_out/d2_step.log
B200_ATTN64_VER=2 B200_ATTN64_CROSS=s timeout 300 python scripts/unet_step_time.py sdxl 2>&1 | grep -v Warn | tail -2 | sed "s/^/VER=2 cross=s /" | tee -a gpurun_out/d2_step.log
B200_ATTN64_VER=2 timeout 300 python scripts/unet_step_time.py sd15 2>&1 | grep -v Warn | tail -1 | sed "s/^/VER=2 /" | tee -a gpurun_out/d2_step.log
