"""block_n sweep on the transformer GEMM shapes whose wave count quantises badly (N = 1280 at M = 16384: 320 pair-tiles on 74
CTA pairs = 4.32 waves) and on the K = 640 shapes; also the epilogue flavours of the UNet (residual + row statistics for the
producers, folded LayerNorm for the consumers).  Usage: python scripts/gemm_bn_sweep.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from b200forge import ops  # noqa: E402

DEV = "cuda"


def timeit(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters * 1e-3


for (M, N, K) in ((16384, 1280, 1280), (16384, 1280, 5120), (16384, 3840, 1280), (65536, 640, 640), (65536, 640, 2560), (65536, 1920, 640)):
    a = torch.randn(M, K, device=DEV, dtype=torch.float16)
    w = torch.randn(N, K, device=DEV, dtype=torch.float16) * K ** -0.5
    b = torch.randn(N, device=DEV, dtype=torch.float16)
    res = torch.randn(M, N, device=DEV, dtype=torch.float16)
    out = torch.empty(M, N, device=DEV, dtype=torch.float16)
    fl = 2.0 * M * N * K
    for bn in (0, 128, 160, 192, 224, 256):
        if bn and (N % bn and bn != 256):
            continue
        line = f"M={M} N={N} K={K} block_n={bn or 'auto':>4}"
        t = timeit(lambda: ops.gemm(a, w, b, out=out, block_n=bn))
        line += f"  plain {fl / t / 1e12:7.1f} TF/s"
        st = ops.row_stats_buffer(M, N, DEV, block_n=bn)
        t = timeit(lambda: ops.gemm(a, w, b, out=out, residual=res, row_stats_out=st, block_n=bn))
        line += f"  res+stats {fl / t / 1e12:7.1f}"
        if K in (640, 1280):
            stin = ops.row_stats_buffer(M, K, DEV)
            prod = torch.empty(M, K, device=DEV, dtype=torch.float16)
            ops.gemm(a, torch.randn(K, K, device=DEV, dtype=torch.float16) * K ** -0.5, None, out=prod, row_stats_out=stin)
            c = torch.randn(N, device=DEV)
            d = torch.randn(N, device=DEV)
            t = timeit(lambda: ops.gemm(prod, w, None, out=out, ln=(stin, c, d, 1e-5), block_n=bn))
            line += f"  ln-fold {fl / t / 1e12:7.1f}"
        print(line, flush=True)
# GEGLU consumer (ff1): M=16384 K=1280 N=10240 and M=65536 K=640 N=5120
for (M, C) in ((16384, 1280), (65536, 640)):
    a = torch.randn(M, C, device=DEV, dtype=torch.float16)
    w = torch.randn(8 * C, C, device=DEV, dtype=torch.float16) * C ** -0.5
    stin = ops.row_stats_buffer(M, C, DEV)
    prod = torch.empty(M, C, device=DEV, dtype=torch.float16)
    ops.gemm(a, torch.randn(C, C, device=DEV, dtype=torch.float16) * C ** -0.5, None, out=prod, row_stats_out=stin)
    wp, cp = ops.pack_geglu(w, torch.randn(8 * C, device=DEV), 256)
    _, dp = ops.pack_geglu(w, torch.randn(8 * C, device=DEV), 256)
    out = torch.empty(M, 4 * C, device=DEV, dtype=torch.float16)
    fl = 2.0 * M * 8 * C * C
    t = timeit(lambda: ops.gemm(prod, wp, None, epilogue=ops.EPI_GEGLU, block_n=256, out=out, ln=(stin, cp, dp, 1e-5)))
    t2 = timeit(lambda: ops.gemm(prod, wp, None, out=torch.empty(M, 8 * C, device=DEV, dtype=torch.float16)))
    print(f"ff1 M={M} C={C}: ln-fold + GEGLU {fl / t / 1e12:7.1f} TF/s   plain (same weights, no epilogue) {fl / t2 / 1e12:7.1f}", flush=True)
