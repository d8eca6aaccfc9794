"""Kernel-level timing on the SDXL / Flux shapes of SURVEY.md §8a (CUDA events, L2 flushed between
iterations by cycling through >126 MB of distinct inputs where cheap, else noted).  Prints one line per
kernel with achieved TFLOP/s or GB/s next to the library kernel the reference would call on the same box
(cuBLAS F.linear / cuDNN conv2d / SDPA).  Development aid; bench.py is the graded measurement.
"""
import json
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, ".")
from b200forge import ops  # noqa: E402

DEV = "cuda"
PEAKS = {"bf16_tflops": 1705.8, "hbm_gbs": 6572.2}
try:
    PEAKS.update(json.load(open("MEASURED_PEAKS.json")))
except Exception:
    pass


def timeit(fn, iters=20, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    s = torch.cuda.Event(enable_timing=True)
    e = torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters * 1e-3


def report(name, secs, flops=None, bytes_=None, ref_secs=None):
    msg = f"{name:58s} {secs * 1e6:9.1f} us"
    if flops:
        tf = flops / secs / 1e12
        msg += f"  {tf:7.1f} TF/s ({tf / PEAKS['bf16_tflops'] * 100:5.1f}% of measured burst)"
    if bytes_:
        gb = bytes_ / secs / 1e9
        msg += f"  {gb:7.1f} GB/s ({gb / PEAKS['hbm_gbs'] * 100:5.1f}%)"
    if ref_secs:
        msg += f"  | torch {ref_secs * 1e6:9.1f} us  speedup x{ref_secs / secs:.2f}"
    print(msg, flush=True)


def bench_gemm():
    shapes = [(16384, 1280, 1280), (16384, 3840, 1280), (16384, 10240, 1280), (16384, 1280, 5120),
              (65536, 640, 640), (65536, 1920, 640), (65536, 5120, 640), (65536, 640, 2560),
              (1232, 2560, 2048), (16, 1280, 1280)]
    for M, N, K in shapes:
        a = torch.randn(M, K, device=DEV, dtype=torch.float16)
        w = torch.randn(N, K, device=DEV, dtype=torch.float16) * K ** -0.5
        b = torch.randn(N, device=DEV, dtype=torch.float16)
        out = torch.empty(M, N, device=DEV, dtype=torch.float16)
        t = timeit(lambda: ops.gemm(a, w, b, out=out))
        tr = timeit(lambda: F.linear(a, w, b))
        report(f"gemm M={M} N={N} K={K}", t, flops=2.0 * M * N * K, ref_secs=tr)
    # GEGLU
    for M, C in [(16384, 1280), (65536, 640)]:
        a = torch.randn(M, C, device=DEV, dtype=torch.float16)
        w = torch.randn(8 * C, C, device=DEV, dtype=torch.float16) * C ** -0.5
        b = torch.randn(8 * C, device=DEV, dtype=torch.float16)
        wp, bp = ops.pack_geglu(w, b, 256)
        out = torch.empty(M, 4 * C, device=DEV, dtype=torch.float16)
        t = timeit(lambda: ops.gemm(a, wp, bp, epilogue=ops.EPI_GEGLU, block_n=256, out=out))

        def ref():
            h = F.linear(a, w, b)
            x, g = h.chunk(2, dim=-1)
            return x * F.gelu(g)
        tr = timeit(ref)
        report(f"geglu M={M} C={C}", t, flops=2.0 * M * 8 * C * C, ref_secs=tr)


def bench_conv():
    shapes = [(16, 128, 128, 320, 0, 320), (16, 64, 64, 640, 0, 640), (16, 32, 32, 1280, 0, 1280),
              (16, 32, 32, 1280, 1280, 1280), (16, 64, 64, 1280, 640, 640), (16, 128, 128, 640, 320, 320)]
    for N, H, W, C1, C2, Co in shapes:
        C = C1 + C2
        x1 = torch.randn(N, H, W, C1, device=DEV, dtype=torch.float16)
        x2 = torch.randn(N, H, W, C2, device=DEV, dtype=torch.float16) if C2 else None
        w = torch.randn(Co, C, 3, 3, device=DEV, dtype=torch.float16) * (9 * C) ** -0.5
        b = torch.randn(Co, device=DEV, dtype=torch.float16)
        wp = ops.pack_conv3x3(w)
        out = torch.empty(N, H, W, Co, device=DEV, dtype=torch.float16)
        t = timeit(lambda: ops.conv3x3(x1, wp, b, x2=x2, out=out), iters=10)
        xr = torch.randn(N, C, H, W, device=DEV, dtype=torch.float16)
        tr = timeit(lambda: F.conv2d(xr, w, b, padding=1), iters=10)
        xcl = xr.to(memory_format=torch.channels_last)
        wcl = w.to(memory_format=torch.channels_last)
        tr2 = timeit(lambda: F.conv2d(xcl, wcl, b, padding=1), iters=10)
        report(f"conv3x3 {N}x{H}x{W} {C1}+{C2}->{Co}", t, flops=2.0 * N * H * W * Co * 9 * C, ref_secs=min(tr, tr2))


def bench_attention():
    shapes = [(16, 10, 4096, 4096, 64), (16, 20, 1024, 1024, 64), (16, 10, 4096, 77, 64), (16, 20, 1024, 77, 64),
              (4, 24, 4352, 4352, 128)]
    for B, H, Lq, Lk, Dh in shapes:
        dt = torch.float16 if Dh == 64 else torch.bfloat16
        q = torch.randn(B, Lq, H * Dh, device=DEV, dtype=dt)
        k = torch.randn(B, Lk, H * Dh, device=DEV, dtype=dt)
        v = torch.randn(B, Lk, H * Dh, device=DEV, dtype=dt)
        out = torch.empty_like(q)
        t = timeit(lambda: ops.attention(q, k, v, H, out=out), iters=10)
        qh, kh, vh = (x.view(B, -1, H, Dh).transpose(1, 2) for x in (q, k, v))
        tr = timeit(lambda: F.scaled_dot_product_attention(qh, kh, vh), iters=10)
        report(f"attention B={B} H={H} Lq={Lq} Lk={Lk} Dh={Dh}", t, flops=4.0 * B * H * Lq * Lk * Dh, ref_secs=tr)


def bench_norms():
    for N, H, W, C in [(16, 128, 128, 320), (16, 64, 64, 640), (16, 32, 32, 1280), (16, 128, 128, 960)]:
        x = torch.randn(N, H, W, C, device=DEV, dtype=torch.float16)
        g = torch.randn(C, device=DEV, dtype=torch.float16)
        b = torch.randn(C, device=DEV, dtype=torch.float16)
        out = torch.empty_like(x)
        sums = torch.empty(N, 32, 2, device=DEV, dtype=torch.float32)
        t = timeit(lambda: ops.groupnorm(x, g, b, silu=True, out=out))
        xr = x.permute(0, 3, 1, 2).contiguous()
        tr = timeit(lambda: F.silu(F.group_norm(xr, 32, g, b)))
        report(f"groupnorm+silu {N}x{H}x{W}x{C} (stats+apply, 2 launches)", t, bytes_=3.0 * x.numel() * 2, ref_secs=tr)
    for rows, C in [(65536, 640), (16384, 1280)]:
        x = torch.randn(rows, C, device=DEV, dtype=torch.float16)
        g = torch.randn(C, device=DEV, dtype=torch.float16)
        b = torch.randn(C, device=DEV, dtype=torch.float16)
        out = torch.empty_like(x)
        t = timeit(lambda: ops.layernorm(x, g, b, out=out))
        tr = timeit(lambda: F.layer_norm(x, (C,), g, b))
        report(f"layernorm {rows}x{C}", t, bytes_=2.0 * x.numel() * 2, ref_secs=tr)


if __name__ == "__main__":
    which = sys.argv[1:] or ["gemm", "conv", "attention", "norms"]
    print(torch.cuda.get_device_name(0), PEAKS)
    for w in which:
        globals()["bench_" + w]()
