"""Tiny launcher for ncu captures: runs a handful of representative kernels once each (after one warm-up)."""
import sys
import torch
sys.path.insert(0, ".")
from b200forge import ops

DEV = "cuda"
which = sys.argv[1] if len(sys.argv) > 1 else "gemm"


def gemm(M, N, K):
    a = torch.randn(M, K, device=DEV, dtype=torch.float16)
    w = torch.randn(N, K, device=DEV, dtype=torch.float16) * K ** -0.5
    b = torch.randn(N, device=DEV, dtype=torch.float16)
    out = torch.empty(M, N, device=DEV, dtype=torch.float16)
    for _ in range(2):
        ops.gemm(a, w, b, out=out)
    torch.cuda.synchronize()


def conv(N, H, W, C, Co):
    x = torch.randn(N, H, W, C, device=DEV, dtype=torch.float16)
    w = torch.randn(Co, 9 * C, device=DEV, dtype=torch.float16) * (9 * C) ** -0.5
    b = torch.randn(Co, device=DEV, dtype=torch.float16)
    out = torch.empty(N, H, W, Co, device=DEV, dtype=torch.float16)
    for _ in range(2):
        ops.conv3x3(x, w, b, out=out)
    torch.cuda.synchronize()


def attn(B, H, Lq, Lk, Dh, dtype=torch.float16):
    q = torch.randn(B, Lq, H * Dh, device=DEV, dtype=dtype)
    k = torch.randn(B, Lk, H * Dh, device=DEV, dtype=dtype)
    v = torch.randn(B, Lk, H * Dh, device=DEV, dtype=dtype)
    out = torch.empty_like(q)
    for _ in range(2):
        ops.attention(q, k, v, H, out=out)
    torch.cuda.synchronize()


def gn(N, H, W, C):
    x = torch.randn(N, H, W, C, device=DEV, dtype=torch.float16)
    g = torch.randn(C, device=DEV, dtype=torch.float16)
    b = torch.randn(C, device=DEV, dtype=torch.float16)
    for _ in range(2):
        ops.groupnorm(x, g, b, silu=True)
    torch.cuda.synchronize()


def unet_gemms():
    """The transformer GEMM flavours of the SDXL UNet at the 1280 level: producer (bias + residual + row statistics), LayerNorm-
    fold consumer (QKV) and LayerNorm-fold + GEGLU (feed-forward), each launched twice (the second is the one to read)."""
    M, C = 16384, 1280
    a = torch.randn(M, C, device=DEV, dtype=torch.float16)
    w = torch.randn(C, C, device=DEV, dtype=torch.float16) * C ** -0.5
    b = torch.randn(C, device=DEV, dtype=torch.float16)
    res = torch.randn(M, C, device=DEV, dtype=torch.float16)
    t = torch.empty(M, C, device=DEV, dtype=torch.float16)
    st = ops.row_stats_buffer(M, C, DEV)
    for _ in range(2):
        ops.gemm(a, w, b, residual=res, out=t, row_stats_out=st)
    wq = torch.randn(3 * C, C, device=DEV, dtype=torch.float16) * C ** -0.5
    c = torch.randn(3 * C, device=DEV)
    d = torch.randn(3 * C, device=DEV)
    for _ in range(2):
        ops.gemm(t, wq, None, ln=(st, c, d, 1e-5))
    wf = torch.randn(8 * C, C, device=DEV, dtype=torch.float16) * C ** -0.5
    wp, cp = ops.pack_geglu(wf, torch.randn(8 * C, device=DEV), 256)
    _, dp = ops.pack_geglu(wf, torch.randn(8 * C, device=DEV), 256)
    for _ in range(2):
        ops.gemm(t, wp, None, epilogue=ops.EPI_GEGLU, block_n=256, ln=(st, cp, dp, 1e-5))
    wo = torch.randn(C, 4 * C, device=DEV, dtype=torch.float16) * (4 * C) ** -0.5
    g = torch.randn(M, 4 * C, device=DEV, dtype=torch.float16)
    for _ in range(2):
        ops.gemm(g, wo, b, residual=res, out=t, row_stats_out=st)
    torch.cuda.synchronize()


def producer_gemms():
    """Producer flavour (bias + residual + row statistics) at the two transformer levels, each twice (read the second)."""
    for M, C in ((16384, 1280), (65536, 640)):
        a = torch.randn(M, C, device=DEV, dtype=torch.float16)
        w = torch.randn(C, C, device=DEV, dtype=torch.float16) * C ** -0.5
        b = torch.randn(C, device=DEV, dtype=torch.float16)
        res = torch.randn(M, C, device=DEV, dtype=torch.float16)
        t = torch.empty(M, C, device=DEV, dtype=torch.float16)
        st = ops.row_stats_buffer(M, C, DEV)
        for _ in range(2):
            ops.gemm(a, w, b, residual=res, out=t, row_stats_out=st)
    torch.cuda.synchronize()


def upconv(N, H, W, C, dtype=torch.float16):
    x = torch.randn(N, H, W, C, device=DEV, dtype=dtype)
    w4 = ops.pack_conv3x3_up2x((torch.randn(C, C, 3, 3, device=DEV) * (9 * C) ** -0.5).to(dtype))
    b = torch.randn(C, device=DEV, dtype=dtype)
    out = torch.empty(N, 2 * H, 2 * W, C, device=DEV, dtype=dtype)
    for _ in range(2):
        ops.conv3x3_up2x(x, w4, b, out=out)
    torch.cuda.synchronize()


if which == "upconv":      # the folded upsample convolutions of the SDXL UNet (32 -> 64) and of the VAE decoder (256 -> 512)
    upconv(16, 32, 32, 1280)
    upconv(8, 256, 256, 512, torch.bfloat16)
elif which == "attn128":   # Flux: 4096 + 256 tokens, 24 heads x 128
    attn(4, 24, 4352, 4352, 128, torch.bfloat16)
elif which == "ksplit":    # conv 1280 -> 1280 at 32 x 32: 320 pair tiles on 74 pairs, tail K-split
    conv(16, 32, 32, 1280, 1280)
elif which == "unetgemm":
    unet_gemms()
elif which == "producer":
    producer_gemms()
elif which == "all":
    gemm(16384, 10240, 1280)
    conv(16, 128, 128, 320, 320)
    attn(16, 10, 4096, 4096, 64)
    attn(4, 24, 4352, 4352, 128, torch.bfloat16)  # Flux.1-dev joint attention
    gn(16, 128, 128, 320)
elif which == "gemm":
    gemm(16384, 10240, 1280)
    gemm(65536, 5120, 640)
    conv(16, 128, 128, 320, 320)
elif which == "attn":
    attn(16, 10, 4096, 4096, 64)
    attn(16, 20, 1024, 1024, 64)
elif which == "attncross":
    attn(16, 20, 1024, 77, 64)
