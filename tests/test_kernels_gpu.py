"""GPU parity tests, kernel level: each C-ABI entry point against the oracle (oracle/ops.py, fp32) on
the same seeded fp16/bf16-rounded inputs.  Tolerances: outputs are fp16/bf16 with fp32 accumulation,
so the bound is one output rounding (2^-11 fp16, 2^-8 bf16 relative) plus accumulation-order noise.
"""
import math

import pytest
import torch

from oracle import ops as O
from tests.util import assert_close

pytestmark = pytest.mark.gpu

DEV = "cuda"


def _ops():
    from b200forge import ops
    return ops


def _rand(*shape, dtype=torch.float16, scale=1.0, seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(dtype).to(DEV)


def _tol(dtype):
    # max_rel (max-abs in units of the reference RMS) catches a single corrupted tile row that an RMS over the matrix hides
    return dict(rel_rms=2e-3, max_rel=1.5e-2) if dtype == torch.float16 else dict(rel_rms=1.2e-2, max_rel=1e-1)


# ------------------------------------------------------------------------------------------------ GEMM
@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (256, 320, 320), (1232, 1280, 2048), (16, 1280, 320),
                                   (4096, 640, 640), (300, 96, 200), (2048, 10240, 1280)])
@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_gemm_plain(M, N, K, dtype):
    ops = _ops()
    a = _rand(M, K, dtype=dtype, seed=1)
    w = _rand(N, K, dtype=dtype, scale=K ** -0.5, seed=2)
    b = _rand(N, dtype=dtype, seed=3)
    y = ops.gemm(a, w, b)
    torch.cuda.synchronize()
    ref = O.linear(a.float(), w.float(), b.float())
    assert_close(f"gemm {M}x{N}x{K} {dtype}", y, ref, **_tol(dtype))


@pytest.mark.parametrize("M,N,K,dtype", [(16384, 1280, 5120, torch.float16),   # 64 pair tiles x 5 = 320 on 74 pairs: tail 24, 3 shares
                                          (5632, 1280, 5120, torch.float16),    # 110 tiles: tail 36, 2 shares
                                          (1024, 1280, 2560, torch.bfloat16),   # 20 pair tiles: fewer tiles than units, one wave, not split
                                          (16, 1280, 4096, torch.float16),      # 5 single-CTA tiles: one wave, not split
                                          (16384, 1280, 1280, torch.float16)])  # same tail, too few k-chunks: not split
def test_gemm_k_split_tail(M, N, K, dtype):
    """Shapes whose tile count leaves a partly filled last wave: the tail tiles are K-split across the idle units (partial
    accumulators through a workspace, share 0 reduces them in a fixed order).  Results against the oracle, with the residual +
    row-statistics epilogue, and bit-identical over repeated launches (flags self-reset, fixed summation order)."""
    ops = _ops()
    a = _rand(M, K, dtype=dtype, seed=120)
    w = _rand(N, K, dtype=dtype, scale=K ** -0.5, seed=121)
    b = _rand(N, dtype=dtype, seed=122)
    res = _rand(M, N, dtype=dtype, seed=123)
    ref = O.linear(a.float(), w.float(), b.float())
    outs = []
    for _ in range(3):
        outs.append(ops.gemm(a, w, b).clone())
    torch.cuda.synchronize()
    assert_close(f"gemm k-split {M}x{N}x{K} {dtype}", outs[0], ref, **_tol(dtype))
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2]), "K-split GEMM must be bit-reproducible"
    if dtype == torch.float16 and N % 32 == 0:
        st = ops.row_stats_buffer(M, N, DEV)
        y = ops.gemm(a, w, b, residual=res, row_stats_out=st)
        torch.cuda.synchronize()
        assert_close("gemm k-split + residual + statistics", y, ref + res.float(), **_tol(dtype))
        cnt = st[..., 0].sum(0)
        mean = (st[..., 0] * st[..., 1]).sum(0) / cnt
        want = (ref + res.float()).mean(-1)
        assert (mean - want).abs().max().item() < 2e-3


def test_conv3x3_k_split_tail():
    """conv 16 x 32 x 32, 1280 -> 1280 (the SDXL 1280-channel level): 320 pair tiles on 74 pairs, 180 k-chunks in 3 shares."""
    ops = _ops()
    N, H, W, C = 16, 32, 32, 1280
    x = _rand(N, H, W, C, seed=124)
    w = _rand(C, C, 3, 3, scale=(9 * C) ** -0.5, seed=125)
    b = _rand(C, seed=126)
    temb = _rand(N, C, seed=127)
    res = _rand(N, H, W, C, seed=128)
    wp = ops.pack_conv3x3(w)
    y = ops.conv3x3(x, wp, b, temb=temb, residual=res).clone()
    y2 = ops.conv3x3(x, wp, b, temb=temb, residual=res)
    torch.cuda.synchronize()
    ref = (O.conv2d(x.float().permute(0, 3, 1, 2), w.float(), b.float()) + temb.float()[:, :, None, None]).permute(0, 2, 3, 1) + res.float()
    assert_close("conv3x3 k-split", y, ref, rel_rms=2e-3, max_rel=1.6e-2)
    assert torch.equal(y, y2)


def test_gemm_no_bias_strided_output():
    ops = _ops()
    M, N, K = 512, 640, 640
    a = _rand(M, K, seed=4)
    w = _rand(N, K, scale=K ** -0.5, seed=5)
    buf = torch.zeros(M, 3 * N, dtype=torch.float16, device=DEV)
    ops.gemm(a, w, out=buf[:, N:2 * N])
    torch.cuda.synchronize()
    assert_close("gemm strided out", buf[:, N:2 * N], O.linear(a.float(), w.float()), rel_rms=2e-3, max_rel=1.6e-02)
    assert buf[:, :N].abs().max().item() == 0 and buf[:, 2 * N:].abs().max().item() == 0


def test_gemm_residual_rowvec_silu():
    ops = _ops()
    M, N, K = 1024, 320, 1280
    a = _rand(M, K, seed=6)
    w = _rand(N, K, scale=K ** -0.5, seed=7)
    b = _rand(N, seed=8)
    res = _rand(M, N, seed=9)
    rv = _rand(4, N, seed=10)
    y = ops.gemm(a, w, b, residual=res, rowvec=rv, rows_per_vec=256, epilogue=ops.EPI_SILU)
    torch.cuda.synchronize()
    pre = O.linear(a.float(), w.float(), b.float()) + rv.float().repeat_interleave(256, dim=0)
    ref = O.silu(pre) + res.float()
    assert_close("gemm silu+rowvec+residual", y, ref, rel_rms=2e-3, max_rel=1.6e-02)


def test_gemm_bias_along_m():
    ops = _ops()
    M, N, K = 512, 1024, 512
    a = _rand(M, K, seed=11)
    w = _rand(N, K, scale=K ** -0.5, seed=12)
    b = _rand(M, seed=13)
    y = ops.gemm(a, w, b, bias_along_m=True)
    torch.cuda.synchronize()
    assert_close("gemm bias_m", y, O.linear(a.float(), w.float()) + b.float()[:, None], rel_rms=2e-3, max_rel=1.6e-02)


@pytest.mark.parametrize("C", [640, 1280])
def test_gemm_geglu(C):
    ops = _ops()
    M = 1024
    a = _rand(M, C, seed=14)
    w = _rand(8 * C, C, scale=C ** -0.5, seed=15)
    b = _rand(8 * C, seed=16, scale=0.1)
    wp, bp = ops.pack_geglu(w, b, 256)
    y = ops.gemm(a, wp, bp, epilogue=ops.EPI_GEGLU, block_n=256)
    torch.cuda.synchronize()
    assert_close(f"geglu C={C}", y, O.geglu(a.float(), w.float(), b.float()), rel_rms=3e-3, max_rel=2.4e-02)


def test_gemm_concat_sources():
    ops = _ops()
    M, K1, K2, N = 1024, 640, 320, 640
    a1 = _rand(M, K1, seed=17)
    a2 = _rand(M, K2, seed=18)
    w = _rand(N, K1 + K2, scale=(K1 + K2) ** -0.5, seed=19)
    y = ops.gemm(a1, w, a2=a2)
    torch.cuda.synchronize()
    assert_close("gemm concat", y, O.linear(torch.cat([a1, a2], 1).float(), w.float()), rel_rms=2e-3, max_rel=1.6e-02)


# ------------------------------------------------------------------------------------------------ conv
@pytest.mark.parametrize("N,H,W,C1,C2,Cout", [(2, 32, 32, 64, 0, 64), (2, 128, 128, 320, 0, 320),
                                              (3, 64, 64, 640, 320, 640), (2, 32, 32, 1280, 1280, 1280),
                                              (4, 8, 8, 1280, 0, 1280), (2, 16, 16, 128, 64, 320)])
def test_conv3x3(N, H, W, C1, C2, Cout):
    ops = _ops()
    C = C1 + C2
    x = _rand(N, C, H, W, seed=20)
    w = _rand(Cout, C, 3, 3, scale=(9 * C) ** -0.5, seed=21)
    b = _rand(Cout, seed=22)
    xn = x.permute(0, 2, 3, 1).contiguous()
    x1 = xn[..., :C1].contiguous()
    x2 = xn[..., C1:].contiguous() if C2 else None
    y = ops.conv3x3(x1, ops.pack_conv3x3(w), b, x2=x2)
    torch.cuda.synchronize()
    ref = O.conv2d(x.float(), w.float(), b.float()).permute(0, 2, 3, 1)
    assert_close(f"conv3x3 {N}x{H}x{W} {C1}+{C2}->{Cout}", y, ref, rel_rms=2e-3, max_rel=1.6e-02)


def test_conv3x3_temb_residual():
    ops = _ops()
    N, H, W, C, Cout = 2, 64, 64, 320, 320
    x = _rand(N, C, H, W, seed=23)
    w = _rand(Cout, C, 3, 3, scale=(9 * C) ** -0.5, seed=24)
    b = _rand(Cout, seed=25)
    temb = _rand(N, Cout, seed=26)
    res = _rand(N, H, W, Cout, seed=27)
    y = ops.conv3x3(x.permute(0, 2, 3, 1).contiguous(), ops.pack_conv3x3(w), b, temb=temb, residual=res)
    torch.cuda.synchronize()
    ref = O.conv2d(x.float(), w.float(), b.float()) + temb.float()[:, :, None, None]
    ref = ref.permute(0, 2, 3, 1) + res.float()
    assert_close("conv3x3 temb+residual", y, ref, rel_rms=2e-3, max_rel=1.6e-02)


def test_conv_via_im2col_stride2_and_small_c():
    ops = _ops()
    # stride-2 downsample (backend/nn/unet.py:358-374)
    N, H, W, C, Cout = 2, 64, 64, 320, 320
    x = _rand(N, C, H, W, seed=28)
    w = _rand(Cout, C, 3, 3, scale=(9 * C) ** -0.5, seed=29)
    b = _rand(Cout, seed=30)
    cols = ops.im2col3x3(x.permute(0, 2, 3, 1).contiguous(), stride=2)
    y = ops.gemm(cols, ops.pack_conv3x3(w), b).view(N, H // 2, W // 2, Cout)
    torch.cuda.synchronize()
    ref = O.conv2d(x.float(), w.float(), b.float(), stride=2).permute(0, 2, 3, 1)
    assert_close("conv s2 via im2col", y, ref, rel_rms=2e-3, max_rel=1.6e-02)
    # 4-channel scalar path
    x4 = _rand(2, 4, 32, 32, seed=31)
    cols4 = ops.im2col3x3(x4.permute(0, 2, 3, 1).contiguous(), ldo=64)
    ref4 = torch.nn.functional.unfold(x4.float(), 3, padding=1)  # [n, c*9, P] with k = c*9 + tap
    ref4 = ref4.view(2, 4, 9, -1).permute(0, 3, 2, 1).reshape(2 * 32 * 32, 36)
    torch.cuda.synchronize()
    assert torch.equal(cols4[:, :36].float().cpu(), ref4.cpu())
    assert cols4[:, 36:].abs().max().item() == 0


# ------------------------------------------------------------------------------------------------ attention
@pytest.mark.parametrize("B,H,Lq,Lk,Dh", [(2, 10, 4096, 4096, 64), (2, 20, 1024, 1024, 64), (2, 20, 1024, 77, 64),
                                          (1, 10, 4096, 77, 64), (2, 4, 200, 333, 64), (1, 24, 1152, 1152, 128),
                                          (2, 3, 130, 77, 128)])
@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_attention(B, H, Lq, Lk, Dh, dtype):
    ops = _ops()
    q = _rand(B, Lq, H * Dh, dtype=dtype, seed=32)
    k = _rand(B, Lk, H * Dh, dtype=dtype, seed=33)
    v = _rand(B, Lk, H * Dh, dtype=dtype, seed=34)
    o = ops.attention(q, k, v, H)
    torch.cuda.synchronize()
    ref = O.attention(q.float(), k.float(), v.float(), H)
    tol = dict(rel_rms=3e-3, max_rel=2e-2) if dtype == torch.float16 else dict(rel_rms=1.5e-2, max_rel=1.2e-1)
    assert_close(f"attention B{B} H{H} {Lq}x{Lk} d{Dh} {dtype}", o, ref, **tol)


def test_attention_fused_qkv_views():
    ops = _ops()
    B, L, H, Dh = 2, 1024, 10, 64
    C = H * Dh
    qkv = _rand(B, L, 3 * C, seed=35)
    o = ops.attention(qkv[:, :, :C], qkv[:, :, C:2 * C], qkv[:, :, 2 * C:], H)
    torch.cuda.synchronize()
    ref = O.attention(qkv[:, :, :C].float(), qkv[:, :, C:2 * C].float(), qkv[:, :, 2 * C:].float(), H)
    assert_close("attention fused qkv", o, ref, rel_rms=3e-3, max_rel=2.4e-02)


def test_attention_peaked_logits():
    # large-magnitude logits exercise the online-softmax rescaling path
    ops = _ops()
    B, L, H, Dh = 1, 512, 2, 64
    q = _rand(B, L, H * Dh, seed=36, scale=4.0)
    k = _rand(B, L, H * Dh, seed=37, scale=4.0)
    v = _rand(B, L, H * Dh, seed=38)
    o = ops.attention(q, k, v, H)
    torch.cuda.synchronize()
    assert_close("attention peaked", o, O.attention(q.float(), k.float(), v.float(), H), rel_rms=5e-3, max_rel=4.0e-02)


@pytest.mark.parametrize("B,H,Lq,Lk", [(1, 24, 4352, 4352), (2, 4, 200, 333), (1, 3, 1000, 129), (2, 2, 64, 4096)])
def test_attention_dh128_flux_shapes(B, H, Lq, Lk):
    """Dh = 128 (Flux / SD3) through the small-CTA kernel (attention64s.cu built for Dh = 128, two CTAs per SM): the benchmark's
    4096 + 256 tokens, ragged query / key counts (last key block partly empty), q / k / v read as column slices of one fused
    QKV buffer like flux_engine does, and large logits that force the lazy O rescale."""
    ops = _ops()
    Dh = 128
    C = H * Dh
    L = max(Lq, Lk)
    qkv = _rand(B, L, 3 * C, dtype=torch.bfloat16, seed=110)
    q, k, v = qkv[:, :Lq, :C], qkv[:, :Lk, C:2 * C], qkv[:, :Lk, 2 * C:]
    o = ops.attention(q, k, v, H)
    torch.cuda.synchronize()
    assert_close(f"attention d128 B{B} H{H} {Lq}x{Lk}", o, O.attention(q.float(), k.float(), v.float(), H), rel_rms=1.5e-2, max_rel=1.2e-1)
    if Lq <= 1000:
        q4 = (q.float() * 3.0).half().contiguous()
        k4 = (k.float() * 3.0).half().contiguous()
        o4 = ops.attention(q4, k4, v.half().contiguous(), H)
        torch.cuda.synchronize()
        assert_close("attention d128 peaked fp16", o4, O.attention(q4.float(), k4.float(), v.half().float(), H), rel_rms=5e-3, max_rel=4e-2)


# ------------------------------------------------------------------------------------------------ norms
@pytest.mark.parametrize("N,H,W,C1,C2", [(2, 32, 32, 320, 0), (2, 64, 64, 640, 320), (3, 16, 16, 1280, 1280),
                                         (2, 128, 128, 320, 0), (1, 8, 8, 2560, 0)])
@pytest.mark.parametrize("silu", [False, True])
def test_groupnorm(N, H, W, C1, C2, silu):
    ops = _ops()
    C = C1 + C2
    x = _rand(N, C, H, W, seed=39) * 2 + 0.5
    g = _rand(C, seed=40) * 0.2 + 1
    b = _rand(C, seed=41) * 0.2
    xn = x.permute(0, 2, 3, 1).contiguous()
    x1 = xn[..., :C1].contiguous()
    x2 = xn[..., C1:].contiguous() if C2 else None
    y = ops.groupnorm(x1, g, b, eps=1e-5, silu=silu, x2=x2)
    torch.cuda.synchronize()
    ref = O.group_norm(x.float(), 32, g.float(), b.float(), 1e-5)
    if silu:
        ref = O.silu(ref)
    assert_close(f"groupnorm {N}x{H}x{W} {C1}+{C2} silu={silu}", y, ref.permute(0, 2, 3, 1), max_abs=2e-2, rel_rms=1.5e-3)


@pytest.mark.parametrize("rows,C", [(4096, 640), (1000, 1280), (77, 320), (512, 3072)])
def test_layernorm(rows, C):
    ops = _ops()
    x = _rand(rows, C, seed=42) * 3 + 1
    g = _rand(C, seed=43) * 0.2 + 1
    b = _rand(C, seed=44) * 0.2
    y = ops.layernorm(x, g, b, 1e-5)
    y2 = ops.layernorm(x, None, None, 1e-6)
    torch.cuda.synchronize()
    assert_close(f"layernorm {rows}x{C}", y, O.layer_norm(x.float(), g.float(), b.float(), 1e-5), rel_rms=1.5e-3, max_rel=1.5e-02)
    assert_close(f"layernorm noaffine {rows}x{C}", y2, O.layer_norm(x.float(), None, None, 1e-6), rel_rms=1.5e-3, max_rel=1.5e-02)


# ------------------------------------------------------------------------------------------------ helpers
def test_layout_upsample_silu_temb():
    ops = _ops()
    x = _rand(2, 64, 16, 16, seed=45)
    xn = ops.nchw_to_nhwc(x, torch.float16)
    assert torch.equal(xn, x.permute(0, 2, 3, 1).contiguous())
    xf = torch.randn(2, 4, 16, 16, device=DEV)
    xh = ops.nchw_to_nhwc(xf, torch.float16)
    assert torch.equal(xh, xf.permute(0, 2, 3, 1).contiguous().half())
    back = ops.nhwc_to_nchw(xn, out_dtype=torch.float32)
    assert torch.equal(back, x.float())
    part = ops.nhwc_to_nchw(xn, channels=4)
    assert torch.equal(part, x[:, :4])
    up = ops.upsample2x(xn)
    assert torch.equal(up, O.upsample_nearest2x(x).permute(0, 2, 3, 1).contiguous())
    s = ops.silu(x)
    assert_close("silu", s, O.silu(x.float()), rel_rms=1e-3, max_rel=1.5e-02)
    t = torch.tensor([0.0, 1.0, 500.0, 999.0], device=DEV)
    e = ops.timestep_embedding(t, 320, torch.float32 if False else torch.float16)
    ref = O.timestep_embedding(t.cpu(), 320)
    torch.cuda.synchronize()
    assert_close("timestep_embedding", e, ref, max_abs=2e-3)


def test_unet_input_im2col():
    ops = _ops()
    B, C, H, W = 2, 4, 16, 16
    x = torch.randn(B, C, H, W, device=DEV)
    sigma = torch.tensor([14.6, 0.5], device=DEV)
    cols = ops.unet_input_im2col(x, sigma, torch.float16, reps=2, ldo=64)
    torch.cuda.synchronize()
    xc = (x / (sigma.view(-1, 1, 1, 1) ** 2 + 1.0) ** 0.5).half()
    ref = torch.nn.functional.unfold(xc.float(), 3, padding=1).view(B, C, 9, -1).permute(0, 3, 2, 1).reshape(B * H * W, 36)
    assert cols.shape == (2 * B * H * W, 64)
    assert_close("unet_input_im2col", cols[:B * H * W, :36], ref, max_abs=1e-3)
    assert torch.equal(cols[:B * H * W], cols[B * H * W:])
    assert cols[:, 36:].abs().max().item() == 0


def test_softmax_rows():
    ops = _ops()
    x = _rand(64, 4096, dtype=torch.bfloat16, seed=46) * 3
    ref = torch.softmax(x.float() * 0.125, dim=-1)
    ops.softmax_rows_(x, 0.125)
    torch.cuda.synchronize()
    assert_close("softmax_rows", x, ref, rel_rms=1e-2, max_rel=8.0e-02)


# ------------------------------------------------------------------------------------------------ sampler step
def test_sampler_step_euler_ancestral_and_dpmpp():
    ops = _ops()
    from oracle import sampling as S
    B, C, H, W = 2, 4, 16, 16
    g = torch.Generator().manual_seed(47)
    x = torch.randn(B, C, H, W, generator=g)
    eps = torch.randn(2 * B, H, W, 8, generator=g).half()
    noise = torch.randn(B, C, H, W, generator=g)
    sigma, sigma_next, cfg = 7.5, 5.0, 7.0
    eps_nchw = eps[..., :C].permute(0, 3, 1, 2).float()
    den_ref = S.cfg_denoised_eps(x, eps_nchw[:B], eps_nchw[B:], sigma, cfg)
    x_ref = S.euler_ancestral_step(x, den_ref, sigma, sigma_next, noise, eta=1.0, s_noise=1.0)
    sd, su = S.get_ancestral_step(sigma, sigma_next, 1.0)
    xd = x.to(DEV).clone()
    den = torch.empty_like(xd)
    ops.sampler_step(xd, eps.to(DEV), den, kind=ops.STEP_EULER, sigma=sigma, cfg_scale=cfg, has_uncond=True,
                     dt=sd - sigma, noise=noise.to(DEV), noise_scale=su)
    torch.cuda.synchronize()
    assert_close("sampler denoised", den, den_ref, max_abs=2e-5)
    assert_close("sampler euler-a x", xd, x_ref, max_abs=2e-5)
    # dpm++ 2m second-order step
    old = torch.randn(B, C, H, W, generator=g)
    sig_prev = 9.0
    x_ref2 = S.dpmpp_2m_step(x, den_ref, old, sig_prev, sigma, sigma_next)
    cx, cd, cold = S.dpmpp_2m_coeffs(sig_prev, sigma, sigma_next, has_old=True)
    xd = x.to(DEV).clone()
    oldd = old.to(DEV).clone()
    ops.sampler_step(xd, eps.to(DEV), den, kind=ops.STEP_DPMPP_2M, sigma=sigma, cfg_scale=cfg, has_uncond=True,
                     old_denoised=oldd, c_x=cx, c_d=cd, c_old=cold)
    torch.cuda.synchronize()
    assert_close("sampler dpmpp2m x", xd, x_ref2, max_abs=5e-5)
    assert_close("sampler dpmpp2m old", oldd, den_ref, max_abs=2e-5)


def _merge_partials(st):
    """(count, mean, M2) partials [P, M, 4] -> (mean, biased variance) per row, in fp64 on the host."""
    st = st.double().cpu()
    cnt = st[:, :, 0].sum(0)
    mean = (st[:, :, 0] * st[:, :, 1]).sum(0) / cnt
    m2 = (st[:, :, 2] + st[:, :, 0] * (st[:, :, 1] - mean[None, :]) ** 2).sum(0)
    return cnt, mean, m2 / cnt


def test_gemm_layernorm_fold_and_row_stats():
    """LayerNorm -> Linear as one GEMM on the raw rows (partial statistics from the producer's epilogue) vs the oracle's
    layer_norm + linear; also the GEGLU variant used by the transformer feed-forward."""
    ops = _ops()
    M, C, N = 1024, 640, 1280
    x_in = _rand(M, C, seed=50)
    w0 = _rand(C, C, scale=C ** -0.5, seed=51)
    res = _rand(M, C, seed=52) * 2 + 0.7
    # producer: t = x_in @ w0^T + res, with row statistics of t taken in its epilogue
    stats = ops.row_stats_buffer(M, C, DEV)
    stats.fill_(float("nan"))  # every partial must be written by the kernel
    t = ops.gemm(x_in, w0, residual=res, row_stats_out=stats)
    torch.cuda.synchronize()
    tf = t.float()
    cnt, mean, var = _merge_partials(stats)
    assert torch.all(cnt == C)
    # the statistics are taken from the fp32 values before they are rounded to fp16 for the store
    assert_close("row stats mean", mean.float(), tf.mean(1).cpu(), max_abs=2e-3, rel_rms=1e-3)
    assert_close("row stats var", var.float(), tf.var(1, unbiased=False).cpu(), rel_rms=1e-3, max_rel=1.5e-02)
    gamma = (_rand(C, seed=53) * 0.2 + 1)
    beta = _rand(C, seed=54) * 0.2
    w1 = _rand(N, C, scale=C ** -0.5, seed=55)
    b1 = _rand(N, seed=56)
    wf, c, d = ops.fold_layernorm(w1, b1, gamma, beta)
    y = ops.gemm(t, wf, None, ln=(stats, c, d, 1e-5))
    torch.cuda.synchronize()
    ref = O.linear(O.layer_norm(tf, gamma.float(), beta.float(), 1e-5), w1.float(), b1.float())
    assert_close("LN folded into GEMM", y, ref, max_abs=2e-2, rel_rms=2e-3)
    # GEGLU consumer
    w2 = _rand(8 * C, C, scale=C ** -0.5, seed=57)
    b2 = _rand(8 * C, seed=58, scale=0.1)
    wf2, c2, d2 = ops.fold_layernorm(w2, b2, gamma, beta)
    wp, cp = ops.pack_geglu(wf2, c2, 256)
    _, dp = ops.pack_geglu(wf2, d2, 256)
    g = ops.gemm(t, wp, None, epilogue=ops.EPI_GEGLU, block_n=256, ln=(stats, cp, dp, 1e-5))
    torch.cuda.synchronize()
    refg = O.geglu(O.layer_norm(tf, gamma.float(), beta.float(), 1e-5), w2.float(), b2.float())
    assert_close("LN folded into GEGLU GEMM", g, refg, max_abs=3e-2, rel_rms=3e-3)
    # bit-reproducible: no atomics anywhere on the statistics path
    stats2 = ops.row_stats_buffer(M, C, DEV)
    t2 = ops.gemm(x_in, w0, residual=res, row_stats_out=stats2)
    y2 = ops.gemm(t2, wf, None, ln=(stats2, c, d, 1e-5))
    torch.cuda.synchronize()
    assert torch.equal(stats, stats2) and torch.equal(y, y2)


@pytest.mark.parametrize("C,mean,sigma", [(640, 100.0, 0.1), (1280, -300.0, 0.25), (320, 30.0, 1.0)])
def test_layernorm_fold_large_row_mean(C, mean, sigma):
    """Rows with |mean| >> sigma (outlier channels of real checkpoints): sumsq/K - mean^2 in fp32 loses the variance
    entirely; the partial (count, mean, M2) statistics must not.  Reference: F.layer_norm (backend/nn/unet.py:171-175)."""
    ops = _ops()
    M, N = 512, 640
    g = torch.Generator().manual_seed(60)
    t_ref = (torch.randn(M, C, generator=g) * sigma + mean + torch.randn(M, 1, generator=g) * sigma * 3)
    # producer with identity-free setup: t = 0 @ w + residual, so t == residual exactly (fp16 grid)
    res = t_ref.half().to(DEV)
    zero_a = torch.zeros(M, 64, dtype=torch.float16, device=DEV)
    w0 = torch.zeros(C, 64, dtype=torch.float16, device=DEV)
    stats = ops.row_stats_buffer(M, C, DEV)
    t = ops.gemm(zero_a, w0, residual=res, row_stats_out=stats)
    torch.cuda.synchronize()
    assert torch.equal(t, res)
    cnt, mu, var = _merge_partials(stats)
    tf = res.double().cpu()
    assert_close("large-mean row mean", mu.float(), tf.mean(1).float(), max_abs=abs(mean) * 2e-6)
    assert_close("large-mean row var", var.float(), tf.var(1, unbiased=False).float(), rel_rms=1e-3, max_rel=1.5e-02)
    gamma = (_rand(C, seed=61) * 0.2 + 1)
    beta = _rand(C, seed=62) * 0.2
    w1 = _rand(N, C, scale=C ** -0.5, seed=63)
    b1 = _rand(N, seed=64)
    wf, c, d = ops.fold_layernorm(w1, b1, gamma, beta)
    y = ops.gemm(t, wf, None, ln=(stats, c, d, 1e-5))
    torch.cuda.synchronize()
    ref = O.linear(O.layer_norm(res.float(), gamma.float(), beta.float(), 1e-5), w1.float(), b1.float())
    # acc - mean*c cancels |mean|/sigma digits of the fp32 accumulator: the bound scales with that ratio
    assert_close(f"LN fold, row mean {mean} sigma {sigma}", y, ref, rel_rms=max(3e-3, 5e-7 * abs(mean) / sigma * C ** 0.5))


@pytest.mark.parametrize("N,H,W,C,mean,sigma", [(2, 32, 32, 320, 50.0, 0.1), (1, 64, 64, 128, -200.0, 0.5), (2, 16, 16, 1280, 8.0, 0.02)])
def test_groupnorm_large_group_mean(N, H, W, C, mean, sigma):
    """GroupNorm with |group mean| >> sigma, against F.group_norm in fp64 on the same fp16 input
    (backend/nn/unet.py:395, backend/nn/vae.py:12)."""
    ops = _ops()
    g = torch.Generator().manual_seed(65)
    x = (torch.randn(N, H, W, C, generator=g) * sigma + mean).half().to(DEV)
    gam = (_rand(C, seed=66) * 0.2 + 1)
    bet = _rand(C, seed=67) * 0.2
    y = ops.groupnorm(x, gam, bet, eps=1e-5, silu=False)
    y2 = ops.groupnorm(x, gam, bet, eps=1e-5, silu=False)
    torch.cuda.synchronize()
    ref = torch.nn.functional.group_norm(x.double().permute(0, 3, 1, 2), 32, gam.double(), bet.double(), 1e-5).permute(0, 2, 3, 1)
    assert_close(f"groupnorm mean {mean} sigma {sigma}", y, ref.float(), max_abs=2e-2, rel_rms=2e-3)
    assert torch.equal(y, y2), "GroupNorm must be bit-reproducible"


def test_attention_blockdiag_vs_oracle():
    """Head dim 160 (SD1.5, 1280 channels / 8 heads): per head one S = Q_h K_h^T over the whole batch, block-diagonal
    softmax, one P V_h — against the oracle's per-sample attention; self (Lk = L) and cross (77 keys) shapes."""
    ops = _ops()
    B, H, Dh = 16, 8, 160
    for L, Lk, seed in ((256, 256, 70), (64, 77, 71)):
        q = _rand(B, L, H * Dh, seed=seed)
        k = _rand(B, Lk, H * Dh, seed=seed + 10)
        v = _rand(B, Lk, H * Dh, seed=seed + 20)
        out = ops.attention_blockdiag(q, k, v, H, scale=Dh ** -0.5)
        torch.cuda.synchronize()
        ref = O.attention(q.float(), k.float(), v.float(), H)
        assert_close(f"attention_blockdiag L={L} Lk={Lk}", out, ref, rel_rms=3e-3, max_rel=2.4e-02)


@pytest.mark.parametrize("N,H,W,C1,C2,Cout", [(2, 104, 152, 320, 0, 320), (2, 52, 76, 640, 640, 640), (3, 26, 38, 1280, 0, 1280),
                                              (2, 13, 19, 1280, 0, 1280), (1, 96, 168, 64, 64, 128)])
def test_conv3x3_generic_tiling(N, H, W, C1, C2, Cout):
    """Image widths that are neither a power of two nor a multiple of 128 (SDXL's non-square buckets: 1216x832 ->
    152x104 latents and their /2, /4 levels; odd sizes): tiles overhang the image, loads zero-fill, stores are masked.
    Includes the time-embedding row and the residual, which are addressed through the decoded pixel index."""
    ops = _ops()
    C = C1 + C2
    x = _rand(N, C, H, W, seed=80)
    w = _rand(Cout, C, 3, 3, scale=(9 * C) ** -0.5, seed=81)
    b = _rand(Cout, seed=82)
    temb = _rand(N, Cout, seed=83)
    res = _rand(N, H, W, Cout, seed=84)
    xn = x.permute(0, 2, 3, 1).contiguous()
    x1 = xn[..., :C1].contiguous()
    x2 = xn[..., C1:].contiguous() if C2 else None
    guard = torch.full((N, H, W, Cout), 7.0, device=DEV, dtype=torch.float16)
    y = ops.conv3x3(x1, ops.pack_conv3x3(w), b, x2=x2, temb=temb, residual=res, out=guard)
    torch.cuda.synchronize()
    ref = (O.conv2d(x.float(), w.float(), b.float()) + temb.float()[:, :, None, None]).permute(0, 2, 3, 1) + res.float()
    assert_close(f"conv3x3 generic {N}x{H}x{W} {C1}+{C2}->{Cout}", y, ref, max_abs=2.5e-2, rel_rms=2e-3)


@pytest.mark.parametrize("N,H,W,C,Cout,dtype", [
    (2, 32, 32, 1280, 1280, torch.float16),   # SDXL 32 -> 64
    (2, 64, 64, 640, 640, torch.float16),     # SDXL 64 -> 128
    (3, 52, 76, 640, 640, torch.float16),     # non-square bucket, 57 tiles per parity (padded to 58 for the CTA pairs)
    (1, 13, 19, 1280, 1280, torch.float16),   # odd sizes: tile_w = 1
    (4, 4, 4, 256, 128, torch.float16),       # image smaller than one tile, Cout != C
    (1, 128, 128, 512, 512, torch.bfloat16),  # VAE 128 -> 256
    (1, 256, 256, 256, 256, torch.bfloat16),  # VAE 256 -> 512 (shape of the 512 -> 1024 level, quarter size)
])
def test_conv3x3_up2x(N, H, W, C, Cout, dtype):
    """Nearest x2 upsample folded into the 3x3 convolution (b200_conv3x3_up2x: four 2x2 parity filters on the low-res
    image) against upsample + conv in fp32 (backend/nn/unet.py:330-355, backend/nn/vae.py:38-58).  The pre-summed weights
    are rounded once to the operand type, so the bound is the plain convolution's plus one weight rounding."""
    import torch.nn.functional as F
    ops = _ops()
    x = _rand(N, C, H, W, dtype=dtype, seed=100)
    w = _rand(Cout, C, 3, 3, dtype=dtype, scale=(9 * C) ** -0.5, seed=101)
    b = _rand(Cout, dtype=dtype, seed=102)
    guard = torch.full((N, 2 * H, 2 * W, Cout), 7.0, device=DEV, dtype=dtype)
    y = ops.conv3x3_up2x(x.permute(0, 2, 3, 1).contiguous(), ops.pack_conv3x3_up2x(w), b, out=guard)
    torch.cuda.synchronize()
    ref = O.conv2d(F.interpolate(x.float(), scale_factor=2, mode="nearest"), w.float(), b.float()).permute(0, 2, 3, 1)
    assert_close(f"conv3x3_up2x {N}x{H}x{W} {C}->{Cout} {dtype}", y, ref, **_tol(dtype))
    # the unfolded route gives the same image within the same bound
    y2 = ops.conv3x3_any(ops.upsample2x(x.permute(0, 2, 3, 1).contiguous()), ops.pack_conv3x3(w), b)
    torch.cuda.synchronize()
    assert_close("upsample2x + conv3x3 (unfolded route)", y2, ref, **_tol(dtype))


def test_sampler_step_and_denoised_v_prediction():
    """prediction = 1 (SD2.x 768-v): denoised = x / (sigma^2 + 1) - v * sigma / sqrt(sigma^2 + 1)
    (backend/modules/k_prediction.py:81-92, sigma_data = 1), inside the fused CFG + Euler step and in b200_eps_to_denoised."""
    ops = _ops()
    from oracle import sampling as S
    B, C, H, W = 2, 4, 16, 16
    g = torch.Generator().manual_seed(90)
    x = torch.randn(B, C, H, W, generator=g) * 3
    v = torch.randn(2 * B, H, W, 8, generator=g).half()
    sigma, sigma_next, cfg = 3.5, 2.0, 6.0
    pred = S.VPrediction()
    vn = v[..., :C].permute(0, 3, 1, 2).float()
    sg = torch.full((B,), sigma)
    du, dc = pred.calculate_denoised(sg, vn[:B], x), pred.calculate_denoised(sg, vn[B:], x)
    den_ref = du + (dc - du) * cfg
    x_ref = S.euler_step(x, den_ref, sigma, sigma_next)
    xd = x.to(DEV).clone()
    den = torch.empty_like(xd)
    ops.sampler_step(xd, v.to(DEV), den, kind=ops.STEP_EULER, sigma=sigma, cfg_scale=cfg, has_uncond=True,
                     dt=sigma_next - sigma, prediction=1)
    torch.cuda.synchronize()
    assert_close("v-pred sampler denoised", den, den_ref, max_abs=3e-5)
    assert_close("v-pred sampler euler x", xd, x_ref, max_abs=3e-5)
    sig4 = torch.tensor([3.5, 3.5, 0.7, 14.0])
    x4 = torch.cat([x, x])
    out = ops.eps_to_denoised(x4.to(DEV), v.to(DEV), sig4.to(DEV), prediction=1)
    torch.cuda.synchronize()
    assert_close("v-pred eps_to_denoised", out, pred.calculate_denoised(sig4, vn, x4), max_abs=3e-5)


def test_any_size_convolution_route_and_ragged_attention():
    """What the im2col route (B200_CONV_ROUTE=im2col) needs on hardware: (1) 3x3 convolutions of non-tiling images as im2col + GEMM with the time-embedding
    row and the residual in the GEMM epilogue, (2) self-attention over token counts that are not multiples of the 128-key /
    256-query tiles (152x104 latents -> 15808 / 3952 / 988 tokens)."""
    ops = _ops()
    N, H, W, C, Cout = 2, 52, 76, 640, 640
    x = _rand(N, C, H, W, seed=100)
    w = _rand(Cout, C, 3, 3, scale=(9 * C) ** -0.5, seed=101)
    b = _rand(Cout, seed=102)
    temb = _rand(N, Cout, seed=103)
    res = _rand(N, H, W, Cout, seed=104)
    y = ops.conv3x3_any(x.permute(0, 2, 3, 1).contiguous(), ops.pack_conv3x3(w), b, temb=temb, residual=res, route="im2col")
    torch.cuda.synchronize()
    ref = (O.conv2d(x.float(), w.float(), b.float()) + temb.float()[:, :, None, None]).permute(0, 2, 3, 1) + res.float()
    assert_close("conv3x3_any 52x76", y, ref, rel_rms=2e-3, max_rel=1.6e-02)
    for (B, Hh, L, Dh) in ((2, 10, 3952, 64), (2, 20, 988, 64), (1, 4, 1000, 128)):
        q, k, v = (_rand(B, L, Hh * Dh, seed=110 + i) for i in range(3))
        out = ops.attention(q, k, v, Hh)
        torch.cuda.synchronize()
        assert_close(f"attention ragged L={L} Dh={Dh}", out, O.attention(q.float(), k.float(), v.float(), Hh), rel_rms=2e-3, max_rel=1.6e-02)
