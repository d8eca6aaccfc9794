"""GPU parity tests for the Flux (DiT) path: the new kernels against the oracle's functions,
and the whole transformer forward against (a) golden vectors made by the imported reference on CPU fp32
(tests/golden/flux_tiny*.pt, oracle/gen_golden.py) and (b) the oracle in fp32 on the GPU at Flux.1-dev width.

Stated tolerance (bf16 compute — the reference's dtype for Flux, fp32 accumulate): rel-RMS <= 3e-2 per forward against
the fp32 reference with O(1) activations; the test prints the oracle-in-bf16 distance for context.
"""
import os

import pytest
import torch

from oracle import flux as OF
from oracle import ops as O
from tests.util import assert_close, err_stats

pytestmark = pytest.mark.gpu
DEV = "cuda"
GOLD = os.path.join(os.path.dirname(__file__), "golden")
BF = torch.bfloat16


def _ops():
    from b200forge import ops
    return ops


def _rand(*shape, dtype=BF, scale=1.0, seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(dtype).to(DEV)


@pytest.mark.parametrize("dtype", [BF, torch.float16])
@pytest.mark.parametrize("C", [256, 3072])
def test_adaln(dtype, C):
    ops = _ops()
    B, L, Lt = 2, 384, 128
    x = _rand(B * L, C, dtype=dtype, seed=1) * 3 + 0.5
    mod = _rand(B, 4 * C, dtype=dtype, scale=0.3, seed=2)
    sh0, sc0, sh1, sc1 = (mod[:, i * C:(i + 1) * C] for i in range(4))
    y = ops.adaln(x, sh0, sc0, shift1=sh1, scale1=sc1, seg_period=L, seg_split=Lt)
    torch.cuda.synchronize()
    xf = O.layer_norm(x.float().view(B, L, C), None, None, 1e-6)
    ref = torch.cat([(1 + sc0.float()[:, None]) * xf[:, :Lt] + sh0.float()[:, None],
                     (1 + sc1.float()[:, None]) * xf[:, Lt:] + sh1.float()[:, None]], 1).view(B * L, C)
    assert_close(f"adaln C={C} {dtype}", y, ref, rel_rms=4e-3 if dtype == BF else 6e-4)
    y1 = ops.adaln(x, sh0, sc0)  # one parameter set
    torch.cuda.synchronize()
    ref1 = ((1 + sc0.float()[:, None]) * xf + sh0.float()[:, None]).view(B * L, C)
    assert_close(f"adaln single C={C} {dtype}", y1, ref1, rel_rms=4e-3 if dtype == BF else 6e-4)


def test_qk_norm_rope():
    """RMSNorm(q, k) * scale + RoPE in place on a fused QKV buffer vs oracle rms_norm + apply_rope (flux.py:128-139, 45-51)."""
    ops = _ops()
    from b200forge.flux_engine import rope_tables
    B, H, hh, ww, Lt = 2, 3, 8, 8, 32
    L = Lt + hh * ww
    extra = 64  # trailing columns (the mlp part of SingleStreamBlock.linear1) must stay untouched
    qkv = _rand(B * L, 3 * H * 128 + extra, seed=3)
    orig = qkv.clone()
    s = [(1 + 0.1 * _rand(128, seed=10 + i, dtype=torch.float32)).to(BF) for i in range(4)]
    cos, sin = rope_tables(hh, ww, Lt, [16, 56, 56], 10000, DEV)
    ocos, osin = OF.rope_tables(OF.position_ids(hh, ww, Lt), [16, 56, 56], 10000)
    assert torch.equal(cos.cpu(), ocos) and torch.equal(sin.cpu(), osin)
    ops.qk_norm_rope_(qkv, H, s[0], s[1], cos, sin, q_scale1=s[2], k_scale1=s[3], seg_split=Lt)
    torch.cuda.synchronize()
    v = orig[:, :3 * H * 128].float().view(B, L, 3, H, 128).permute(2, 0, 3, 1, 4)  # [3][B, H, L, D]
    for part, (s_txt, s_img) in enumerate(((s[0], s[2]), (s[1], s[3]))):
        t = torch.cat([OF.rms_norm(v[part][:, :, :Lt], s_txt.float()), OF.rms_norm(v[part][:, :, Lt:], s_img.float())], 2)
        ref = OF.apply_rope(t, cos, sin)
        got = qkv[:, part * H * 128:(part + 1) * H * 128].view(B, L, H, 128).permute(0, 2, 1, 3)
        assert_close(f"qk_norm_rope part {part}", got, ref, rel_rms=4e-3)
    assert torch.equal(qkv[:, 2 * H * 128:], orig[:, 2 * H * 128:])  # v and the trailing columns untouched


@pytest.mark.parametrize("in_f32", [True, False])
def test_patchify_roundtrip(in_f32):
    ops = _ops()
    B, C, H, W = 2, 16, 12, 20
    x = _rand(B, C, H, W, dtype=torch.float32 if in_f32 else BF, seed=4)
    tok = ops.flux_patchify(x, BF)
    torch.cuda.synchronize()
    ref = OF.patchify(x.float()).reshape(-1, 4 * C)
    assert torch.equal(tok.float(), ref.to(BF).float())
    back = ops.flux_unpatchify(tok, B, C, H, W, nchw_f32=True)
    nhwc = ops.flux_unpatchify(tok, B, C, H, W, nchw_f32=False)
    torch.cuda.synchronize()
    assert torch.equal(back, x.to(BF).float())
    assert torch.equal(nhwc.float(), x.to(BF).float().permute(0, 2, 3, 1))


def test_gemm_two_segments_gate_residual():
    """One GEMM over a joint [txt | img] activation with per-segment weights, modulation gate and in-place residual
    (DoubleStreamBlock, flux.py:252-258) vs per-stream oracle linears."""
    ops = _ops()
    B, Lt, Li, K, N = 2, 256, 512, 384, 512
    L = Lt + Li
    a = _rand(B * L, K, seed=5)
    wt, wi = _rand(N, K, scale=K ** -0.5, seed=6), _rand(N, K, scale=K ** -0.5, seed=7)
    bt, bi = _rand(N, seed=8, scale=0.1), _rand(N, seed=9, scale=0.1)
    gates = _rand(B, 2 * N, seed=10, scale=0.5)
    gt, gi = gates[:, :N], gates[:, N:]
    res = _rand(B * L, N, seed=11)
    out = res.clone()
    ops.gemm(a, wt, bt, rowvec=gt, rows_per_vec=L, rowvec_mul=True, residual=out, out=out, seg=(L, Lt, wi, bi, gi))
    torch.cuda.synchronize()
    af = a.float().view(B, L, K)
    ref = torch.cat([gt.float()[:, None] * O.linear(af[:, :Lt], wt.float(), bt.float()),
                     gi.float()[:, None] * O.linear(af[:, Lt:], wi.float(), bi.float())], 1).view(B * L, N) + res.float()
    assert_close("two-segment GEMM + gate + residual", out, ref, rel_rms=6e-3)
    # no-bias, activation variant (the QKV / MLP-in projections)
    y = ops.gemm(a, wt, None, epilogue=ops.EPI_GELU_TANH, seg=(L, Lt, wi, None, None))
    torch.cuda.synchronize()
    ref2 = OF.gelu_tanh(torch.cat([O.linear(af[:, :Lt], wt.float()), O.linear(af[:, Lt:], wi.float())], 1)).view(B * L, N)
    assert_close("two-segment GEMM + tanh GELU", y, ref2, rel_rms=6e-3)


def test_gemm_partial_activation_and_concat_gate():
    """SingleStreamBlock: linear1 with GELU on the mlp columns only, linear2 on [attn | gelu(mlp)] with gate + residual
    (flux.py:289-300)."""
    ops = _ops()
    B, L, hs, mlp = 2, 300, 256, 1024
    x = _rand(B * L, hs, seed=12)
    w1, b1 = _rand(3 * hs + mlp, hs, scale=hs ** -0.5, seed=13), _rand(3 * hs + mlp, seed=14, scale=0.1)
    y1 = ops.gemm(x, w1, b1, epilogue=ops.EPI_GELU_TANH, act_col0=3 * hs)
    torch.cuda.synchronize()
    lin = O.linear(x.float(), w1.float(), b1.float())
    ref1 = torch.cat([lin[:, :3 * hs], OF.gelu_tanh(lin[:, 3 * hs:])], 1)
    assert_close("linear1 partial GELU", y1, ref1, rel_rms=6e-3)
    attn = _rand(B * L, hs, seed=15)
    w2, b2 = _rand(hs, hs + mlp, scale=(hs + mlp) ** -0.5, seed=16), _rand(hs, seed=17, scale=0.1)
    gate = _rand(B, hs, seed=18, scale=0.5)
    xres = x.clone()
    ops.gemm(attn, w2, b2, a2=y1[:, 3 * hs:], rowvec=gate, rows_per_vec=L, rowvec_mul=True, residual=xres, out=xres)
    torch.cuda.synchronize()
    ref2 = x.float() + (gate.float()[:, None] * O.linear(torch.cat([attn.float(), y1[:, 3 * hs:].float()], 1), w2.float(), b2.float())
                        .view(B, L, hs)).view(B * L, hs)
    assert_close("linear2 concat + gate + residual", xres, ref2, rel_rms=6e-3)


def _engine(cfg, sd):
    from b200forge.flux_engine import FluxEngine
    return FluxEngine(cfg, sd, dtype=BF, device=DEV)


@pytest.mark.parametrize("fname", ["flux_tiny.pt", "flux_tiny_seg.pt", "flux_tiny_odd.pt"])
def test_flux_forward_vs_reference_golden(fname):
    """flux_tiny: 64 img + 128 txt tokens (per-stream launches); flux_tiny_seg: 256 + 256 tokens (two-segment GEMMs);
    flux_tiny_odd: 15 x 18 latent (circular pad to the patch size + crop, flux.py:394-397, 412)."""
    g = torch.load(os.path.join(GOLD, fname), weights_only=False)
    cfg = OF.CONFIGS[g["config"]]
    sd = OF.random_state_dict(cfg, seed=g["weight_seed"])
    eng = _engine(cfg, sd)
    out = eng.forward(g["x"].to(DEV), g["t"].to(DEV), g["context"].to(DEV), g["y"].to(DEV), g["guidance"].to(DEV))
    torch.cuda.synchronize()
    assert_close(f"flux {fname} bf16 engine vs reference fp32 golden", out, g["out"], rel_rms=3e-2, max_abs=4e-1)
    sd16 = {k: v.to(DEV).to(BF) for k, v in sd.items()}
    with torch.no_grad():
        ref16 = OF.flux_forward(sd16, cfg, g["x"].to(DEV).to(BF), g["t"].to(DEV), g["context"].to(DEV).to(BF), g["y"].to(DEV).to(BF),
                                g["guidance"].to(DEV))
    m, r = err_stats(ref16, g["out"])
    print(f"[parity] oracle-in-bf16 vs fp32 golden ({fname}): max_abs={m:.3e} rel_rms={r:.3e}")
    if "odd" in fname:  # 15 x 18 latent: the circular-pad / crop branch lives in forward() (NCHW out), like the reference's
        return
    # channels-last output for the fused sampler step is the same tensor, permuted
    nhwc = eng.forward_nhwc(g["x"].to(DEV), g["t"].to(DEV), g["context"].to(DEV).to(BF), g["y"].to(DEV).to(BF), g["guidance"].to(DEV))
    torch.cuda.synchronize()
    assert torch.equal(nhwc.float().permute(0, 3, 1, 2), out)


def test_flux_dev_width_vs_oracle_fp32():
    """Flux.1-dev width (hidden 3072, 24 heads, mlp 12288, T5 width 4096) with one double and one single block, 256 txt +
    256 img tokens, against the oracle in fp32 on the GPU with the same bf16-rounded weights."""
    torch.backends.cuda.matmul.allow_tf32 = False
    cfg = dict(OF.FLUX_DEV, depth=1, depth_single_blocks=1)
    sd = {k: v.to(BF) for k, v in OF.random_state_dict(cfg, seed=21).items()}
    eng = _engine(cfg, sd)
    g = torch.Generator().manual_seed(22)
    B, hw, Lt = 2, 32, 256
    x = torch.randn(B, 16, hw, hw, generator=g).to(DEV)
    ctx = torch.randn(B, Lt, cfg["context_in_dim"], generator=g).to(BF).to(DEV)
    y = torch.randn(B, cfg["vec_in_dim"], generator=g).to(BF).to(DEV)
    t = torch.tensor([0.8, 0.3], device=DEV)
    gd = torch.tensor([4.0, 4.0], device=DEV)
    out = eng.forward(x, t, ctx, y, gd)
    torch.cuda.synchronize()
    sd32 = {k: v.float().to(DEV) for k, v in sd.items()}
    with torch.no_grad():
        ref = OF.flux_forward(sd32, cfg, x.to(BF).float(), t, ctx.float(), y.float(), gd)
    assert_close("flux-dev width (1+1 blocks) bf16 engine vs oracle fp32", out, ref, rel_rms=3e-2)


@pytest.mark.parametrize("use_graph", [False, True])
def test_flux_trajectory_vs_oracle(use_graph):
    """4 Euler steps over the Simple schedule ('const' prediction, CFG 1, distilled guidance 4.0) through the public
    pipeline (CUDA graph + fused sampler step) vs the oracle loop in fp32 on the GPU."""
    from b200forge.pipeline import FluxTxt2ImgPipeline
    from oracle import sampling as OS
    torch.backends.cuda.matmul.allow_tf32 = False
    cfg = OF.TINY_FLUX
    sd = OF.random_state_dict(cfg, seed=31)
    pipe = FluxTxt2ImgPipeline(cfg, sd, device=DEV, use_graph=use_graph)
    g = torch.Generator().manual_seed(32)
    B, hw, Lt, steps = 2, 32, 256, 4
    noise = torch.randn(B, 16, hw, hw, generator=g)
    cond = dict(crossattn=torch.randn(B, Lt, cfg["context_in_dim"], generator=g), vector=torch.randn(B, cfg["vec_in_dim"], generator=g))
    x = pipe.sample(cond, noise, steps=steps, guidance=4.0)
    torch.cuda.synchronize()
    sig = OS.simple_scheduler(steps, OS.flux_sigma_table(seq_len=(hw // 2) ** 2))
    sd32 = {k: v.to(DEV) for k, v in sd.items()}
    ctx, y, gd = cond["crossattn"].to(DEV), cond["vector"].to(DEV), torch.full((B,), 4.0, device=DEV)

    def model(xx, sigma):
        with torch.no_grad():
            v = OF.flux_forward(sd32, cfg, xx, sigma, ctx, y, gd)
        return OS.const_denoised(xx, v, sigma.view(-1, 1, 1, 1))

    ref = OS.sample_euler(model, OS.const_noise_scaling(float(sig[0]), noise, torch.zeros_like(noise)).to(DEV), sig.to(DEV))
    mse = (x - ref).pow(2).mean()
    psnr = float(10 * torch.log10(ref.abs().max() ** 2 / mse))
    m, r = err_stats(x, ref)
    print(f"[parity] flux trajectory (graph={use_graph}): PSNR={psnr:.1f} dB max_abs={m:.3e} rel_rms={r:.3e}")
    assert psnr >= 30.0, psnr


def test_flux_guidance_follows_reference_bf16_cast():
    """KModel casts `guidance` to the computation dtype before the model scales it by 1000 (k_model.py:37-42,
    flux.py:53): in bf16, 3.5 * 1000 = 3504.  The engine reproduces that, so 3.5 and 3.504 give the same output."""
    cfg = OF.TINY_FLUX
    eng = _engine(cfg, OF.random_state_dict(cfg, seed=41))
    g = torch.Generator().manual_seed(42)
    x = torch.randn(1, 16, 16, 16, generator=g).to(DEV)
    ctx = torch.randn(1, 128, cfg["context_in_dim"], generator=g).to(BF).to(DEV)
    y = torch.randn(1, cfg["vec_in_dim"], generator=g).to(BF).to(DEV)
    t = torch.tensor([0.5], device=DEV)
    a = eng.forward(x, t, ctx, y, torch.tensor([3.5], device=DEV)).clone()
    b = eng.forward(x, t, ctx, y, torch.tensor([3.504], device=DEV)).clone()
    c = eng.forward(x, t, ctx, y, torch.tensor([3.0], device=DEV)).clone()
    torch.cuda.synchronize()
    assert torch.equal(a, b)
    assert not torch.equal(a, c)
