"""GPU parity tests of the reference-facing plug-in layer: every entry point is called with
the reference's own signature/layout and compared with the oracle; the tests also assert that the sm_100a kernels —
not a deferred torch call — produced the result."""
import pytest
import torch

from oracle import configs as CF
from oracle import ops as O
from oracle import sampling as S
from oracle import unet as OU
from oracle import vae as OV
from tests.util import assert_close

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _r(*shape, seed=0, dtype=torch.float16, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(dtype).to(DEV)


def test_p1_attention_function_contract():
    from b200forge import attention as A
    from b200forge import ops
    q, k, v = _r(2, 300, 640, seed=1), _r(2, 77, 640, seed=2), _r(2, 77, 640, seed=3)
    n0 = ops.LAUNCHES
    out = A.attention_function(q, k, v, 10)
    assert ops.LAUNCHES == n0 + 1
    assert_close("P1 attention_function", out, O.attention(q.float(), k.float(), v.float(), 10), rel_rms=3e-3)
    # skip_reshape layout [b, H, L, Dh] (Flux call site, backend/nn/flux.py:15-18)
    qh = _r(1, 6, 256, 128, seed=4, dtype=torch.bfloat16)
    out2 = A.attention_function(qh, qh, qh, 6, skip_reshape=True)
    qf = qh.float().permute(0, 2, 1, 3).reshape(1, 256, 768)
    assert_close("P1 skip_reshape", out2, O.attention(qf, qf, qf, 6), rel_rms=1.5e-2)
    # non-contiguous view of a fused projection
    qkv = _r(2, 128, 3 * 128, seed=5)
    out3 = A.attention_function(qkv[..., :128], qkv[..., 128:256], qkv[..., 256:], 2)
    assert_close("P1 strided views", out3, O.attention(qkv[..., :128].float(), qkv[..., 128:256].float(), qkv[..., 256:].float(), 2), rel_rms=3e-3)


def test_p1_single_head_spatial():
    from b200forge import attention as A
    q, k, v = (_r(2, 128, 16, 16, seed=s, dtype=torch.bfloat16) for s in (6, 7, 8))
    out = A.attention_function_single_head_spatial(q, k, v)
    b, c, h, w = q.shape
    qt, kt, vt = (t.float().reshape(b, c, h * w).transpose(1, 2) for t in (q, k, v))
    ref = O.attention(qt, kt, vt, 1).transpose(1, 2).reshape(b, c, h, w)
    assert_close("P1 single-head spatial", out, ref, rel_rms=2e-2)


def test_p2_operations_modules():
    from b200forge import operations as BO
    from b200forge import ops
    ops_cls = BO.B200Operations
    d0 = BO.DEFERRED
    torch.manual_seed(0)
    lin = ops_cls.Linear(320, 640).to(DEV).half()
    x = _r(2, 77, 320, seed=9)
    assert_close("P2 Linear", lin(x), O.linear(x.float(), lin.weight.float(), lin.bias.float()), rel_rms=2e-3)
    for (cin, cout, k, s, p, hw) in [(64, 128, 3, 1, 1, 32), (320, 320, 3, 2, 1, 32), (128, 64, 1, 1, 0, 16), (8, 64, 3, 1, 1, 16)]:
        conv = ops_cls.Conv2d(cin, cout, k, stride=s, padding=p).to(DEV).half()
        xc = _r(2, cin, hw, hw, seed=10)
        ref = O.conv2d(xc.float(), conv.weight.float(), conv.bias.float(), stride=s, padding=p)
        assert_close(f"P2 Conv2d {cin}->{cout} k{k} s{s}", conv(xc), ref, rel_rms=2e-3)
    gn = ops_cls.GroupNorm(32, 320, eps=1e-6).to(DEV).half()
    with torch.no_grad():
        gn.weight.copy_(1 + 0.1 * torch.randn(320))
        gn.bias.copy_(0.1 * torch.randn(320))
    xg = _r(2, 320, 16, 16, seed=11)
    assert_close("P2 GroupNorm", gn(xg), O.group_norm(xg.float(), 32, gn.weight.float(), gn.bias.float(), 1e-6), rel_rms=2e-3)
    ln = ops_cls.LayerNorm(640).to(DEV).half()
    xl = _r(2, 100, 640, seed=12)
    assert_close("P2 LayerNorm", ln(xl), O.layer_norm(xl.float(), ln.weight.float(), ln.bias.float(), 1e-5), rel_rms=2e-3)
    assert BO.DEFERRED == d0, "a supported fp16 CUDA call was deferred to torch"
    # state-dict compatibility + LoRA-style parameter replacement invalidates the packed weight
    conv = ops_cls.Conv2d(64, 64, 3, padding=1).to(DEV).half()
    xc = _r(1, 64, 16, 16, seed=13)
    y1 = conv(xc)
    with torch.no_grad():
        conv.weight.mul_(2.0)
    y2 = conv(xc)
    assert_close("P2 repack after in-place weight edit", y2, O.conv2d(xc.float(), conv.weight.float(), conv.bias.float()), rel_rms=2e-3)
    assert not torch.allclose(y1, y2)


def test_p3_model_function_wrapper_vs_oracle():
    from b200forge import plugin
    from b200forge.unet_engine import UNetEngine
    cfg = CF.CONFIGS["tiny_xl"]
    sd = OU.random_state_dict(cfg, seed=1)
    eng = UNetEngine(cfg, sd, dtype=torch.float16, device=DEV)
    pred = S.EpsPrediction()

    class P:  # the two things the wrapper reads from Forge's predictor
        prediction_type = "epsilon"
        timestep = staticmethod(lambda s: pred.timestep(s.cpu()).to(s.device))

    w = plugin.UNetWrapper(eng, P())
    g = torch.Generator().manual_seed(3)
    n = 4
    x = (torch.randn(n, 4, 16, 16, generator=g) * 5).to(DEV)
    sigma = torch.tensor([14.6, 3.0, 0.7, 0.05], device=DEV)
    ctx = torch.randn(n, 77, cfg["context_dim"], generator=g).to(DEV)
    y = torch.randn(n, cfg["adm_in_channels"], generator=g).to(DEV)
    c = {"c_crossattn": ctx, "y": y, "transformer_options": {"cond_or_uncond": [1, 0]}}
    called = []
    out = w(lambda *a, **k: called.append(1), {"input": x, "timestep": sigma, "c": c, "cond_or_uncond": [1, 0]})
    assert not called and w.calls_fast == 1
    den = S.Denoiser(lambda xc, t, cx, yy: OU.unet_forward(sd, cfg, xc, t, cx, yy), pred, {}, {}, 1.0)
    with torch.no_grad():
        ref = den.apply_model(x.cpu(), sigma.cpu(), ctx.cpu(), y.cpu())
    assert_close("P3 wrapper denoised vs oracle KModel.apply_model", out, ref, rel_rms=3e-3)
    # a hook the fused forward cannot honour -> the reference callable is used, untouched
    c2 = dict(c, transformer_options={"patches": {"attn2_patch": [lambda *a: a]}})
    sentinel = torch.zeros(1)
    out2 = w(lambda xx, ss, **kw: sentinel, {"input": x, "timestep": sigma, "c": c2, "cond_or_uncond": [1, 0]})
    assert out2 is sentinel and w.calls_reference == 1


@pytest.mark.parametrize("name", ["sample_euler", "sample_euler_ancestral", "sample_dpmpp_2m"])
def test_p4_k_diffusion_sampler_contract(name):
    from b200forge import k_samplers
    torch.manual_seed(0)
    B, C, H, W, steps = 2, 4, 16, 16, 7
    pred = S.EpsPrediction()
    sig = S.get_sigmas_karras(steps, float(pred.sigma_min), float(pred.sigma_max)) if name == "sample_dpmpp_2m" \
        else S.get_sigmas_uniform(pred, steps)
    eps_seq = [torch.randn(B, C, H, W) for _ in range(steps)]
    noise_seq = [torch.randn(B, C, H, W) for _ in range(steps)]
    x0 = torch.randn(B, C, H, W) * float(sig[0])

    def make_model(dev):
        it = iter(range(steps))
        return lambda x, sigma, **kw: x - eps_seq[next(it)].to(dev) * sigma.view(-1, 1, 1, 1)

    if name == "sample_euler":
        ref = S.sample_euler(make_model("cpu"), x0.clone(), sig)
    elif name == "sample_euler_ancestral":
        k = iter(range(steps))
        ref = S.sample_euler_ancestral(make_model("cpu"), x0.clone(), sig, lambda: noise_seq[next(k)])
    else:
        ref = S.sample_dpmpp_2m(make_model("cpu"), x0.clone(), sig)
    seen = []
    cb = lambda d: seen.append((d["i"], float(d["sigma"]), d["denoised"].shape))  # noqa: E731
    kw = {}
    if name == "sample_euler_ancestral":
        k2 = iter(range(steps))
        kw["noise_sampler"] = lambda s, sn: noise_seq[next(k2)].to(DEV)
    out = getattr(k_samplers, name)(make_model(DEV), x0.to(DEV), sig.to(DEV), extra_args={}, callback=cb, disable=True, **kw)
    torch.cuda.synchronize()
    assert [s[0] for s in seen] == list(range(steps))
    assert_close(f"P4 {name} vs oracle loop", out, ref, rel_rms=2e-6)


def test_p5_vae_decode_wrapper():
    from b200forge import plugin
    from b200forge.vae_engine import VAEDecoderEngine
    cfg = CF.VAE_CONFIGS["tiny"]
    sd = OV.random_state_dict(cfg, seed=3)
    eng = VAEDecoderEngine(cfg, sd, dtype=torch.float16, device=DEV)
    w = plugin.VAEDecodeWrapper(eng)
    z = _r(2, 4, 16, 16, seed=14, dtype=torch.float32)  # processed-out latent, as Forge passes it
    img = w(lambda s: None, z)
    with torch.no_grad():
        ref = torch.clamp((OV.decode(sd, cfg, z.cpu()) + 1.0) / 2.0, 0.0, 1.0).movedim(1, -1)
    assert_close("P5 VAE decode wrapper", img, ref, max_abs=2e-2, rel_rms=4e-3)
    assert eng.scaling == cfg["scaling_factor"]


def test_p3_flux_wrapper_vs_oracle():
    """The same hook with a Flux KModel: `c` carries T5 states, pooled CLIP and the distilled guidance; the result is
    PredictionFlux.calculate_denoised = x - sigma * v (k_prediction.py:81-92 'const')."""
    from b200forge import plugin
    from b200forge.flux_engine import FluxEngine
    from oracle import flux as OF
    cfg = OF.TINY_FLUX
    sd = OF.random_state_dict(cfg, seed=5)
    eng = FluxEngine(cfg, sd, dtype=torch.bfloat16, device=DEV)

    class P:
        prediction_type = "const"
        timestep = staticmethod(lambda s: s)

    w = plugin.FluxWrapper(eng, P())
    g = torch.Generator().manual_seed(6)
    n = 2
    x = torch.randn(n, 16, 16, 16, generator=g).to(DEV)
    sigma = torch.tensor([0.9, 0.2], device=DEV)
    ctx = torch.randn(n, 128, cfg["context_in_dim"], generator=g).to(DEV)
    y = torch.randn(n, cfg["vec_in_dim"], generator=g).to(DEV)
    gd = torch.full((n,), 4.0)
    c = {"c_crossattn": ctx, "y": y, "guidance": gd, "transformer_options": {}}
    called = []
    out = w(lambda *a, **k: called.append(1), {"input": x, "timestep": sigma, "c": c, "cond_or_uncond": [0]})
    torch.cuda.synchronize()
    assert not called and w.calls_fast == 1 and out.dtype == torch.float32 and out.shape == x.shape
    with torch.no_grad():
        v = OF.flux_forward(sd, cfg, x.cpu(), sigma.cpu(), ctx.cpu(), y.cpu(), gd)
    ref = x.cpu() - v * sigma.cpu().view(-1, 1, 1, 1)
    assert_close("P3 Flux wrapper denoised vs oracle", out, ref, rel_rms=3e-2)
    # odd latent size: the reference's circular-pad branch (flux.py:394-397, 412) is served by the fused path too
    sentinel = torch.zeros(1)
    xo = torch.randn(n, 16, 15, 16, generator=g).to(DEV)
    outo = w(lambda xx, ss, **kw: sentinel, {"input": xo, "timestep": sigma, "c": c, "cond_or_uncond": [0]})
    torch.cuda.synchronize()
    with torch.no_grad():
        vo = OF.flux_forward(sd, cfg, xo.cpu(), sigma.cpu(), ctx.cpu(), y.cpu(), gd)
    assert outo is not sentinel
    assert_close("P3 Flux wrapper, odd latent, vs oracle", outo, xo.cpu() - vo * sigma.cpu().view(-1, 1, 1, 1), rel_rms=3e-2)
    # a patch hook -> Forge's own apply_model
    c2 = dict(c, transformer_options={"patches_replace": {"dit": {}}})
    assert w(lambda xx, ss, **kw: sentinel, {"input": x, "timestep": sigma, "c": c2, "cond_or_uncond": [0]}) is sentinel
    assert w.calls_reference == 1


def test_p5_vae_encode_wrapper():
    """`model_vae_encode_wrapper` (backend/patcher/vae.py:186-191): NHWC pixels in [0,1] -> un-scaled posterior sample."""
    from b200forge import plugin
    from b200forge.vae_engine import VAEEncoderEngine
    from oracle import vae as OV
    cfg = CF.VAE_CONFIGS["tiny"]
    sd = OV.random_encoder_state_dict(cfg, seed=8)
    eng = VAEEncoderEngine(cfg, sd, dtype=torch.float16, device=DEV)
    w = plugin.VAEEncodeWrapper(eng, output_device="cpu")
    g = torch.Generator().manual_seed(9)
    px = torch.rand(2, 64, 64, 3, generator=g)
    torch.manual_seed(77)
    out = w(lambda p: None, px)
    torch.manual_seed(77)
    noise = torch.randn(2, cfg["latent_channels"], 32, 32)  # the tiny VAE has one downsample: 64 -> 32
    with torch.no_grad():
        ref = OV.posterior(OV.encode_moments(sd, cfg, 2.0 * px.movedim(-1, 1) - 1.0), noise)
    assert out.device.type == "cpu" and out.dtype == torch.float32
    assert_close("P5 encode wrapper vs oracle", out, ref, max_abs=2e-2, rel_rms=5e-3)
    sentinel = torch.zeros(1)
    assert w(lambda p: sentinel, torch.rand(1, 61, 64, 3)) is sentinel  # odd size: the stride-2 level does not divide -> Forge's own encode


@pytest.mark.parametrize("name", ["sample_heun", "sample_dpm_2", "sample_dpm_2_ancestral", "sample_dpmpp_2s_ancestral"])
def test_p4_two_evaluation_samplers_vs_reference_golden(name):
    """Fused Heun / DPM2 / DPM2 a / DPM++ 2S a (k_samplers.py) against the reference's own loops on CPU fp32 around the
    same toy denoiser (tests/golden/samplers_toy.pt): same call contract, callback per step, noise order."""
    import os
    from b200forge import k_samplers
    g = torch.load(os.path.join(os.path.dirname(__file__), "golden", "samplers_toy.pt"), weights_only=False)
    steps = len(g["sigmas"]) - 1
    k = iter(range(g["noise"].shape[0]))
    kw = dict(noise_sampler=lambda s, sn: g["noise"][next(k)].to(DEV)) if "ancestral" in name else {}
    seen = []
    out = getattr(k_samplers, name)(lambda x, sigma, **kwargs: S.toy_denoiser(x, sigma), g["x0"].to(DEV), g["sigmas"].to(DEV),
                                    extra_args={}, callback=lambda d: seen.append(d["i"]), disable=True, **kw)
    torch.cuda.synchronize()
    assert seen == list(range(steps))
    assert_close(f"P4 {name} vs reference golden", out, g[name], rel_rms=2e-5)
