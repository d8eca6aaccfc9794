"""Shared helpers for the parity tests."""
import torch


def err_stats(got: torch.Tensor, ref: torch.Tensor):
    got = got.detach().float().cpu()
    ref = ref.detach().float().cpu()
    diff = (got - ref).abs()
    max_abs = diff.max().item()
    rel_rms = (diff.pow(2).mean().sqrt() / ref.pow(2).mean().sqrt().clamp_min(1e-20)).item()
    return max_abs, rel_rms


def assert_close(name, got, ref, *, max_abs=None, rel_rms=None, max_rel=None):
    """max_abs: bound on max |got - ref|; max_rel: the same bound in units of the reference's RMS (for outputs whose scale is
    O(1) but not exactly 1: SURVEY 8d states max-abs <= 2e-2 for unit-scale activations); rel_rms: RMS error / RMS reference."""
    assert got.shape == ref.shape, (name, got.shape, ref.shape)
    assert torch.isfinite(got.float()).all(), f"{name}: non-finite values"
    m, r = err_stats(got, ref)
    scale = ref.detach().float().abs().max().item()
    rms = ref.detach().float().pow(2).mean().sqrt().item()
    print(f"[parity] {name}: max_abs={m:.3e} rel_rms={r:.3e} (ref max {scale:.3e}, rms {rms:.3e})")
    if max_abs is not None:
        assert m <= max_abs, f"{name}: max_abs {m:.3e} > {max_abs:.3e}"
    if max_rel is not None:
        assert m <= max_rel * max(rms, 1e-20), f"{name}: max_abs {m:.3e} > {max_rel:.1e} x ref rms {rms:.3e}"
    if rel_rms is not None:
        assert r <= rel_rms, f"{name}: rel_rms {r:.3e} > {rel_rms:.3e}"
