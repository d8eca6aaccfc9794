"""A/B timing of the LayerNorm-folded GEMMs against LayerNorm kernel + plain GEMM (development aid)."""
import sys
import torch
sys.path.insert(0, ".")
sys.path.insert(0, "scripts")
from b200forge import ops
from kernel_perf import timeit

DEV = "cuda"
for M, C in [(16384, 1280), (65536, 640)]:
    t = torch.randn(M, C, device=DEV, dtype=torch.float16)
    g = torch.randn(C, device=DEV, dtype=torch.float16) * 0.1 + 1
    b = torch.randn(C, device=DEV, dtype=torch.float16) * 0.1
    res = torch.randn(M, C, device=DEV, dtype=torch.float16)
    wo = torch.randn(C, C, device=DEV, dtype=torch.float16) * C ** -0.5
    bo = torch.randn(C, device=DEV, dtype=torch.float16)
    stats = torch.zeros(M, 2, device=DEV)
    for name, N, epi in [("qkv", 3 * C, ops.EPI_NONE), ("ff1-geglu", 8 * C, ops.EPI_GEGLU)]:
        w = torch.randn(N, C, device=DEV, dtype=torch.float16) * C ** -0.5
        bias = torch.randn(N, device=DEV, dtype=torch.float16) if epi == ops.EPI_GEGLU else None
        wf, c, d = ops.fold_layernorm(w, bias, g, b)
        if epi == ops.EPI_GEGLU:
            wp, bp = ops.pack_geglu(w, bias, 256)
            wfp, cp = ops.pack_geglu(wf, c, 256)
            _, dp = ops.pack_geglu(wf, d, 256)
        else:
            wp, bp, wfp, cp, dp = w, None, wf, c, d
        n_out = N // 2 if epi == ops.EPI_GEGLU else N
        out = torch.empty(M, n_out, device=DEV, dtype=torch.float16)
        nrm = torch.empty_like(t)

        def unfused():
            ops.layernorm(t, g, b, out=nrm)
            ops.gemm(nrm, wp, bp, epilogue=epi, block_n=256 if epi else 0, out=out)

        def fused():
            ops.gemm(t, wfp, None, epilogue=epi, block_n=256 if epi else 0, out=out, ln=(stats, cp, dp, 1e-5))
        a, f = timeit(unfused), timeit(fused)
        print(f"M={M} C={C} {name:10s} LN+GEMM {a*1e6:8.1f} us   folded {f*1e6:8.1f} us")
    att = torch.randn(M, C, device=DEV, dtype=torch.float16)
    o1 = torch.empty(M, C, device=DEV, dtype=torch.float16)
    p0 = timeit(lambda: ops.gemm(att, wo, bo, residual=res, out=o1))
    p1 = timeit(lambda: ops.gemm(att, wo, bo, residual=res, out=o1, row_stats_out=ops.zero_(stats)))
    print(f"M={M} C={C} producer  plain {p0*1e6:8.1f} us   +row stats (+memset) {p1*1e6:8.1f} us")
