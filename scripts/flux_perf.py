"""Flux.1-dev (BASELINE.json configs[4]: 1024x1024, batch 4, bf16) transformer forward on the sm_100a path: ms per
forward (CUDA events, CUDA graph replay), achieved TFLOP/s against SURVEY.md §8d's 69 466.6 GFLOP/sample, and a
per-kernel-family breakdown from an instrumented eager pass.  Development aid; synthetic weights generated on the GPU.

    python scripts/flux_perf.py [--batch 4] [--hw 128] [--txt 256] [--iters 5] [--depth 19 --single 38]
"""
import argparse
import json
import sys
from collections import defaultdict

import torch

sys.path.insert(0, ".")
from b200forge import ops, synthetic  # noqa: E402
from b200forge.flux_engine import FluxEngine  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=4)
    ap.add_argument("--hw", type=int, default=128)
    ap.add_argument("--txt", type=int, default=256)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--depth", type=int, default=19)
    ap.add_argument("--single", type=int, default=38)
    a = ap.parse_args()
    dev = "cuda"
    cfg = dict(synthetic.FLUX_DEV, depth=a.depth, depth_single_blocks=a.single)
    sd = synthetic.random_flux_state_dict(cfg, device=dev)
    eng = FluxEngine(cfg, sd, device=dev)
    del sd
    torch.cuda.empty_cache()
    B = a.batch
    x = torch.randn(B, 16, a.hw, a.hw, device=dev)
    ctx = torch.randn(B, a.txt, cfg["context_in_dim"], device=dev, dtype=torch.bfloat16)
    y = torch.randn(B, cfg["vec_in_dim"], device=dev, dtype=torch.bfloat16)
    t = torch.full((B,), 0.7, device=dev)
    gd = torch.full((B,), 3.5, device=dev)
    out = torch.empty(B, a.hw, a.hw, 16, device=dev, dtype=torch.bfloat16)
    for _ in range(2):
        eng.forward_nhwc(x, t, ctx, y, gd, out=out)
    torch.cuda.synchronize()
    assert torch.isfinite(out.float()).all()
    print("output std", float(out.float().std()), "mem GB", torch.cuda.max_memory_allocated() / 1e9, flush=True)
    # instrumented eager pass
    ops.PROFILE = []
    n0 = ops.LAUNCHES
    eng.forward_nhwc(x, t, ctx, y, gd, out=out)
    torch.cuda.synchronize()
    launches = ops.LAUNCHES - n0
    fam = defaultdict(lambda: [0, 0.0, 0.0, 0.0])
    for name, fl, by, s, e, *_ in ops.PROFILE:
        f = fam[name]
        f[0] += 1
        f[1] += s.elapsed_time(e)
        f[2] += fl
        f[3] += by
    ops.PROFILE = None
    tot = sum(v[1] for v in fam.values())
    for k, v in sorted(fam.items(), key=lambda kv: -kv[1][1]):
        extra = f"{v[2] / v[1] / 1e9:8.1f} TF/s" if v[2] else f"{v[3] / v[1] / 1e6:8.1f} GB/s"
        print(f"  {k:18s} n={v[0]:4d} {v[1]:8.2f} ms ({v[1] / tot * 100:5.1f}%) {extra}", flush=True)
    # graph replay timing
    g = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        eng.forward_nhwc(x, t, ctx, y, gd, out=out)
        with torch.cuda.graph(g, stream=side):
            eng.forward_nhwc(x, t, ctx, y, gd, out=out)
    torch.cuda.current_stream().wait_stream(side)
    g.replay()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(a.iters):
        g.replay()
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e) / a.iters
    # algorithmic flops: SURVEY §8d figure for the full model at 4096 + 256 tokens, else the summed GEMM/attention flops
    full = a.depth == 19 and a.single == 38 and a.hw == 128 and a.txt == 256
    flops = 69466.6e9 * B if full else sum(v[2] for v in fam.values())
    print(json.dumps({"flux_forward_ms": ms, "batch": B, "tokens": a.txt + (a.hw // 2) ** 2, "launches": launches,
                      "tflops_per_s": flops / ms / 1e9, "flops_source": "SURVEY 8d" if full else "sum of kernel flops",
                      "eager_sum_ms": tot}), flush=True)


if __name__ == "__main__":
    main()
