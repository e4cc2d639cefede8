#!/usr/bin/env python
"""Per-kernel micro-benchmarks (CUDA events, L2 flushed between iterations) at the shapes of a workload.
Prints one JSON object per kernel family: ms, achieved GB/s or TFLOP/s and the fraction of the measured
peak (MEASURED_PEAKS.json).  Used to fill profiles/*.md; not part of the driver contract."""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rtucker_b200 import ops  # noqa: E402

SHAPES = {"wn18rr": (40943, 22, (10, 200, 200)), "fb15k237": (14541, 474, (200, 20, 20)),
          "synthetic-1m": (1000000, 1000, (200, 200, 200))}


def timeit(fn, iters=10, flush=None):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush is not None:
            flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="wn18rr", choices=list(SHAPES))
    ap.add_argument("--only", default="")
    ap.add_argument("--variant", type=int, default=0)
    args = ap.parse_args()
    N, M, rank = SHAPES[args.workload]
    B = 512
    dev = torch.device("cuda:0")
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.isfile(
        os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    g = torch.Generator(device="cpu").manual_seed(0)
    r0, r1, r2 = rank
    out = {}

    def want(name):
        return not args.only or name in args.only.split(",")

    if want("rank"):
        P = torch.rand(B, N, device=dev)
        tgt = torch.randint(0, N, (B,), device=dev, dtype=torch.int32)
        off = torch.arange(0, (B + 1) * 16, 16, device=dev, dtype=torch.int32)
        idx = torch.randint(0, N, (B * 16,), device=dev, dtype=torch.int32)
        ms = timeit(lambda: ops.rank_filtered(P, tgt, off, idx), flush=flush)
        gbs = 4.0 * B * N / ms / 1e6
        out["rank_filtered"] = dict(ms=ms, GBps=gbs, frac_hbm=gbs / peaks["hbm_gbs"], bytes=4.0 * B * N)
    if want("score"):
        q = torch.randn(B, r2, device=dev) / r2 ** 0.5
        O = torch.randn(N, r2, device=dev)
        off = torch.arange(0, (B + 1) * 2, 2, device=dev, dtype=torch.int32)
        idx = torch.randint(0, N, (B * 2,), device=dev, dtype=torch.int32)
        ms = timeit(lambda: ops.score_bce_fwd_bwd(q, q, O, off, idx, 0.1, variant=args.variant), flush=flush)
        tf = 6.0 * B * N * r2 / ms / 1e9
        out["score_bce_fwd_bwd"] = dict(ms=ms, TFLOPs=tf, frac_bf16_peak=tf / peaks["bf16_tflops"],
                                        flops=6.0 * B * N * r2, bytes=8.0 * N * r2, variant=args.variant)
        tgt = torch.randint(0, N, (B,), device=dev, dtype=torch.int32)
        pt = ops.target_prob(q, O, tgt)
        ms = timeit(lambda: ops.score_rank_fused(q, O, tgt, pt, off, idx), flush=flush)
        out["score_rank_fused"] = dict(ms=ms, TFLOPs=2.0 * B * N * r2 / ms / 1e9)
    if want("query"):
        core = torch.randn(rank, device=dev)
        rr, sr, H = torch.randn(B, r0, device=dev), torch.randn(B, r1, device=dev), torch.randn(B, r2, device=dev)
        ms = timeit(lambda: ops.query_fwd(core, rr, sr), flush=flush)
        out["query_fwd"] = dict(ms=ms, TFLOPs=2.0 * B * r0 * r1 * r2 / ms / 1e9)
        ms = timeit(lambda: ops.query_bwd(core, rr, sr, H), flush=flush)
        out["query_bwd"] = dict(ms=ms, TFLOPs=6.0 * B * r0 * r1 * r2 / ms / 1e9)
    if want("tall"):
        r = r1
        U, V = torch.randn(N, r, device=dev), torch.randn(N, r, device=dev)
        for precise in (False, True):
            ms = timeit(lambda: ops.gram(U, V, precise=precise), flush=flush)
            out["gram_precise" if precise else "gram"] = dict(ms=ms, TFLOPs=2.0 * N * r * r / ms / 1e9,
                                                              GBps=8.0 * N * r / ms / 1e6)
        K = torch.randn(r, r, device=dev, dtype=torch.float64)
        Y = torch.empty(N, r, device=dev)
        ms = timeit(lambda: ops.apply(Y, None, None, [(U, K), (V, K)]), flush=flush)
        out["apply_2term"] = dict(ms=ms, TFLOPs=4.0 * N * r * r / ms / 1e9, GBps=12.0 * N * r / ms / 1e6)
    if want("eig"):
        for n in sorted({2 * r0, 2 * r1}):
            X = torch.randn(n, 4 * n, dtype=torch.float64, device=dev)
            A = X @ X.T
            ms = timeit(lambda: ops.eigh(A), iters=5)
            out[f"eigh_{n}"] = dict(ms=ms)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
