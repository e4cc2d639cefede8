"""Times the HOSVD subspace kernel (csrc/subspace.cu) against the block-Jacobi eigensolver (csrc/eig.cu) on
graded spectra like those of a training run (run on the GPU box: python tools/subspace_bench.py)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rtucker_b200 import ops  # noqa: E402


def graded(n, r, top, bottom, ratio, seed):
    g = torch.Generator().manual_seed(seed)
    Q, _ = torch.linalg.qr(torch.randn(n, n, generator=g, dtype=torch.float64))
    lam = torch.cat([torch.logspace(np.log10(top), np.log10(bottom), r, dtype=torch.float64),
                     torch.logspace(np.log10(bottom / ratio), np.log10(bottom / ratio) - 6, n - r, dtype=torch.float64)])
    A = (Q * lam) @ Q.T
    return 0.5 * (A + A.T)


def timeit(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    dev = torch.device("cuda:0")
    for (n, r, top, bottom, ratio) in [(400, 200, 1e6, 0.3, 1.3), (400, 200, 1e5, 0.1, 50.0), (400, 200, 10.0, 1.0, 2.0),
                                       (40, 20, 1e3, 1.0, 3.0)]:
        A = graded(n, r, top, bottom, ratio, 1).to(dev)
        Y, info = ops.dominant_subspace(A, r)
        ms = timeit(lambda: ops.dominant_subspace(A, r))
        ms_e = timeit(lambda: ops.eigh(A), reps=2)
        w, V = torch.linalg.eigh(A)
        Vr = V[:, -r:]
        perr = float((Y @ Y.T - Vr @ Vr.T).norm())
        orth = float((Y.T @ Y - torch.eye(r, dtype=torch.float64, device=dev)).abs().max())
        print(f"n={n} r={r} top={top:g} bottom={bottom:g} ratio={ratio:g}: subspace {ms:.3f} ms "
              f"(tc2 {int(info[0])} its, ns {int(info[1])} its, {1e3 * ms / max(1, int(info[0]) + 2 * int(info[1])):.1f} us/round), "
              f"jacobi eigh {ms_e:.3f} ms; projector err {perr:.2e}, orth {orth:.1e}", flush=True)


if __name__ == "__main__":
    main()
