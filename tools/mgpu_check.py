#!/usr/bin/env python
"""One process driving several GPUs (bof_mgpu_*, what BOF_GPUS=n enables behind flash::): gemm in every layout incl.
shards without rows, csrmm, csrgemv N/T, k-means with the in-library allreduce -- each against the oracle.
Run by tests/test_gpu_zz_multi.py in a child process so that a hang is a test failure, not a stalled suite."""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import __graft_entry__ as g  # noqa: E402
import oracle  # noqa: E402

TOL = 1e-5
NGPU = torch.cuda.device_count()


def main():
    bof = g.load_package()
    n_use = min(NGPU, 4)
    with bof.MultiGpu(ndev=n_use) as mg:
        assert mg.count() == n_use
        # M = 2100 leaves the last of 4 GPUs without rows (shares are rounded up to 256 rows), M = 200 the last of 2
        for o, ta, tb, M in (("R", "N", "N", 2100), ("R", "T", "N", 2100), ("C", "N", "T", 2100), ("R", "N", "T", 2100),
                             ("R", "N", "N", 200)):
            N, K = 1700, 1300
            ar, ac = (M, K) if ta == "N" else (K, M)
            br, bc = (K, N) if tb == "N" else (N, K)
            cr, cc = M, N
            if o == "C":
                ar, ac, br, bc, cr, cc = ac, ar, bc, br, cc, cr
            a, b, c0 = oracle.gen_dense((ar, ac), seed=1), oracle.gen_dense((br, bc), seed=2), oracle.gen_dense((cr, cc), seed=3)
            c = c0.copy()
            mg.gemm(o, ta, tb, M, N, K, 1.5, 0.5, a, b, c)
            assert oracle.rel_fro(c, oracle.gemm(o, ta, tb, M, N, K, 1.5, 0.5, a, b, c0, acc64=True)) <= TOL, (o, ta, tb)
            print("gemm", o, ta, tb, M, "ok", flush=True)
        m, n, k = 60000, 45000, 128
        av, ia, ja = oracle.gen_csr(m, n, 20, seed=4)
        B, C0 = oracle.gen_dense((n, k), seed=5), oracle.gen_dense((m, k), seed=6)
        C = C0.copy()
        mg.csrmm("N", m, n, k, 1.25, 0.75, av, ia, ja, "R", B, C)
        assert oracle.rel_fro(C, oracle.csrmm("N", m, n, k, 1.25, 0.75, av, ia, ja, "R", B, C0, acc64=True)) <= TOL
        print("csrmm ok", flush=True)
        x, xt = oracle.gen_dense((n,), seed=7), oracle.gen_dense((m,), seed=8)
        y = np.full(m, np.nan, np.float32)
        mg.csrgemv("N", m, n, av, ia, ja, x, y)
        assert oracle.rel_fro(y, oracle.csrgemv("N", m, n, av, ia, ja, x, acc64=True)) <= TOL
        yt = np.full(n, np.nan, np.float32)
        mg.csrgemv("T", m, n, av, ia, ja, xt, yt)
        assert oracle.rel_fro(yt, oracle.csrgemv("T", m, n, av, ia, ja, xt, acc64=True)) <= TOL
        yt2 = np.full(n, np.nan, np.float32)
        mg.csrgemv("T", m, n, av, ia, ja, xt, yt2)
        assert oracle.rel_fro(yt2, yt) <= 1e-6
        print("csrgemv ok", flush=True)
        rng = np.random.default_rng(9)
        Kc, d, P = 32, 24, 40000
        mu = (rng.normal(size=(Kc, d)) * 4).astype(np.float32)
        pts = (mu[rng.integers(0, Kc, P)] + 0.3 * rng.normal(size=(P, d))).astype(np.float32)
        c0 = (mu + 0.05 * rng.normal(size=(Kc, d))).astype(np.float32)
        cent = c0.copy(); asg = np.zeros(P, np.int64)
        mg.kmeans_lloyd(P, Kc, d, pts, cent, 3, asg)
        c = c0.copy()
        for _ in range(3):
            c, a_ref, _ = oracle.lloyd_iter(pts, c)
        assert oracle.rel_fro(cent, c) <= TOL
    print("mgpu ok", n_use, flush=True)


if __name__ == "__main__":
    main()
