#!/usr/bin/env python
"""Parity of the library-owned multi-GPU paths under torchrun (one process per GPU):
bof_dist_gemm, bof_dist_csrmm, bof_kmeans_lloyd with the NCCL allreduce -- each against the oracle.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 tools/dist_check.py
"""
import json
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import __graft_entry__ as g  # noqa: E402
import oracle  # noqa: E402


def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    bof = g.load_package()
    from bof_b200 import dist as bdist
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = bof.Context(device=local)
    bdist.init_comm(ctx)
    out = {"world": world, "comm_world": ctx.comm_world()}
    TOL = 1e-5

    # ---- gemm: 3000 x 2100 x 1500, every (ta, tb), beta != 0; each rank owns a row block of C
    M, N, K = 3000, 2100, 1500
    for ta in "NT":
        for tb in "NT":
            a = oracle.gen_dense((M, K) if ta == "N" else (K, M), seed=1)
            b = oracle.gen_dense((K, N) if tb == "N" else (N, K), seed=2)
            c0 = oracle.gen_dense((M, N), seed=3)
            r0, r1 = bdist.row_shard(M, world, rank)
            a_loc = np.ascontiguousarray(a[r0:r1]) if ta == "N" else a[:, r0:r1]   # 'T': a column range, lda = M
            c_loc = c0[r0:r1].copy()
            lda = (K if ta == "N" else M)
            a_ptr = a_loc if ta == "N" else a.reshape(-1)[r0:]                      # view starting at column r0
            ctx.dist_gemm(ta, tb, r1 - r0, N, K, 1.5, 0.5, a_ptr, b, c_loc, lda, 0, N)
            ref = oracle.gemm("R", ta, tb, M, N, K, 1.5, 0.5, a, b, c0, acc64=True)
            err = oracle.rel_fro(c_loc, ref[r0:r1])
            out[f"gemm_{ta}{tb}"] = err
            assert err <= TOL, (ta, tb, err)
    # a size that takes the tensor-core panel path with several panels and blocks
    M, N, K = 2048 * world, 8192, 4096
    a, b = oracle.gen_dense((M, K), seed=4), oracle.gen_dense((K, N), seed=5)
    r0, r1 = bdist.row_shard(M, world, rank)
    c_loc = np.full((r1 - r0, N), np.nan, np.float32)
    ctx.dist_gemm("N", "N", r1 - r0, N, K, 1.0, 0.0, np.ascontiguousarray(a[r0:r1]), b, c_loc)
    ii = np.random.default_rng(rank).integers(0, r1 - r0, 64); jj = np.random.default_rng(7).integers(0, N, 64)
    ref = (a[r0:r1][ii].astype(np.float64) * b[:, jj].T.astype(np.float64)).sum(1)
    out["gemm_big_sampled"] = float(np.abs(c_loc[ii, jj] - ref).max() / np.abs(ref).max())
    assert out["gemm_big_sampled"] <= TOL and not np.isnan(c_loc).any()

    # uneven shards: the last rank owns no rows of C / a single row (a tiny product) -- it still owns panels of B and
    # must take part in the exchange
    for last_rows in ((0, 1) if world > 1 else (1,)):
        M, N, K = 600 * (world - 1) + last_rows, 1300, 900
        a, b, c0 = oracle.gen_dense((M, K), seed=14), oracle.gen_dense((K, N), seed=15), oracle.gen_dense((M, N), seed=16)
        r0, r1 = 600 * rank, (600 * (rank + 1) if rank < world - 1 else M)
        c_loc = c0[r0:r1].copy()
        ctx.dist_gemm("N", "N", r1 - r0, N, K, 1.5, 0.5, np.ascontiguousarray(a[r0:r1]), b, c_loc)
        ref = oracle.gemm("R", "N", "N", M, N, K, 1.5, 0.5, a, b, c0, acc64=True)
        assert r1 == r0 or oracle.rel_fro(c_loc, ref[r0:r1]) <= TOL, (rank, r0, r1)
    out["gemm_uneven_rows"] = True

    # ---- csrmm: nnz-balanced row shards, B shared
    m, n, k = 50000, 40000, 96
    av, ia, ja = oracle.gen_csr(m, n, 24, seed=6)
    B = oracle.gen_dense((n, k), seed=7); C0 = oracle.gen_dense((m, k), seed=8)
    r0, r1 = bdist.nnz_balanced_shard(ia, world, rank)
    z0, z1 = int(ia[r0]), int(ia[r1])
    c_loc = C0[r0:r1].copy()
    ctx.dist_csrmm(r1 - r0, n, k, 1.25, 0.75, av[z0:z1], ia[r0:r1 + 1], ja[z0:z1], B, c_loc)   # un-rebased offsets
    ref = oracle.csrmm("N", m, n, k, 1.25, 0.75, av, ia, ja, "R", B, C0, acc64=True)
    out["csrmm"] = oracle.rel_fro(c_loc, ref[r0:r1])
    assert out["csrmm"] <= TOL

    # empty shard on the last rank (all rows on the others)
    r0, r1 = (bdist.nnz_balanced_shard(ia, world - 1, rank) if rank < world - 1 else (m, m)) if world > 1 else (0, m)
    z0, z1 = int(ia[r0]), int(ia[r1])
    c_loc = C0[r0:r1].copy()
    ctx.dist_csrmm(r1 - r0, n, k, 1.25, 0.75, av[z0:z1], ia[r0:r1 + 1], ja[z0:z1], B, c_loc)
    assert r1 == r0 or oracle.rel_fro(c_loc, ref[r0:r1]) <= TOL
    out["csrmm_empty_last_shard"] = True

    # ---- kmeans: sharded + allreduce inside the library == the oracle on the whole set (ties-free start)
    rng = np.random.default_rng(9)
    Kc, d, P = 48, 40, 60000
    mu = (rng.normal(size=(Kc, d)) * 4).astype(np.float32)
    pts = (mu[rng.integers(0, Kc, P)] + 0.3 * rng.normal(size=(P, d))).astype(np.float32)
    c0 = (mu + 0.05 * rng.normal(size=(Kc, d))).astype(np.float32)
    p0, p1 = bdist.row_shard(P, world, rank)
    km = bof.KMeans(ctx, p1 - p0, Kc, d, np.ascontiguousarray(pts[p0:p1]), c0)
    km.lloyd(4)
    cent = np.zeros((Kc, d), np.float32); asg = np.zeros(p1 - p0, np.int64)
    km.get(cent, asg); km.close()
    c = c0.copy()
    for _ in range(4):
        c, a_ref, _ = oracle.lloyd_iter(pts, c)
    out["kmeans_centroids"] = oracle.rel_fro(cent, c)
    assert out["kmeans_centroids"] <= TOL
    if rank == 0:
        print(json.dumps(out), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
