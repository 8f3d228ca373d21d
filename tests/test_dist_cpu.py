"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: shard rules tile the problem, sharded
results concatenate to the single-process result, and the k-means sums/counts allreduce composes to
the reference update.  The arithmetic here is the oracle's -- what is under test is the sharding."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q):
    import __graft_entry__ as g
    bof = g.load_package()
    from bof_b200 import dist as bdist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # ---- csrmm: nnz-balanced row shards, no collective on the data path ----
        rng = np.random.default_rng(0)
        m, n, k = 500, 300, 16
        counts = rng.integers(0, 40, size=m); counts[100:200] = 0
        ia = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
        ja = np.concatenate([np.sort(rng.choice(n, c, replace=False)) for c in counts]).astype(np.int64)
        a = rng.random(ia[-1], dtype=np.float32)
        B = rng.random((n, k), dtype=np.float32)
        r0, r1 = bdist.nnz_balanced_shard(ia, world, rank)
        z0, z1 = ia[r0], ia[r1]
        mine = oracle.csrmm("N", r1 - r0, n, k, 1.0, 0.0, a[z0:z1], ia[r0:r1 + 1], ja[z0:z1], "R", B,
                            np.zeros((r1 - r0, k), np.float32))
        full = oracle.csrmm("N", m, n, k, 1.0, 0.0, a, ia, ja, "R", B, np.zeros((m, k), np.float32))
        ok_spmm = np.array_equal(mine, full[r0:r1])
        spans = [None] * world
        dist.all_gather_object(spans, (r0, r1, int(z1 - z0)))
        # ---- csrgemv: 'N' shards are disjoint rows; 'T' partials summed on the host in rank order ----
        x_n = rng.random(n, dtype=np.float32); x_t = rng.random(m, dtype=np.float32)
        y_n = oracle.csrgemv("N", r1 - r0, n, a[z0:z1], ia[r0:r1 + 1], ja[z0:z1], x_n)
        ok_gemv = np.array_equal(y_n, oracle.csrgemv("N", m, n, a, ia, ja, x_n)[r0:r1])
        part = oracle.csrgemv("T", r1 - r0, n, a[z0:z1], ia[r0:r1 + 1], ja[z0:z1], x_t[r0:r1])
        y_t = bdist.sum_partials_fixed_order(part)
        ok_gemv = ok_gemv and oracle.rel_fro(y_t, oracle.csrgemv("T", m, n, a, ia, ja, x_t, acc64=True)) < 1e-5
        sums_t = [None] * world
        dist.all_gather_object(sums_t, y_t.tobytes())
        ok_gemv = ok_gemv and all(b == sums_t[0] for b in sums_t)  # bitwise identical on every rank
        ok_spmm = ok_spmm and ok_gemv
        # ---- gemm: equal row shards ----
        g0, g1 = bdist.row_shard(37, world, rank)
        # ---- kmeans: local sums/counts -> allreduce -> divide == single-process update ----
        P, K, d = 400, 5, 8
        pts = rng.normal(size=(P, d)).astype(np.float32)
        cent = pts[:K].copy()
        p0, p1 = bdist.row_shard(P, world, rank)
        assign, _ = oracle.kmeans_assign(pts[p0:p1], cent)
        sums = np.zeros((K, d), np.float32); cnt = np.zeros(K, np.float32)
        np.add.at(sums, assign, pts[p0:p1]); np.add.at(cnt, assign, 1.0)
        buf = torch.from_numpy(np.concatenate([sums.ravel(), cnt]))
        dist.all_reduce(buf)  # the only collective on the path
        tot = buf.numpy()
        s, c = tot[:K * d].reshape(K, d), tot[K * d:]
        new_c = np.where(c[:, None] > 0, s / np.maximum(c[:, None], 1), 0).astype(np.float32)
        ref_c, ref_a, _ = oracle.lloyd_iter(pts, cent)
        ok_km = oracle.rel_fro(new_c, ref_c) < 1e-5 and np.array_equal(assign, ref_a[p0:p1]) and int(c.sum()) == P
        q.put((rank, ok_spmm, spans, (g0, g1), ok_km))
    finally:
        dist.destroy_process_group()


def test_world2_gloo_sharding():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    [p.start() for p in procs]
    res = sorted(q.get(timeout=120) for _ in range(world))
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    for rank, ok_spmm, spans, gspan, ok_km in res:
        assert ok_spmm and ok_km
        assert spans[0][0] == 0 and spans[0][1] == spans[1][0] and spans[1][1] == 500  # shards tile the rows
        total = spans[0][2] + spans[1][2]
        assert abs(spans[0][2] - total / 2) <= 40  # nnz-balanced to within one row
    assert res[0][3] == (0, 19) and res[1][3] == (19, 37)


@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_shard_rules_tile(bof, world):
    from bof_b200 import dist as bdist
    ia = np.concatenate([[0], np.cumsum(np.random.default_rng(1).integers(0, 9, 1000))]).astype(np.int64)
    prev = 0
    for r in range(world):
        a, b = bdist.nnz_balanced_shard(ia, world, r)
        assert a == prev and b >= a
        prev = b
    assert prev == 1000
    prev = 0
    for r in range(world):
        a, b = bdist.row_shard(1001, world, r)
        assert a == prev and b - a in (1001 // world, 1001 // world + 1)
        prev = b
    assert prev == 1001
    # degenerate: empty matrix
    assert bdist.nnz_balanced_shard(np.zeros(1, np.int64), world, 0) == (0, 0)
