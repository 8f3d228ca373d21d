"""CPU tests: the oracle (oracle/oracle.c) against the committed golden vectors (MKL 2024.2 / scipy
outputs, tests/golden/make_golden.py), against live MKL where libtorch exports it, and against the
structural properties of the reference's tilers."""
from pathlib import Path

import numpy as np
import pytest

import oracle

G = np.load(Path(__file__).parent / "golden" / "golden_small.npz")
TOL = 1e-5  # relative Frobenius, BASELINE.json


def test_csrmm_rowmajor_vs_golden():
    m, n, k = int(G["sp_m"]), int(G["sp_n"]), int(G["sp_k"])
    a, ia, ja, B, C0 = G["sp_a"], G["sp_ia"], G["sp_ja"], G["sp_B"], G["sp_C0"]
    for acc64 in (False, True):
        c = oracle.csrmm("N", m, n, k, 1.0, 0.0, a, ia, ja, "R", B, np.full((m, k), np.nan, np.float32), acc64)
        assert oracle.rel_fro(c, G["spmm_R_a1_b0"]) < 1e-6  # beta == 0: NaNs in C must not propagate
        c = oracle.csrmm("N", m, n, k, 1.5, 0.5, a, ia, ja, "R", B, C0, acc64)
        assert oracle.rel_fro(c, G["spmm_R_a15_b05"]) < 1e-6


def test_csrmm_colmajor_and_trans_vs_golden():
    m, n, k = int(G["sp_m"]), int(G["sp_n"]), int(G["sp_k"])
    a, ia, ja, B, C0 = G["sp_a"], G["sp_ia"], G["sp_ja"], G["sp_B"], G["sp_C0"]
    Bc = np.asfortranarray(B).T.copy().reshape(-1)
    Cc = np.asfortranarray(C0).T.copy().reshape(-1)
    c = oracle.csrmm("N", m, n, k, 1.5, 0.5, a, ia, ja, "C", Bc, Cc)
    assert oracle.rel_fro(c, G["spmm_C_a15_b05"]) < 1e-6
    c = oracle.csrmm("T", m, n, k, 1.0, 0.0, a, ia, ja, "R", G["sp_Bt"], np.zeros((n, k), np.float32))
    assert oracle.rel_fro(c, G["spmm_T_R"]) < 1e-6


def test_csrmm_bad_args():
    with pytest.raises(ValueError):
        oracle.csrmm("X", 1, 1, 1, 1.0, 0.0, np.ones(1), np.array([0, 1]), np.array([0]), "R", np.ones(1), np.ones(1))
    with pytest.raises(ValueError):
        oracle.csrmm("N", 1, 1, 1, 1.0, 0.0, np.ones(1), np.array([0, 1]), np.array([0]), "Q", np.ones(1), np.ones(1))


def test_csrmm_unrebased_offsets():
    """SparseBlock slices keep file offsets (src/blas/csrmm.cpp:87-103 rebases; ours indexes by ia[i]-ia[0])."""
    m, n, k = int(G["sp_m"]), int(G["sp_n"]), int(G["sp_k"])
    a, ia, ja, B = G["sp_a"], G["sp_ia"], G["sp_ja"], G["sp_B"]
    r0, r1 = 10, 40
    z0, z1 = ia[r0], ia[r1]
    c = oracle.csrmm("N", r1 - r0, n, k, 1.0, 0.0, a[z0:z1], ia[r0:r1 + 1], ja[z0:z1], "R", B,
                     np.zeros((r1 - r0, k), np.float32))
    assert oracle.rel_fro(c, G["spmm_R_a1_b0"][r0:r1]) < 1e-6


def test_csrgemv_vs_golden():
    m, n = int(G["sp_m"]), int(G["sp_n"])
    a, ia, ja = G["sp_a"], G["sp_ia"], G["sp_ja"]
    assert oracle.rel_fro(oracle.csrgemv("N", m, n, a, ia, ja, G["sp_x"]), G["spmv_N"]) < 1e-6
    assert oracle.rel_fro(oracle.csrgemv("T", m, n, a, ia, ja, G["sp_xt"]), G["spmv_T"]) < 1e-6
    assert oracle.rel_fro(oracle.csrgemv("T", m, n, a, ia, ja, G["sp_xt"], acc64=True), G["spmv_T"]) < 1e-6


def test_csrcsc_vs_golden_bit_exact():
    ia_t, ja_t, a_t = oracle.csrcsc(int(G["tr_m"]), int(G["tr_n"]), G["tr_ia"], G["tr_ja"], G["tr_a"])
    assert np.array_equal(ia_t, G["tr_ia_t"]) and np.array_equal(ja_t, G["tr_ja_t"])
    assert np.array_equal(a_t.view(np.uint32), G["tr_a_t"].view(np.uint32))
    ia_t, ja_t, a_t = oracle.csrcsc(int(G["sp_m"]), int(G["sp_n"]), G["sp_ia"], G["sp_ja"], G["sp_a"])
    assert np.array_equal(ia_t, G["sp_csc_indptr"]) and np.array_equal(ja_t, G["sp_csc_indices"])
    assert np.array_equal(a_t.view(np.uint32), G["sp_csc_data"].view(np.uint32))


@pytest.mark.parametrize("rblk", [10, 13, 64])
def test_csrcsc_blocked_reference_algorithm_equals_stable_sort(rblk):
    """Row-block transposes + in-order merge (src/blas/csrcsc.cpp, csrcsc_task.h) == stable counting sort."""
    a, ia, ja = oracle.gen_csr(300, 257, 9, seed=7)
    plain = oracle.csrcsc(300, 257, ia, ja, a)
    blocked = oracle.csrcsc(300, 257, ia, ja, a, blocked_rblk=rblk, max_nnzs=100)
    for x, y in zip(plain, blocked):
        assert np.array_equal(x, y)


def test_csrcsc_edge_cases():
    ia_t, ja_t, a_t = oracle.csrcsc(5, 4, np.zeros(6, np.int64), np.zeros(0, np.int64), np.zeros(0, np.float32))
    assert np.array_equal(ia_t, np.zeros(5, np.int64)) and ja_t.size == 0
    # transposing twice is the identity for a duplicate-free sorted CSR
    a, ia, ja = oracle.gen_csr(64, 80, 5, seed=3)
    t = oracle.csrcsc(64, 80, ia, ja, a)
    tt = oracle.csrcsc(80, 64, *t)
    assert np.array_equal(tt[0], ia) and np.array_equal(tt[1], ja) and np.array_equal(tt[2], a)


def test_next_blk_size_rule():
    """include/blas_utils.h:72-82."""
    offs = np.arange(0, 1001 * 10, 10, dtype=np.int64)  # 10 nnz per row, 1000 rows
    assert oracle.next_blk_size(offs, 1000, 128, 131072, max_nnzs=5000) == 501  # first blk with nnz > 5000
    assert oracle.next_blk_size(offs, 1000, 128, 256, max_nnzs=5000) == 256     # capped
    assert oracle.next_blk_size(offs, 1000, 128, 131072, max_nnzs=10**9) == 1000
    assert oracle.next_blk_size(offs, 50, 128, 131072, max_nnzs=10**9) == 128   # min_size wins (reference quirk)


def test_gemm_vs_golden():
    M, N, K = int(G["ge_M"]), int(G["ge_N"]), int(G["ge_K"])
    A, B, C0 = G["ge_A"], G["ge_B"], G["ge_C0"]
    c = oracle.gemm("R", "N", "N", M, N, K, 1.25, 0.75, A, B, C0)
    assert oracle.rel_fro(c, G["gemm_a125_b075"]) < 1e-6
    c = oracle.gemm("R", "N", "N", M, N, K, 1.0, 0.0, A, B, np.full((M, N), np.nan, np.float32))
    assert oracle.rel_fro(c, G["gemm_a1_b0"]) < 1e-6


@pytest.mark.parametrize("ord_", "RC")
@pytest.mark.parametrize("ta", "NT")
@pytest.mark.parametrize("tb", "NT")
def test_gemm_all_layouts_vs_numpy(ord_, ta, tb):
    """The 8 (order, transA, transB) configurations of misc/gemm_run.sh:31-38."""
    rng = np.random.default_rng(1)
    M, N, K = 33, 21, 45
    A = rng.random((M, K), dtype=np.float32)
    B = rng.random((K, N), dtype=np.float32)
    C0 = rng.random((M, N), dtype=np.float32)
    ref = (1.5 * (A.astype(np.float64) @ B.astype(np.float64)) + 0.5 * C0).astype(np.float32)

    def store(X, trans):  # how the caller lays the operand out
        X = X.T if trans == "T" else X
        return np.ascontiguousarray(X if ord_ == "R" else X.T).reshape(-1)

    c = oracle.gemm(ord_, ta, tb, M, N, K, 1.5, 0.5, store(A, ta), store(B, tb), store(C0, "N"))
    c = c.reshape(M, N) if ord_ == "R" else c.reshape(N, M).T
    assert oracle.rel_fro(c, ref) < 1e-6
    # explicit leading dimensions larger than tight
    if ord_ == "R" and ta == "N" and tb == "N":
        Ap = np.zeros((M, K + 3), np.float32); Ap[:, :K] = A
        Bp = np.zeros((K, N + 5), np.float32); Bp[:, :N] = B
        Cp = np.zeros((M, N + 2), np.float32); Cp[:, :N] = C0
        c = oracle.gemm("R", "N", "N", M, N, K, 1.5, 0.5, Ap, Bp, Cp, K + 3, N + 5, N + 2).reshape(M, N + 2)
        assert oracle.rel_fro(c[:, :N], ref) < 1e-6 and np.all(c[:, N:] == 0)


def test_gemm_tiler_chain_matches_monolithic():
    """src/blas/gemm.cpp:46-129: blocks + tail merge + beta=1 accumulate chain."""
    rng = np.random.default_rng(2)
    M, N, K = 300, 260, 530  # with blk=128: tails of 44 (<128: merged), 4 (merged), 18 (merged)
    A = rng.random((M, K), dtype=np.float32)
    B = rng.random((K, N), dtype=np.float32)
    C0 = rng.random((M, N), dtype=np.float32)
    mono = oracle.gemm("R", "N", "N", M, N, K, 1.0, 0.5, A, B, C0, acc64=True)
    tiled = oracle.gemm("R", "N", "N", M, N, K, 1.0, 0.5, A, B, C0, tiled_blk=128)
    assert oracle.rel_fro(tiled, mono) < TOL
    tiled = oracle.gemm("C", "T", "N", M, N, K, 1.0, 0.5, A.reshape(-1), B.T.copy().reshape(-1),
                        C0.T.copy().reshape(-1), tiled_blk=128)
    assert oracle.rel_fro(tiled.reshape(N, M).T, mono) < TOL


def test_integer_compat_data_is_exact():
    """misc/sparse_create.cpp:52-55 and misc/dense_create.cpp:28-32 emit small integers: every fp32
    summation order gives the same bits."""
    a, ia, ja = oracle.gen_csr(128, 96, 7, seed=5, val_mode=0)
    assert set(np.unique(a)) <= set(range(1, 10))
    B = oracle.gen_dense((96, 8), mode=0)
    c32 = oracle.csrmm("N", 128, 96, 8, 1.0, 0.0, a, ia, ja, "R", B, np.zeros((128, 8), np.float32))
    c64 = oracle.csrmm("N", 128, 96, 8, 1.0, 0.0, a, ia, ja, "R", B, np.zeros((128, 8), np.float32), acc64=True)
    assert np.array_equal(c32, c64)


def test_gen_csr_structure():
    a, ia, ja = oracle.gen_csr(200, 50, 20, seed=9)
    assert np.array_equal(np.diff(ia), np.full(200, 20))
    cols = ja.reshape(200, 20)
    assert np.all(np.diff(cols, axis=1) > 0) and cols.min() >= 0 and cols.max() < 50
    assert a.min() >= 0 and a.max() < 1


def test_kmeans_assign_vs_golden_and_semantics():
    pts, cent = G["km_points"], G["km_centers"]
    assign, margin = oracle.kmeans_assign(pts, cent)
    ok = margin > 1e-3  # ties-free points only: the dot product's rounding is MKL-internal
    assert ok.mean() > 0.99
    assert np.array_equal(assign[ok], G["km_assign"][ok])
    # isamin = first index of the minimum ABSOLUTE value (drivers/in_mem_kmeans.cpp:84-85)
    p = np.zeros((1, 2), np.float32)
    c = np.array([[1, 0], [1, 0], [0.5, 0]], np.float32)
    a2, _ = oracle.kmeans_assign(p, c)
    assert a2[0] == 2
    c = np.array([[1, 0], [1, 0]], np.float32)
    a2, _ = oracle.kmeans_assign(p, c)
    assert a2[0] == 0  # tie -> first


def test_kmeans_update_modes_and_empty_cluster():
    rng = np.random.default_rng(3)
    pts = rng.normal(size=(500, 8)).astype(np.float32)
    assign = rng.integers(0, 5, 500).astype(np.int64)
    assign[assign == 3] = 1  # cluster 3 empty
    ref, counts = oracle.kmeans_update(pts, assign, 6, mode=0)
    assert counts[3] == 0 and counts[5] == 0 and np.all(ref[3] == 0) and np.all(ref[5] == 0)
    for mode in (1, 2):
        c, _ = oracle.kmeans_update(pts, assign, 6, mode=mode)
        assert oracle.rel_fro(c, ref) < TOL
    want = np.stack([pts[assign == c].mean(axis=0) if (assign == c).any() else np.zeros(8) for c in range(6)])
    assert oracle.rel_fro(ref, want) < TOL


def test_lloyd_iter_decreases_residual():
    rng = np.random.default_rng(4)
    cent = rng.normal(size=(4, 6)).astype(np.float32) * 5
    pts = (cent[rng.integers(0, 4, 400)] + rng.normal(size=(400, 6))).astype(np.float32)
    c0 = pts[:4].copy()
    c1, a1, r1 = oracle.lloyd_iter(pts, c0)
    c2, a2, r2 = oracle.lloyd_iter(pts, c1)
    assert r2 <= r1 * (1 + 1e-6)


def test_oracle_vs_live_mkl():
    """Independent cross-check at a size the fixtures do not cover (skips if libtorch lacks MKL)."""
    try:
        from oracle import mkl
        mkl.lib()
    except Exception as e:  # pragma: no cover
        pytest.skip(f"MKL not reachable: {e}")
    m, n, k = 2000, 1500, 64
    a, ia, ja = oracle.gen_csr(m, n, 30, seed=11)
    B = oracle.gen_dense((n, k), seed=12)
    h = mkl.Csr(m, n, a, ia, ja)
    ref = h.mm("N", k, 1.0, B, 0.0, np.zeros((m, k), np.float32))
    assert oracle.rel_fro(oracle.csrmm("N", m, n, k, 1.0, 0.0, a, ia, ja, "R", B, np.zeros((m, k), np.float32)), ref) < 1e-6
    x = oracle.gen_dense((n,), seed=13)
    assert oracle.rel_fro(oracle.csrgemv("N", m, n, a, ia, ja, x), h.mv("N", x, np.zeros(m, np.float32))) < 1e-6
    h.close()
    M = N = K = 384
    A = oracle.gen_dense((M, K), seed=14)
    Bd = oracle.gen_dense((K, N), seed=15)
    ref = mkl.sgemm_rowmajor(M, N, K, 1.0, A, Bd, 0.0, np.zeros((M, N), np.float32))
    assert oracle.rel_fro(oracle.gemm("R", "N", "N", M, N, K, 1.0, 0.0, A, Bd, np.zeros((M, N), np.float32)), ref) < 1e-6
