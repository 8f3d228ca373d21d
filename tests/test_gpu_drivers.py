"""GPU parity through the C++ drop-in layer: the drivers built from drivers/*.cpp use include/flash_blas.h
(reference signatures) on file-backed flash_ptrs with the reference's positional CLIs; outputs are diffed
against the oracle exactly as the reference's in_mem_X / X driver pairs are meant to be diffed."""
import subprocess
from pathlib import Path

import numpy as np
import pytest

import oracle
from gpu_util import ragged_csr

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent
BIN = ROOT / "build"
TOL = 1e-5


def run(exe, *args):
    r = subprocess.run([str(BIN / exe), *map(str, args)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, f"{exe} failed:\n{r.stdout}\n{r.stderr}"
    return r.stdout


@pytest.fixture(scope="module", autouse=True)
def built():
    import __graft_entry__ as g
    if not (BIN / "gemm").exists():
        g.build()


@pytest.mark.parametrize("ta,tb,ord_", [("N", "N", "R"), ("T", "N", "C"), ("N", "T", "R")])
def test_gemm_driver(tmp_path, ta, tb, ord_):
    rng = np.random.default_rng(1)
    M, N, K = 700, 520, 610
    A = rng.random((M, K), dtype=np.float32); B = rng.random((K, N), dtype=np.float32)
    C0 = rng.random((M, N), dtype=np.float32)

    def store(X, t):
        X = X.T if t == "T" else X
        return np.ascontiguousarray(X if ord_ == "R" else X.T)

    store(A, ta).tofile(tmp_path / "A.bin"); store(B, tb).tofile(tmp_path / "B.bin"); store(C0, "N").tofile(tmp_path / "C.bin")
    out = run("gemm", tmp_path / "A.bin", tmp_path / "B.bin", tmp_path / "C.bin", M, K, N, 1.0, 0.5, ta, tb, ord_, 0, 0, 0)
    assert "returned 0" in out
    got = np.fromfile(tmp_path / "C.bin", dtype=np.float32)
    got = got.reshape(M, N) if ord_ == "R" else got.reshape(N, M).T
    ref = oracle.gemm("R", "N", "N", M, N, K, 1.0, 0.5, A, B, C0, acc64=True)
    assert oracle.rel_fro(got, ref) <= TOL


def write_csr(tmp_path, a, ia, ja):
    a.tofile(tmp_path / "A.csr"); ja.tofile(tmp_path / "A.col"); ia.tofile(tmp_path / "A.off")  # misc/sparse_create.cpp:27-29
    return tmp_path / "A.csr", tmp_path / "A.col", tmp_path / "A.off"


def test_csrmm_and_csrgemv_drivers(tmp_path):
    rng = np.random.default_rng(2)
    m, n, k = 2500, 1800, 64
    a, ia, ja = ragged_csr(rng, m, n, 45)
    fa, fj, fi = write_csr(tmp_path, a, ia, ja)
    B = rng.random((n, k), dtype=np.float32); C0 = rng.random((m, k), dtype=np.float32)
    B.tofile(tmp_path / "B.bin"); C0.tofile(tmp_path / "C.bin")
    run("csrmm", fa, fj, fi, tmp_path / "B.bin", tmp_path / "C.bin", m, n, k, 2.0, 0.25, "N", "R")
    got = np.fromfile(tmp_path / "C.bin", dtype=np.float32).reshape(m, k)
    assert oracle.rel_fro(got, oracle.csrmm("N", m, n, k, 2.0, 0.25, a, ia, ja, "R", B, C0, acc64=True)) <= TOL
    for trans in "NT":
        x = rng.random(n if trans == "N" else m, dtype=np.float32)
        x.tofile(tmp_path / "x.bin")
        run("csrgemv", fa, fj, fi, tmp_path / "x.bin", tmp_path / "y.bin", m, n, trans)
        y = np.fromfile(tmp_path / "y.bin", dtype=np.float32)
        assert oracle.rel_fro(y, oracle.csrgemv(trans, m, n, a, ia, ja, x, acc64=True)) <= TOL


def test_csrcsc_driver_bit_exact(tmp_path):
    rng = np.random.default_rng(3)
    m, n = 3000, 4100
    a, ia, ja = ragged_csr(rng, m, n, 30, dups=True)
    fa, fj, fi = write_csr(tmp_path, a, ia, ja)
    run("csrcsc", fa, fj, fi, tmp_path / "T.csr", tmp_path / "T.col", tmp_path / "T.off", m, n)
    r_ia, r_ja, r_a = oracle.csrcsc(m, n, ia, ja, a)
    assert np.array_equal(np.fromfile(tmp_path / "T.off", dtype=np.int64), r_ia)
    assert np.array_equal(np.fromfile(tmp_path / "T.col", dtype=np.int64)[: r_ja.size], r_ja)
    assert np.array_equal(np.fromfile(tmp_path / "T.csr", dtype=np.uint32)[: r_a.size], r_a.view(np.uint32))


def test_kmeans_driver_one_iteration(tmp_path):
    """drivers/kmeans.cpp / in_mem_kmeans.cpp run ONE Lloyd iteration and leave the centers in the centers file."""
    rng = np.random.default_rng(4)
    P, K, d = 20000, 32, 24
    cent = (rng.normal(size=(K, d)) * 5).astype(np.float32)
    pts = (cent[rng.integers(0, K, P)] + 0.2 * rng.normal(size=(P, d))).astype(np.float32)
    c0 = pts[:K].copy()
    pts.tofile(tmp_path / "points.bin"); c0.tofile(tmp_path / "centers.bin")
    run("kmeans", tmp_path / "points.bin", tmp_path / "centers.bin", P, d, K)
    got = np.fromfile(tmp_path / "centers.bin", dtype=np.float32).reshape(K, d)
    ref_c, ref_a, _ = oracle.lloyd_iter(pts, c0)
    _, margin = oracle.kmeans_assign(pts, c0)
    if (margin > 1e-3 * (1 + np.einsum("ij,ij->i", pts, pts))).all():
        assert oracle.rel_fro(got, ref_c) <= TOL
    else:  # a near-tie may move one point between clusters; centroids still agree closely
        assert oracle.rel_fro(got, ref_c) <= 1e-3


@pytest.mark.parametrize("trans,n_calls", [("N", 1), ("N", 3), ("T", 2)])
def test_csrmm_pmem_driver(tmp_path, trans, n_calls):
    """drivers/csrmm_pmem.cpp: B, C in host memory (flash_blas.h:43-46); n_calls > 1 pins A in HBM (flash::csr_pin)."""
    rng = np.random.default_rng(12)
    m, n, k = 2100, 1700, 160
    a, ia, ja = ragged_csr(rng, m, n, 35)
    fa, fj, fi = write_csr(tmp_path, a, ia, ja)
    rows_b, rows_c = (n, m) if trans == "N" else (m, n)
    B = rng.random((rows_b, k), dtype=np.float32); C0 = rng.random((rows_c, k), dtype=np.float32)
    B.tofile(tmp_path / "B.bin"); C0.tofile(tmp_path / "C.bin")
    args = [fa, fj, fi, tmp_path / "B.bin", tmp_path / "C.bin", m, n, k, 1.25, 0.5, trans, "R"]
    out = run("csrmm_pmem", *args, *([n_calls] if n_calls > 1 else []))
    assert out.count("returned 0") == n_calls + (1 if n_calls > 1 else 0)
    got = np.fromfile(tmp_path / "C.bin", dtype=np.float32).reshape(rows_c, k)
    ref = oracle.csrmm(trans, m, n, k, 1.25, 0.5, a, ia, ja, "R", B, C0, acc64=True)
    assert oracle.rel_fro(got, ref) <= TOL
