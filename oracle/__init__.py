"""Python face of the CPU oracle (``oracle.c``).  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package; the product path never does.
PARITY PINNED TO THE REFERENCE: tests/golden/golden_ref.npz holds outputs written by the reference's own
binaries (``oracle/_ref``, unmodified sources built by ``oracle/Makefile.ref``); ``tests/test_oracle_ref.py``
checks every function here against them and against live runs of those binaries (``oracle/ref_run.py``).
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

_DIR = Path(__file__).resolve().parent
_LIB = _DIR / "liboracle.so"
_lib = None

_p, _i64, _f32, _ch, _int = C.c_void_p, C.c_int64, C.c_float, C.c_char, C.c_int


def build(force: bool = False) -> Path:
    src = _DIR / "oracle.c"
    if force or not _LIB.exists() or _LIB.stat().st_mtime < src.stat().st_mtime:
        subprocess.run(["make", "-C", str(_DIR), "liboracle.so"], check=True, capture_output=True)
    return _LIB


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not _LIB.exists():
            build()
        L = C.CDLL(str(_LIB))
        sig = {
            "orc_csrmm": (_int, [_ch, _i64, _i64, _i64, _f32, _f32, _p, _p, _p, _ch, _p, _p, _int]),
            "orc_gemm": (_int, [_ch, _ch, _ch, _i64, _i64, _i64, _f32, _f32, _p, _i64, _p, _i64, _p, _i64, _int]),
            "orc_gemm_tiled": (_int, [_ch, _ch, _ch, _i64, _i64, _i64, _f32, _f32, _p, _i64, _p, _i64, _p, _i64,
                                      _i64]),
            "orc_csrgemv": (_int, [_ch, _i64, _i64, _p, _p, _p, _p, _p, _int]),
            "orc_csrcsc": (_int, [_i64, _i64, _p, _p, _p, _p, _p, _p]),
            "orc_csrcsc_blocked": (_int, [_i64, _i64, _p, _p, _p, _p, _p, _p, _i64, _i64]),
            "orc_next_blk_size": (_i64, [_p, _i64, _i64, _i64, _i64]),
            "orc_row_sqnorm": (None, [_i64, _i64, _p, _p]),
            "orc_kmeans_assign": (None, [_i64, _i64, _i64, _p, _p, _p, _p, _p, _p, _int]),
            "orc_kmeans_update": (None, [_i64, _i64, _i64, _p, _p, _p, _p, _int]),
            "orc_kmeans_residual": (C.c_double, [_i64, _i64, _p, _p, _p]),
            "orc_lloyd_iter": (C.c_double, [_i64, _i64, _i64, _p, _p, _p, _p, _int]),
            "orc_gen_csr": (_int, [_i64, _i64, _i64, C.c_uint64, _int, _p, _p, _p]),
            "orc_gen_dense": (None, [_i64, C.c_uint64, _int, _p]),
        }
        for name, (res, args) in sig.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


def _c(c: str) -> bytes:
    return c.encode("ascii")[:1]


def _a(x, dt):
    x = np.ascontiguousarray(x, dtype=dt)
    return x


def _ptr(x):
    return None if x is None else x.ctypes.data


def csrmm(trans, m, n, k, alpha, beta, a, ia, ja, ord_b, b, c, acc64=False):
    """C updated in place (a copy is returned); layouts as drivers/in_mem_csrmm.cpp."""
    a, ia, ja, b = _a(a, np.float32), _a(ia, np.int64), _a(ja, np.int64), _a(b, np.float32)
    c = np.array(c, dtype=np.float32, copy=True)
    rc = lib().orc_csrmm(_c(trans), m, n, k, alpha, beta, _ptr(a), _ptr(ia), _ptr(ja), _c(ord_b), _ptr(b), _ptr(c),
                         int(acc64))
    if rc:
        raise ValueError(f"orc_csrmm returned {rc}")
    return c


def gemm(ord_, ta, tb, m, n, k, alpha, beta, a, b, c, lda=0, ldb=0, ldc=0, acc64=False, tiled_blk=0):
    a, b = _a(a, np.float32), _a(b, np.float32)
    c = np.array(c, dtype=np.float32, copy=True)
    if tiled_blk:
        rc = lib().orc_gemm_tiled(_c(ord_), _c(ta), _c(tb), m, n, k, alpha, beta, _ptr(a), lda, _ptr(b), ldb,
                                  _ptr(c), ldc, tiled_blk)
    else:
        rc = lib().orc_gemm(_c(ord_), _c(ta), _c(tb), m, n, k, alpha, beta, _ptr(a), lda, _ptr(b), ldb, _ptr(c),
                            ldc, int(acc64))
    if rc:
        raise ValueError(f"orc_gemm returned {rc}")
    return c


def csrgemv(trans, m, n, a, ia, ja, x, acc64=False):
    a, ia, ja, x = _a(a, np.float32), _a(ia, np.int64), _a(ja, np.int64), _a(x, np.float32)
    y = np.zeros(m if trans == "N" else n, dtype=np.float32)
    rc = lib().orc_csrgemv(_c(trans), m, n, _ptr(a), _ptr(ia), _ptr(ja), _ptr(x), _ptr(y), int(acc64))
    if rc:
        raise ValueError(f"orc_csrgemv returned {rc}")
    return y


def csrcsc(m, n, ia, ja, a, blocked_rblk=0, max_nnzs=10_000_000):
    a, ia, ja = _a(a, np.float32), _a(ia, np.int64), _a(ja, np.int64)
    nnz = int(ia[m] - ia[0])
    ia_tr = np.zeros(n + 1, dtype=np.int64)
    ja_tr = np.zeros(max(nnz, 1), dtype=np.int64)
    a_tr = np.zeros(max(nnz, 1), dtype=np.float32)
    if blocked_rblk:
        rc = lib().orc_csrcsc_blocked(m, n, _ptr(ia), _ptr(ja), _ptr(a), _ptr(ia_tr), _ptr(ja_tr), _ptr(a_tr),
                                      blocked_rblk, max_nnzs)
    else:
        rc = lib().orc_csrcsc(m, n, _ptr(ia), _ptr(ja), _ptr(a), _ptr(ia_tr), _ptr(ja_tr), _ptr(a_tr))
    if rc:
        raise ValueError(f"orc_csrcsc returned {rc}")
    return ia_tr, ja_tr[:nnz], a_tr[:nnz]


def next_blk_size(offs, nrows, min_size, max_size, max_nnzs=10_000_000):
    offs = _a(offs, np.int64)
    return int(lib().orc_next_blk_size(_ptr(offs), nrows, min_size, max_size, max_nnzs))


def row_sqnorm(x):
    x = _a(x, np.float32)
    out = np.zeros(x.shape[0], dtype=np.float32)
    lib().orc_row_sqnorm(x.shape[0], x.shape[1], _ptr(x), _ptr(out))
    return out


def kmeans_assign(points, centers, c_l2sq=None, p_l2sq=None, acc64=False):
    points, centers = _a(points, np.float32), _a(centers, np.float32)
    c2 = row_sqnorm(centers) if c_l2sq is None else _a(c_l2sq, np.float32)
    p2 = row_sqnorm(points) if p_l2sq is None else _a(p_l2sq, np.float32)
    P, K, d = points.shape[0], centers.shape[0], points.shape[1]
    assign = np.zeros(P, dtype=np.int64)
    margin = np.zeros(P, dtype=np.float32)
    lib().orc_kmeans_assign(P, K, d, _ptr(points), _ptr(centers), _ptr(c2), _ptr(p2), _ptr(assign), _ptr(margin),
                            int(acc64))
    return assign, margin


def kmeans_update(points, assign, ncenters, mode=0):
    points, assign = _a(points, np.float32), _a(assign, np.int64)
    P, d = points.shape
    centers = np.zeros((ncenters, d), dtype=np.float32)
    counts = np.zeros(ncenters, dtype=np.int64)
    lib().orc_kmeans_update(P, ncenters, d, _ptr(points), _ptr(assign), _ptr(centers), _ptr(counts), mode)
    return centers, counts


def lloyd_iter(points, centers, p_l2sq=None, update_mode=0):
    points = _a(points, np.float32)
    centers = np.array(centers, dtype=np.float32, copy=True)
    p2 = row_sqnorm(points) if p_l2sq is None else _a(p_l2sq, np.float32)
    P, d = points.shape
    assign = np.zeros(P, dtype=np.int64)
    res = lib().orc_lloyd_iter(P, centers.shape[0], d, _ptr(points), _ptr(centers), _ptr(p2), _ptr(assign),
                               update_mode)
    return centers, assign, float(res)


def gen_csr(m, n, nnz_per_row, seed=0x5EED0001, val_mode=1):
    """(a, ia, ja) with exactly nnz_per_row sorted unique columns per row (file-format dtypes)."""
    ia = np.zeros(m + 1, dtype=np.int64)
    ja = np.zeros(max(m * nnz_per_row, 1), dtype=np.int64)
    a = np.zeros(max(m * nnz_per_row, 1), dtype=np.float32)
    rc = lib().orc_gen_csr(m, n, nnz_per_row, seed, val_mode, _ptr(ia), _ptr(ja), _ptr(a))
    if rc:
        raise ValueError("gen_csr: nnz_per_row > n")
    return a[: m * nnz_per_row], ia, ja[: m * nnz_per_row]


def gen_dense(shape, seed=0x5EED0002, mode=1):
    out = np.zeros(shape, dtype=np.float32)
    lib().orc_gen_dense(out.size, seed, mode, _ptr(out))
    return out


def rel_fro(x, ref) -> float:
    """relative Frobenius error ||x - ref||_F / ||ref||_F in float64 (the tolerance metric of BASELINE.json)."""
    x = np.asarray(x, dtype=np.float64).ravel()
    ref = np.asarray(ref, dtype=np.float64).ravel()
    den = np.linalg.norm(ref)
    return float(np.linalg.norm(x - ref) / (den if den > 0 else 1.0))
