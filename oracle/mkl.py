"""oneMKL 2024.2 as exported by ``libtorch_cpu.so`` -- independent cross-check for the oracle and the
timed CPU baseline.  TEST INFRASTRUCTURE ONLY (same rule as ``oracle/__init__.py``).

The reference calls MKL through ``cblas_sgemm`` / ``mkl_scsrmm`` / ``mkl_cspblas_scsrgemv``
(include/bof_types.h:18-29).  Those deprecated entry points are not exported by the MKL build
inside torch, but the same library's ``sgemm_`` (Fortran interface), ``mkl_sparse_s_mm`` and
``mkl_sparse_s_mv`` are (LP64: 32-bit indices).  They compute the same operations with MKL's own
threading and summation order, which is exactly what the reference's in_mem_* drivers time.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_lib = None

SPARSE_OPERATION_NON_TRANSPOSE = 10
SPARSE_OPERATION_TRANSPOSE = 11
SPARSE_INDEX_BASE_ZERO = 0
SPARSE_MATRIX_TYPE_GENERAL = 20
SPARSE_LAYOUT_ROW_MAJOR = 101
SPARSE_LAYOUT_COLUMN_MAJOR = 102


class MatrixDescr(C.Structure):
    _fields_ = [("type", C.c_int), ("mode", C.c_int), ("diag", C.c_int)]


def lib():
    global _lib
    if _lib is None:
        import torch

        path = os.path.join(os.path.dirname(torch.__file__), "lib", "libtorch_cpu.so")
        L = C.CDLL(path)
        L.sgemm_.restype = None
        L.mkl_sparse_s_create_csr.restype = C.c_int
        L.mkl_sparse_s_create_csr.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int, C.c_void_p,
                                              C.c_void_p, C.c_void_p, C.c_void_p]
        L.mkl_sparse_s_mm.restype = C.c_int
        L.mkl_sparse_s_mm.argtypes = [C.c_int, C.c_float, C.c_void_p, MatrixDescr, C.c_int, C.c_void_p, C.c_int,
                                      C.c_int, C.c_float, C.c_void_p, C.c_int]
        L.mkl_sparse_s_mv.restype = C.c_int
        L.mkl_sparse_s_mv.argtypes = [C.c_int, C.c_float, C.c_void_p, MatrixDescr, C.c_void_p, C.c_float,
                                      C.c_void_p]
        L.mkl_sparse_destroy.restype = C.c_int
        L.mkl_sparse_destroy.argtypes = [C.c_void_p]
        L.mkl_get_max_threads.restype = C.c_int
        _lib = L
    return _lib


def max_threads() -> int:
    return int(lib().mkl_get_max_threads())


def sgemm_rowmajor(m, n, k, alpha, a, b, beta, c):
    """C(m x n) = alpha*A(m x k)*B(k x n) + beta*C, all row-major tight, through Fortran sgemm_
    using the identity C^T = B^T A^T (what drivers/in_mem_gemm.cpp:58-67 computes with
    CblasRowMajor).  c is updated in place."""
    L = lib()
    ci = lambda v: C.byref(C.c_int(v))
    cf = lambda v: C.byref(C.c_float(v))
    N = C.c_char(b"N")
    L.sgemm_(C.byref(N), C.byref(N), ci(n), ci(m), ci(k), cf(alpha), C.c_void_p(b.ctypes.data), ci(n),
             C.c_void_p(a.ctypes.data), ci(k), cf(beta), C.c_void_p(c.ctypes.data), ci(n))
    return c


class Csr:
    """MKL sparse handle over int32 copies of (ia, ja); 0-based."""

    def __init__(self, m, n, a, ia, ja):
        self.a = np.ascontiguousarray(a, dtype=np.float32)
        ia = np.asarray(ia, dtype=np.int64)
        self.ia = np.ascontiguousarray(ia - ia[0], dtype=np.int32)
        self.ja = np.ascontiguousarray(ja, dtype=np.int32)
        self.m, self.n = m, n
        self.h = C.c_void_p()
        st = lib().mkl_sparse_s_create_csr(C.byref(self.h), SPARSE_INDEX_BASE_ZERO, m, n, self.ia.ctypes.data,
                                           self.ia.ctypes.data + 4, self.ja.ctypes.data, self.a.ctypes.data)
        if st != 0:
            raise RuntimeError(f"mkl_sparse_s_create_csr status {st}")
        self.descr = MatrixDescr(SPARSE_MATRIX_TYPE_GENERAL, 0, 0)

    def mm(self, trans, k, alpha, b, beta, c, ord_b="R"):
        """c updated in place; layouts as drivers/in_mem_csrmm.cpp:96-120."""
        op = SPARSE_OPERATION_TRANSPOSE if trans == "T" else SPARSE_OPERATION_NON_TRANSPOSE
        brows = self.m if trans == "T" else self.n
        crows = self.n if trans == "T" else self.m
        if ord_b == "R":
            layout, ldb, ldc = SPARSE_LAYOUT_ROW_MAJOR, k, k
        else:
            layout, ldb, ldc = SPARSE_LAYOUT_COLUMN_MAJOR, brows, crows
        st = lib().mkl_sparse_s_mm(op, alpha, self.h, self.descr, layout, b.ctypes.data, k, ldb, beta,
                                   c.ctypes.data, ldc)
        if st != 0:
            raise RuntimeError(f"mkl_sparse_s_mm status {st}")
        return c

    def mv(self, trans, x, y):
        op = SPARSE_OPERATION_TRANSPOSE if trans == "T" else SPARSE_OPERATION_NON_TRANSPOSE
        st = lib().mkl_sparse_s_mv(op, 1.0, self.h, self.descr, x.ctypes.data, 0.0, y.ctypes.data)
        if st != 0:
            raise RuntimeError(f"mkl_sparse_s_mv status {st}")
        return y

    def close(self):
        if self.h:
            lib().mkl_sparse_destroy(self.h)
            self.h = None
