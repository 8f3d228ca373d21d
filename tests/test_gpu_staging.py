"""GPU parity of the pageable (file-mapping) path under stress: tiny pinned staging buffers and many row blocks, so
that every pipeline cycles its buffer generations several times while the background download queue is busy
(tickets + fences of blas-on-flash_b200/csrc/staging.cu).  numpy arrays are pageable host memory, like the mmap
behind a flash_ptr."""
import numpy as np
import pytest

import oracle
from gpu_util import ragged_csr

pytestmark = pytest.mark.gpu
TOL = 1e-5


@pytest.fixture(scope="module")
def sctx(bof):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    c = bof.Context(device=0, stage_bytes=1 << 20, n_stage_bufs=2, n_copy_threads=3, gemm_row_block=256,
                    csrmm_max_nnz=20000)
    yield c
    c.close()


@pytest.mark.parametrize("beta", [0.0, 0.5])
@pytest.mark.parametrize("ord_,ta,tb", [("R", "N", "N"), ("C", "T", "N")])
def test_gemm_pageable_many_blocks(sctx, beta, ord_, ta, tb):
    rng = np.random.default_rng(41)
    M, N, K = 3000, 1500, 700           # 12 row blocks of 256 through a ring of 5; a C block is 1.5 MB = 2 chunks
    A = rng.random((M, K), dtype=np.float32); B = rng.random((K, N), dtype=np.float32)
    C0 = rng.random((M, N), dtype=np.float32)

    def store(X, t):
        X = X.T if t == "T" else X
        return np.ascontiguousarray(X if ord_ == "R" else X.T)

    a, b = store(A, ta), store(B, tb)
    for rep in range(2):                # a second call reuses rings, slots and the drainer thread
        c = store(C0, "N").copy() if beta else np.full_like(store(C0, "N"), np.nan)  # copy: store() may alias C0
        sctx.host_gemm(ord_, ta, tb, M, N, K, 1.25, beta, a, b, c)
        got = c if ord_ == "R" else c.T
        ref = oracle.gemm("R", "N", "N", M, N, K, 1.25, beta, A, B, C0, acc64=True)
        assert oracle.rel_fro(got, ref) <= TOL


@pytest.mark.parametrize("beta", [0.0, 0.5])
@pytest.mark.parametrize("ord_b", ["R", "C"])
def test_csrmm_pageable_many_blocks(sctx, beta, ord_b):
    rng = np.random.default_rng(42)
    m, n, k = 20000, 5000, 256          # ~30 row blocks of 20000 nnz; a C block is ~0.7 MB
    a, ia, ja = ragged_csr(rng, m, n, 60)
    B = rng.random((n, k), dtype=np.float32); C0 = rng.random((m, k), dtype=np.float32)
    b = B if ord_b == "R" else np.ascontiguousarray(B.T)
    c = (C0 if ord_b == "R" else np.ascontiguousarray(C0.T)).copy()
    if beta == 0.0:
        c[:] = np.nan
    sctx.host_csrmm("N", m, n, k, 2.0, beta, a, ia, ja, ord_b, b, c)
    got = c if ord_b == "R" else c.T
    assert oracle.rel_fro(got, oracle.csrmm("N", m, n, k, 2.0, beta, a, ia, ja, "R", B, C0, acc64=True)) <= TOL


def test_resident_and_transpose_pageable(bof, sctx):
    rng = np.random.default_rng(43)
    m, n, k = 9000, 7000, 192
    a, ia, ja = ragged_csr(rng, m, n, 40)
    h = bof.ResidentCsr(sctx, m, n, a, ia, ja)
    try:
        for trans, rows_b, rows_c in (("N", n, m), ("T", m, n)):
            B = rng.random((rows_b, k), dtype=np.float32); C0 = rng.random((rows_c, k), dtype=np.float32)
            c = C0.copy()
            h.mm(trans, k, 1.0, 0.25, "R", B, c)
            assert oracle.rel_fro(c, oracle.csrmm(trans, m, n, k, 1.0, 0.25, a, ia, ja, "R", B, C0, acc64=True)) <= TOL
    finally:
        h.close()
    # csrcsc through the same tiny rings: bit-exact
    ia_t = np.zeros(n + 1, np.int64); ja_t = np.zeros(len(ja), np.int64); a_t = np.zeros(len(a), np.float32)
    sctx.host_csrcsc(m, n, ia, ja, a, ia_t, ja_t, a_t)
    r_ia, r_ja, r_a = oracle.csrcsc(m, n, ia, ja, a)
    assert np.array_equal(ia_t, r_ia) and np.array_equal(ja_t, r_ja) and np.array_equal(a_t.view(np.int32), r_a.view(np.int32))
