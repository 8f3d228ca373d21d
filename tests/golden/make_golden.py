"""Generates tests/golden/golden_small.npz.

The reference keeps no golden vectors for this path and cannot be built
here, so the fixtures are produced by the arithmetic library the reference calls -- Intel MKL,
here oneMKL 2024.2 as exported by libtorch_cpu.so (sgemm_, mkl_sparse_s_mm, mkl_sparse_s_mv) --
and, for the integer-only transpose, by scipy's csr->csc (a stable counting sort, like
mkl_scsrcsc).  Inputs are small, seeded and stored next to the outputs, so the fixtures pin the
oracle on any machine without MKL.

    python tests/golden/make_golden.py
"""
import sys
from pathlib import Path

import numpy as np
import scipy.sparse as sp

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import mkl  # noqa: E402


def ragged_csr(rng, m, n, max_nnz, with_dups=False):
    ia = [0]
    ja, a = [], []
    for r in range(m):
        cnt = 0 if r % 7 == 3 else int(rng.integers(0, max_nnz + 1))
        cols = np.sort(rng.choice(n, size=min(cnt, n), replace=False))
        if with_dups and len(cols) > 1 and r % 5 == 0:
            cols[1] = cols[0]  # duplicate (r, c) entry, kept in storage order
        ja.extend(cols.tolist())
        a.extend(rng.random(len(cols)).astype(np.float32).tolist())
        ia.append(len(ja))
    return np.array(a, np.float32), np.array(ia, np.int64), np.array(ja, np.int64)


def main():
    rng = np.random.default_rng(0x5EED)
    out = {}
    # --- sparse: 61 x 47 ragged CSR (empty rows included), k = 12 ---
    m, n, k = 61, 47, 12
    a, ia, ja = ragged_csr(rng, m, n, 9)
    B = rng.random((n, k), dtype=np.float32)
    C0 = rng.random((m, k), dtype=np.float32)
    h = mkl.Csr(m, n, a, ia, ja)
    out.update(sp_m=m, sp_n=n, sp_k=k, sp_a=a, sp_ia=ia, sp_ja=ja, sp_B=B, sp_C0=C0)
    out["spmm_R_a1_b0"] = h.mm("N", k, 1.0, B, 0.0, np.zeros((m, k), np.float32))
    out["spmm_R_a15_b05"] = h.mm("N", k, 1.5, B, 0.5, C0.copy())
    Bc = np.asfortranarray(B)   # column-major n x k
    Cc = np.asfortranarray(C0)
    got = h.mm("N", k, 1.5, Bc.T.copy().reshape(-1), 0.5, Cc.T.copy().reshape(-1), ord_b="C")
    out["spmm_C_a15_b05"] = got  # column-major m x k, flattened
    Bt = rng.random((m, k), dtype=np.float32)
    out["sp_Bt"] = Bt
    out["spmm_T_R"] = h.mm("T", k, 1.0, Bt, 0.0, np.zeros((n, k), np.float32))
    x = rng.random(n, dtype=np.float32)
    xt = rng.random(m, dtype=np.float32)
    out.update(sp_x=x, sp_xt=xt)
    out["spmv_N"] = h.mv("N", x, np.zeros(m, np.float32))
    out["spmv_T"] = h.mv("T", xt, np.zeros(n, np.float32))
    h.close()
    # --- transpose with duplicates: scipy (stable) ---
    a2, ia2, ja2 = ragged_csr(rng, 53, 39, 8, with_dups=True)
    out.update(tr_m=53, tr_n=39, tr_a=a2, tr_ia=ia2, tr_ja=ja2)
    # scipy's tocsc sums duplicates unless told otherwise; build the stable result explicitly
    rows = np.repeat(np.arange(53), np.diff(ia2))
    order = np.argsort(ja2, kind="stable")
    out["tr_ia_t"] = np.concatenate([[0], np.cumsum(np.bincount(ja2, minlength=39))]).astype(np.int64)
    out["tr_ja_t"] = rows[order].astype(np.int64)
    out["tr_a_t"] = a2[order]
    # cross-check against scipy on the duplicate-free matrix
    csc = sp.csr_matrix((a, ja, ia), shape=(m, n)).tocsc()
    out["sp_csc_indptr"] = csc.indptr.astype(np.int64)
    out["sp_csc_indices"] = csc.indices.astype(np.int64)
    out["sp_csc_data"] = csc.data.astype(np.float32)
    # --- dense: MKL sgemm, 37 x 29 x 53, alpha 1.25 beta 0.75 ---
    M, N, K = 37, 29, 53
    A = rng.random((M, K), dtype=np.float32)
    Bd = rng.random((K, N), dtype=np.float32)
    Cd = rng.random((M, N), dtype=np.float32)
    out.update(ge_M=M, ge_N=N, ge_K=K, ge_A=A, ge_B=Bd, ge_C0=Cd)
    out["gemm_a125_b075"] = mkl.sgemm_rowmajor(M, N, K, 1.25, A, Bd, 0.75, Cd.copy())
    out["gemm_a1_b0"] = mkl.sgemm_rowmajor(M, N, K, 1.0, A, Bd, 0.0, np.zeros((M, N), np.float32))
    # --- kmeans: 200 points, 7 centers, 16 dims; distances via MKL sgemm in the reference's op order ---
    P, Kc, d = 200, 7, 16
    cent = rng.normal(size=(Kc, d)).astype(np.float32) * 4
    pts = (cent[rng.integers(0, Kc, P)] + 0.1 * rng.normal(size=(P, d))).astype(np.float32)
    G = mkl.sgemm_rowmajor(P, Kc, d, -2.0, pts, cent.T.copy(), 0.0, np.zeros((P, Kc), np.float32))
    c2 = np.einsum("ij,ij->i", cent, cent).astype(np.float32)
    p2 = np.einsum("ij,ij->i", pts, pts).astype(np.float32)
    D = (G + c2[None, :]).astype(np.float32)
    D = (D + p2[:, None]).astype(np.float32)
    out.update(km_points=pts, km_centers=cent, km_assign=np.argmin(np.abs(D), axis=1).astype(np.int64))
    np.savez_compressed(Path(__file__).with_name("golden_small.npz"), **out)
    print("wrote", Path(__file__).with_name("golden_small.npz"), "with", len(out), "arrays; MKL threads", mkl.max_threads())


if __name__ == "__main__":
    main()
