"""Generates tests/golden/golden_ref.npz from the REFERENCE ITSELF.

Every output array in the fixture was written by a binary of ``oracle/_ref/`` -- the unmodified sources of
microsoft/BLAS-on-flash compiled by ``oracle/Makefile.ref`` (in_mem_* drivers = the bare MKL call the reference
times itself against; flash drivers = flash::gemm / csrmm / csrgemv / csrcsc / kmeans through the reference's own
scheduler, cache and libaio file handles).  The ``_small`` build only shrinks the tile constants
(-DGEMM_BLK_SIZE=256 ...) so that these kilobyte-sized problems run through several row / column / k blocks,
the beta=1 accumulate chains and the csrcsc merge phase.  Sparse inputs are stored next to the outputs; large dense
inputs are regenerated from the stored seeds with the oracle's counter-based generator (oracle.gen_dense, U[0,1)),
so the fixture pins the oracle and the CUDA path on machines that have neither /root/reference nor oracle/_ref.

    make -C oracle -f Makefile.ref && python tests/golden/make_golden_ref.py

Cases the reference gets wrong are left out and listed in DESIGN.md (csrmm with a ragged last column block,
column-major csrmm over more than one column block, flash csrmm 'T', in_mem_csrmm 'T'+column-major on a
non-square matrix).
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
import oracle  # noqa: E402
from oracle import ref_run as rr  # noqa: E402

S = "_small"


def ragged_csr(rng, m, n, max_nnz, with_dups=False):
    ia = [0]
    ja, a = [], []
    for r in range(m):
        cnt = 0 if r % 7 == 3 else int(rng.integers(1, max_nnz + 1))
        cols = np.sort(rng.choice(n, size=min(cnt, n), replace=False))
        if with_dups and len(cols) > 1 and r % 5 == 0:
            cols[1] = cols[0]  # duplicate (r, c) entry, kept in storage order
        ja.extend(cols.tolist())
        a.extend(rng.random(len(cols)).astype(np.float32).tolist())
        ia.append(len(ja))
    return np.array(a, np.float32), np.array(ia, np.int64), np.array(ja, np.int64)


def dense(shape, seed):
    return oracle.gen_dense(shape, seed=seed)


def main():
    assert rr.available() and rr.available(S), "build oracle/_ref first: make -C oracle -f Makefile.ref"
    rng = np.random.default_rng(0x0BEF)
    out = {}

    # ---- gemm: flash::gemm, 2 x 2 row/k blocks of 256 with ragged tails, beta=1 chains; all 8 layouts ----
    M, N, K = 400, 40, 420
    out.update(ge_M=M, ge_N=N, ge_K=K, ge_alpha=1.5, ge_beta=0.5)
    ge_a, ge_b, ge_c = dense(M * K, 11), dense(K * N, 12), dense(M * N, 13)
    out.update(ge_seeds=np.array([11, 12, 13]))  # flat buffers, interpreted per layout with tight ld
    for o in "RC":
        for ta in "NT":
            for tb in "NT":
                ar, ac = (M, K) if ta == "N" else (K, M)
                br, bc = (K, N) if tb == "N" else (N, K)
                cr, cc = M, N
                if o == "C":
                    ar, ac, br, bc, cr, cc = ac, ar, bc, br, cc, cr
                got = rr.gemm(o, ta, tb, M, N, K, 1.5, 0.5, ge_a, ge_b, ge_c, ac, bc, cc, flash=True, suffix=S)
                out[f"gemm_flash_{o}{ta}{tb}"] = got
    out["gemm_inmem_RNN"] = rr.gemm("R", "N", "N", M, N, K, 1.5, 0.5, ge_a, ge_b, ge_c, K, N, N)
    # two N blocks too (beta = 0: C must not be read)
    M2, N2, K2 = 130, 420, 140
    g2a, g2b = dense(M2 * K2, 14), dense(K2 * N2, 15)
    out.update(ge2_M=M2, ge2_N=N2, ge2_K=K2, ge2_seeds=np.array([14, 15]))
    out["gemm2_flash_RNN_b0"] = rr.gemm("R", "N", "N", M2, N2, K2, 1.0, 0.0, g2a, g2b,
                                        np.full(M2 * N2, np.nan, np.float32), K2, N2, N2, flash=True, suffix=S)

    # ---- csrmm: 700 x 530 ragged CSR (empty rows), several nnz-budgeted row blocks in the _small build ----
    m, n = 700, 530
    a, ia, ja = ragged_csr(rng, m, n, 24)
    out.update(sp_m=m, sp_n=n, sp_a=a, sp_ia=ia, sp_ja=ja)
    k = 32
    B, C0 = dense((n, k), 21), dense((m, k), 22)
    out.update(sp_k=k, sp_seeds=np.array([21, 22, 23, 24]))  # B, C0, Bt, B2
    out["csrmm_flash_NR_a15_b05"] = rr.csrmm("N", m, n, k, 1.5, 0.5, a, ia, ja, "R", B, C0, flash=True, suffix=S)
    out["csrmm_flash_NR_a1_b0"] = rr.csrmm("N", m, n, k, 1.0, 0.0, a, ia, ja, "R", B, np.zeros_like(C0), flash=True,
                                           suffix=S)
    out["csrmm_inmem_NR_a15_b05"] = rr.csrmm("N", m, n, k, 1.5, 0.5, a, ia, ja, "R", B, C0)
    # column-major B (k x n in memory) and C (k x m in memory)
    Bc, Cc = np.ascontiguousarray(B.T), np.ascontiguousarray(C0.T)
    out["csrmm_flash_NC_a15_b05"] = rr.csrmm("N", m, n, k, 1.5, 0.5, a, ia, ja, "C", Bc, Cc, flash=True, suffix=S)
    out["csrmm_inmem_NC_a15_b05"] = rr.csrmm("N", m, n, k, 1.5, 0.5, a, ia, ja, "C", Bc, Cc)
    # transposed product through the in-memory driver (the flash 'T' path does not terminate, App. A-3)
    Bt = dense((m, k), 23)
    out["csrmm_inmem_TR_a1_b0"] = rr.csrmm("T", m, n, k, 1.0, 0.0, a, ia, ja, "R", Bt, np.zeros((n, k), np.float32))
    # two column blocks of 128 (strided B / C slices, csrmm_task.h:175-199)
    k2 = 256
    B2 = dense((n, k2), 24)
    out.update(sp_k2=k2)
    out["csrmm_flash_NR_k256"] = rr.csrmm("N", m, n, k2, 1.0, 0.0, a, ia, ja, "R", B2, np.zeros((m, k2), np.float32),
                                          flash=True, suffix=S)
    # (drivers/csrmm_pmem.cpp, the in-memory B/C overload, cannot run as shipped: no flash_setup on its main thread,
    #  and it "writes" C through a std::fstream opened without ios::out)

    # ---- csrgemv: flash::csrgemv 'N' and 'T' ----
    x, xt = rng.random(n, dtype=np.float32), rng.random(m, dtype=np.float32)
    out.update(sp_x=x, sp_xt=xt)
    out["csrgemv_flash_N"] = rr.csrgemv("N", m, n, a, ia, ja, x, flash=True, suffix=S)
    out["csrgemv_flash_T"] = rr.csrgemv("T", m, n, a, ia, ja, xt, flash=True, suffix=S)
    out["csrgemv_inmem_N"] = rr.csrgemv("N", m, n, a, ia, ja, x)
    out["csrgemv_inmem_T"] = rr.csrgemv("T", m, n, a, ia, ja, xt)

    # ---- csrcsc: flash::csrcsc (row blocks -> per-block transpose -> column-block merge), duplicates kept ----
    # (n % CSRCSC_CBLK_SIZE must be 0 or >= 10: get_next_blk_size starts at min_size even when fewer rows remain,
    #  include/blas_utils.h:72-82, and the reference then walks off the offsets array)
    a2, ia2, ja2 = ragged_csr(rng, 1300, 906, 10, with_dups=True)
    out.update(tr_m=1300, tr_n=906, tr_a=a2, tr_ia=ia2, tr_ja=ja2)
    t_ia, t_ja, t_a = rr.csrcsc(1300, 906, ia2, ja2, a2, flash=True, suffix=S)
    out.update(csrcsc_flash_ia=t_ia, csrcsc_flash_ja=t_ja, csrcsc_flash_a=t_a)
    i_ia, i_ja, i_a = rr.csrcsc(1300, 906, ia2, ja2, a2)
    assert np.array_equal(i_ia, t_ia) and np.array_equal(i_ja, t_ja) and np.array_equal(i_a.view(np.uint32),
                                                                                       t_a.view(np.uint32))
    # the matrix of the products above, wide (n < m) -- padded-square path of the in-memory driver
    s_ia, s_ja, s_a = rr.csrcsc(m, n, ia, ja, a, flash=True, suffix=S)
    out.update(csrcsc_sp_ia=s_ia, csrcsc_sp_ja=s_ja, csrcsc_sp_a=s_a)

    # ---- kmeans: in_mem_kmeans driver, 3 Lloyd iterations; flash::kmeans distance matrix ----
    P, Kc, d = 2048, 16, 32
    mu = (rng.normal(size=(Kc, d)) * 4).astype(np.float32)
    pts = (mu[rng.integers(0, Kc, P)] + 0.3 * rng.normal(size=(P, d))).astype(np.float32)
    c0 = pts[:Kc].copy()
    out.update(km_points=pts, km_centers0=c0)
    out["km_centers_iter1"] = rr.kmeans_iters(pts, c0, iters=1)
    out["km_centers_iter3"] = rr.kmeans_iters(pts, c0, iters=3)
    D = rr.kmeans_dist(pts[:600], c0, suffix=S)
    out["km_dist_600"] = D
    out["km_assign_600"] = np.argmin(np.abs(D), axis=1).astype(np.int64)  # cblas_isamin, in_mem_kmeans.cpp:84-85

    path = Path(__file__).with_name("golden_ref.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, "with", len(out), "arrays,", path.stat().st_size, "bytes")


if __name__ == "__main__":
    main()
