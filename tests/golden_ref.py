"""Cases of tests/golden/golden_ref.npz -- outputs written by the reference's own binaries (oracle/_ref, see
tests/golden/make_golden_ref.py).  Shared by the CPU oracle tests and the GPU parity tests: one callable per
kernel computes every case of that kernel and the tests compare with the stored reference output."""
from pathlib import Path

import numpy as np

import oracle

G = np.load(Path(__file__).parent / "golden" / "golden_ref.npz")
TOL = 1e-5  # relative Frobenius error (BASELINE.json north_star); integer / byte outputs are compared bit for bit


def dense(shape, seed):
    return oracle.gen_dense(shape, seed=int(seed))


def gemm_layout(o, ta, tb, M, N, K):
    """(lda, ldb, ldc) of tight operands for a layout (rows x cols swap under column-major)."""
    ac = K if ta == "N" else M
    bc = N if tb == "N" else K
    cc = N
    if o == "C":
        ac = M if ta == "N" else K
        bc = K if tb == "N" else N
        cc = M
    return ac, bc, cc


def gemm_cases():
    M, N, K = int(G["ge_M"]), int(G["ge_N"]), int(G["ge_K"])
    sa, sb, sc = G["ge_seeds"]
    a, b, c = dense(M * K, sa), dense(K * N, sb), dense(M * N, sc)
    for o in "RC":
        for ta in "NT":
            for tb in "NT":
                yield (f"flash_{o}{ta}{tb}", o, ta, tb, M, N, K, 1.5, 0.5, a, b, c, G[f"gemm_flash_{o}{ta}{tb}"])
    yield ("inmem_RNN", "R", "N", "N", M, N, K, 1.5, 0.5, a, b, c, G["gemm_inmem_RNN"])
    M2, N2, K2 = int(G["ge2_M"]), int(G["ge2_N"]), int(G["ge2_K"])
    s1, s2 = G["ge2_seeds"]
    yield ("flash2_RNN_b0", "R", "N", "N", M2, N2, K2, 1.0, 0.0, dense(M2 * K2, s1), dense(K2 * N2, s2),
           np.full(M2 * N2, np.nan, np.float32), G["gemm2_flash_RNN_b0"])


def sparse_inputs():
    m, n, k, k2 = int(G["sp_m"]), int(G["sp_n"]), int(G["sp_k"]), int(G["sp_k2"])
    sB, sC, sBt, sB2 = G["sp_seeds"]
    return dict(m=m, n=n, k=k, k2=k2, a=G["sp_a"], ia=G["sp_ia"], ja=G["sp_ja"], B=dense((n, k), sB),
                C0=dense((m, k), sC), Bt=dense((m, k), sBt), B2=dense((n, k2), sB2), x=G["sp_x"], xt=G["sp_xt"])


def csrmm_cases():
    """(name, trans, k, alpha, beta, ord_b, B, C, expected) -- B/C in the memory layout the call takes."""
    s = sparse_inputs()
    m, n, k = s["m"], s["n"], s["k"]
    B, C0 = s["B"], s["C0"]
    Bc, Cc = np.ascontiguousarray(B.T), np.ascontiguousarray(C0.T)
    yield ("flash_NR_a15_b05", "N", k, 1.5, 0.5, "R", B, C0, G["csrmm_flash_NR_a15_b05"])
    yield ("flash_NR_a1_b0", "N", k, 1.0, 0.0, "R", B, np.full_like(C0, np.nan), G["csrmm_flash_NR_a1_b0"])
    yield ("inmem_NR_a15_b05", "N", k, 1.5, 0.5, "R", B, C0, G["csrmm_inmem_NR_a15_b05"])
    yield ("flash_NC_a15_b05", "N", k, 1.5, 0.5, "C", Bc, Cc, G["csrmm_flash_NC_a15_b05"])
    yield ("inmem_NC_a15_b05", "N", k, 1.5, 0.5, "C", Bc, Cc, G["csrmm_inmem_NC_a15_b05"])
    yield ("inmem_TR_a1_b0", "T", k, 1.0, 0.0, "R", s["Bt"], np.full((n, k), np.nan, np.float32),
           G["csrmm_inmem_TR_a1_b0"])
    yield ("flash_NR_k256", "N", s["k2"], 1.0, 0.0, "R", s["B2"], np.full((m, s["k2"]), np.nan, np.float32),
           G["csrmm_flash_NR_k256"])


def csrgemv_cases():
    s = sparse_inputs()
    for kind in ("flash", "inmem"):
        yield (f"{kind}_N", "N", s["x"], G[f"csrgemv_{kind}_N"])
        yield (f"{kind}_T", "T", s["xt"], G[f"csrgemv_{kind}_T"])


def csrcsc_cases():
    """(name, m, n, ia, ja, a, (ia_t, ja_t, a_t))"""
    yield ("flash_dups", int(G["tr_m"]), int(G["tr_n"]), G["tr_ia"], G["tr_ja"], G["tr_a"],
           (G["csrcsc_flash_ia"], G["csrcsc_flash_ja"], G["csrcsc_flash_a"]))
    yield ("flash_sp", int(G["sp_m"]), int(G["sp_n"]), G["sp_ia"], G["sp_ja"], G["sp_a"],
           (G["csrcsc_sp_ia"], G["csrcsc_sp_ja"], G["csrcsc_sp_a"]))


def same_csc(got, want) -> bool:
    return (np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1])
            and np.array_equal(np.asarray(got[2], np.float32).view(np.uint32), want[2].view(np.uint32)))
