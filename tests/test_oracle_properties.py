"""Property tests (hypothesis) of the CPU oracle: size-independent identities the GPU parity tests also lean on at
full size -- double transpose, stability, linearity, tile/chain invariance, assignment optimality."""
import numpy as np
from hypothesis import given, settings, strategies as st

import oracle

FAST = settings(max_examples=25, deadline=None)


@st.composite
def csr_matrices(draw, max_m=40, max_n=40, max_row=12, dups=False):
    m = draw(st.integers(0, max_m)); n = draw(st.integers(1, max_n))
    seed = draw(st.integers(0, 2**31 - 1))
    rng = np.random.default_rng(seed)
    counts = rng.integers(0, min(max_row, n) + 1, size=m)
    ia = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    cols = [np.sort(rng.choice(n, c, replace=dups)) for c in counts]
    ja = (np.concatenate(cols) if m and ia[-1] else np.zeros(0)).astype(np.int64)
    a = (rng.integers(1, 10, size=int(ia[-1]))).astype(np.float32)       # (i % 9) + 1 style values: sums are exact
    return m, n, a, ia, ja


@FAST
@given(csr_matrices(dups=True))
def test_csrcsc_is_a_stable_involution(mat):
    m, n, a, ia, ja = mat
    ia_t, ja_t, a_t = oracle.csrcsc(m, n, ia, ja, a)
    assert ia_t[0] == 0 and ia_t[-1] == len(a) and np.all(np.diff(ia_t) >= 0)
    assert np.array_equal(np.diff(ia_t), np.bincount(ja, minlength=n))          # column histogram
    for c in range(n):                                                          # rows ascending inside a column: stable
        assert np.all(np.diff(ja_t[ia_t[c]:ia_t[c + 1]]) >= 0)
    ia2, ja2, a2 = oracle.csrcsc(n, m, ia_t, ja_t, a_t)
    assert np.array_equal(ia2, ia) and np.array_equal(ja2, ja) and np.array_equal(a2.view(np.int32), a.view(np.int32))


@FAST
@given(csr_matrices(), st.integers(1, 9), st.sampled_from(["R", "C"]), st.sampled_from(["N", "T"]))
def test_csrmm_is_linear_and_matches_dense(mat, k, ord_b, trans):
    m, n, a, ia, ja = mat
    rows_b, rows_c = (n, m) if trans == "N" else (m, n)
    rng = np.random.default_rng(k)
    B1 = rng.integers(0, 10, size=(rows_b, k)).astype(np.float32); B2 = rng.integers(0, 10, size=(rows_b, k)).astype(np.float32)
    lay = (lambda X: X) if ord_b == "R" else (lambda X: np.ascontiguousarray(X.T))
    unlay = (lambda X: X) if ord_b == "R" else (lambda X: X.reshape(k, rows_c).T)
    Z = np.zeros((rows_c, k), np.float32)
    f = lambda B: unlay(oracle.csrmm(trans, m, n, k, 1.0, 0.0, a, ia, ja, ord_b, lay(B), lay(Z)).reshape(lay(Z).shape))
    dense = np.zeros((m, n), np.float64)
    for r in range(m):
        np.add.at(dense[r], ja[ia[r]:ia[r + 1]], a[ia[r]:ia[r + 1]])
    op = dense if trans == "N" else dense.T
    assert np.array_equal(f(B1), (op @ B1.astype(np.float64)).astype(np.float32))   # integer data: exact in fp32
    assert np.array_equal(f(B1 + 2 * B2), f(B1) + 2 * f(B2))


@FAST
@given(st.integers(1, 40), st.integers(1, 40), st.integers(1, 60), st.integers(1, 16), st.integers(0, 2**31 - 1))
def test_gemm_tile_chain_equals_monolithic_on_exact_data(m, n, k, blk, seed):
    """The reference's 3-D tiling with beta=1 chains (src/blas/gemm.cpp:83-129) cannot change an exact result."""
    rng = np.random.default_rng(seed)
    A = rng.integers(0, 10, size=(m, k)).astype(np.float32); B = rng.integers(0, 10, size=(k, n)).astype(np.float32)
    C0 = rng.integers(0, 10, size=(m, n)).astype(np.float32)
    mono = oracle.gemm("R", "N", "N", m, n, k, 2.0, 1.0, A, B, C0)
    tiled = oracle.gemm("R", "N", "N", m, n, k, 2.0, 1.0, A, B, C0, tiled_blk=blk)
    assert np.array_equal(mono, tiled) and np.array_equal(mono, 2 * (A @ B) + C0)


@FAST
@given(st.integers(1, 60), st.integers(1, 8), st.integers(1, 6), st.integers(0, 2**31 - 1))
def test_kmeans_assignment_minimises_the_reference_distance(P, K, d, seed):
    rng = np.random.default_rng(seed)
    pts = rng.normal(size=(P, d)).astype(np.float32); cent = rng.normal(size=(K, d)).astype(np.float32)
    assign, margin = oracle.kmeans_assign(pts, cent)
    c2 = (cent.astype(np.float64) ** 2).sum(1); p2 = (pts.astype(np.float64) ** 2).sum(1)
    D = np.abs(-2.0 * pts.astype(np.float64) @ cent.astype(np.float64).T + c2[None, :] + p2[:, None])
    best = D.min(axis=1)
    assert np.all(D[np.arange(P), assign] <= best + 1e-4 * (1 + best))          # optimal up to fp32 rounding
    assert np.all(margin >= 0)
    new_c, a2, _ = oracle.lloyd_iter(pts, cent)
    for c in range(K):                                                          # centroid = mean of its points, empty => 0
        sel = a2 == c
        want = pts[sel].astype(np.float64).mean(0) if sel.any() else np.zeros(d)
        assert np.allclose(new_c[c], want, rtol=1e-4, atol=1e-5)
