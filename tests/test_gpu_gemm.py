"""GPU parity: fp32 GEMM (3xTF32 on tcgen05, and the CUDA-core edge path) vs the oracle."""
from pathlib import Path

import numpy as np
import pytest
import torch

import oracle
from gpu_util import dev

pytestmark = pytest.mark.gpu
TOL = 1e-5  # relative Frobenius error, BASELINE.json north_star
G = np.load(Path(__file__).parent / "golden" / "golden_small.npz")

# name -> (gemm_force_path, gemm_split): tc = pure 3xTF32, hyb = TF32 hi*hi + BF16 cross terms
PATHS = {"tc1": (1, 1), "tc2": (2, 1), "ffma": (3, 0), "hyb1": (1, 2), "hyb2": (2, 2)}


@pytest.fixture(scope="module", params=list(PATHS))
def pctx(bof, request):
    path, split = PATHS[request.param]
    c = bof.Context(device=0, gemm_force_path=path, gemm_split=split)
    c.path = request.param
    yield c
    c.close()


def store(X, trans, ord_):
    X = X.T if trans == "T" else X
    return np.ascontiguousarray(X if ord_ == "R" else X.T)


def run_gemm(c, ord_, ta, tb, A, B, C0, alpha, beta):
    M, K = A.shape
    N = B.shape[1]
    Cd = dev(store(C0, "N", ord_))
    c.sgemm(ord_, ta, tb, M, N, K, alpha, dev(store(A, ta, ord_)), 0, dev(store(B, tb, ord_)), 0, beta, Cd, 0)
    got = Cd.cpu().numpy()
    return got if ord_ == "R" else got.T


@pytest.mark.parametrize("shape", [(128, 128, 32), (256, 256, 64), (512, 384, 640), (300, 260, 530), (1, 1, 1),
                                   (129, 257, 33), (1000, 24, 2000), (64, 1024, 100)])
def test_gemm_shapes(pctx, shape):
    M, N, K = shape
    rng = np.random.default_rng(M + N + K)
    A = rng.random((M, K), dtype=np.float32)
    B = rng.random((K, N), dtype=np.float32)
    C0 = rng.random((M, N), dtype=np.float32)
    for alpha, beta in ((1.0, 0.0), (-0.75, 1.5)):
        C_in = C0 if beta else np.full((M, N), np.nan, np.float32)  # beta == 0: C must not be read
        got = run_gemm(pctx, "R", "N", "N", A, B, C_in, alpha, beta)
        ref = oracle.gemm("R", "N", "N", M, N, K, alpha, beta, A, B, C0, acc64=True)
        assert oracle.rel_fro(got, ref) <= TOL, (pctx.path, shape, alpha, beta, oracle.rel_fro(got, ref))


@pytest.mark.parametrize("ord_", "RC")
@pytest.mark.parametrize("ta", "NT")
@pytest.mark.parametrize("tb", "NT")
def test_gemm_eight_layouts(pctx, ord_, ta, tb):
    """misc/gemm_run.sh:31-38 -- the 8 (transA, transB, order) configurations, U[0,1) inputs."""
    D = 768 if pctx.path != "ffma" else 256
    rng = np.random.default_rng(3)
    A = rng.random((D, D + 64), dtype=np.float32)
    B = rng.random((D + 64, D - 32), dtype=np.float32)
    C0 = np.zeros((D, D - 32), np.float32)
    got = run_gemm(pctx, ord_, ta, tb, A, B, C0, 1.0, 0.0)
    ref = oracle.gemm("R", "N", "N", D, D - 32, D + 64, 1.0, 0.0, A, B, C0, acc64=True)
    assert oracle.rel_fro(got, ref) <= TOL, (pctx.path, ord_, ta, tb, oracle.rel_fro(got, ref))


def test_gemm_golden(pctx):
    M, N, K = int(G["ge_M"]), int(G["ge_N"]), int(G["ge_K"])
    got = run_gemm(pctx, "R", "N", "N", G["ge_A"], G["ge_B"], G["ge_C0"], 1.25, 0.75)
    assert oracle.rel_fro(got, G["gemm_a125_b075"]) <= TOL


def test_gemm_padded_leading_dims(pctx):
    rng = np.random.default_rng(4)
    M, N, K = 200, 136, 72
    A = np.zeros((M, K + 5), np.float32); A[:, :K] = rng.random((M, K), dtype=np.float32)
    B = np.zeros((K, N + 3), np.float32); B[:, :N] = rng.random((K, N), dtype=np.float32)
    Cp = np.full((M, N + 7), 3.0, np.float32)
    Cd = dev(Cp)
    pctx.sgemm("R", "N", "N", M, N, K, 1.0, dev(A), K + 5, dev(B), N + 3, 0.0, Cd, N + 7)
    got = Cd.cpu().numpy()
    ref = A[:, :K].astype(np.float64) @ B[:, :N].astype(np.float64)
    assert oracle.rel_fro(got[:, :N], ref) <= TOL
    assert np.all(got[:, N:] == 3.0)


def test_gemm_3072_like_gemm_run_sh(ctx):
    """DIM=3072 of misc/gemm_run.sh:4 on the default (2-CTA tensor-core) path."""
    D = 3072
    A = oracle.gen_dense((D, D), seed=21)
    B = oracle.gen_dense((D, D), seed=22)
    Cd = torch.empty((D, D), device="cuda")
    ctx.sgemm("R", "N", "N", D, D, D, 1.0, dev(A), 0, dev(B), 0, 0.0, Cd, 0)
    ref = (A.astype(np.float64) @ B.astype(np.float64))
    err = oracle.rel_fro(Cd.cpu().numpy(), ref)
    mre = np.max(np.abs(Cd.cpu().numpy() - ref) / ref)  # the script's max-relative-error
    assert err <= TOL and mre <= 1e-4, (err, mre)


def test_gemm_integer_data_exact(ctx):
    """dense_create 's' pattern (i % 10): all products and sums are exact in fp32 => bit-exact."""
    M, N, K = 512, 512, 1024
    A = oracle.gen_dense((M, K), mode=0)
    B = oracle.gen_dense((K, N), mode=0)
    Cd = torch.empty((M, N), device="cuda")
    ctx.sgemm("R", "N", "N", M, N, K, 1.0, dev(A), 0, dev(B), 0, 0.0, Cd, 0)
    ref = (A.astype(np.float64) @ B.astype(np.float64)).astype(np.float32)
    assert np.array_equal(Cd.cpu().numpy(), ref)


@pytest.mark.parametrize("split", [1, 2])
def test_gemm_split_modes_worst_case_data(bof, split):
    """Constant matrices make every product carry the same rounding error (no cancellation across k): the
    adversarial case for the bf16 cross terms of the hybrid split (bound 2^-19 per product)."""
    M, N, K = 512, 512, 4096
    for va, vb in ((1.2345678, 0.87654321), (1.9999999, 1.0000001), (3.1415927, 2.7182817)):
        A = np.full((M, K), va, np.float32); B = np.full((K, N), vb, np.float32)
        ref = A.astype(np.float64) @ B.astype(np.float64)
        with bof.Context(device=0, gemm_split=split) as c2:
            Cd = torch.empty((M, N), device="cuda")
            c2.sgemm("R", "N", "N", M, N, K, 1.0, dev(A), 0, dev(B), 0, 0.0, Cd, 0)
            err = oracle.rel_fro(Cd.cpu().numpy(), ref)
        print(f"split={split} constant data ({va}, {vb}): rel-Frobenius {err:.3e}")
        assert err <= TOL, (split, va, vb, err)


@pytest.mark.parametrize("split", [1, 2])
@pytest.mark.parametrize("k_chunk", [-1, 256, 2048])
def test_gemm_long_k_accuracy(bof, k_chunk, split):
    """k = 32768 (cfg-2's reduction length) with all-positive data: the worst case for accumulator
    rounding.  Reference = fp64 accumulation."""
    M, N, K = 256, 256, 32768
    A = oracle.gen_dense((M, K), seed=31)
    B = oracle.gen_dense((K, N), seed=32)
    ref = A.astype(np.float64) @ B.astype(np.float64)
    with bof.Context(device=0, gemm_k_chunk=k_chunk, gemm_split=split) as c2:
        Cd = torch.empty((M, N), device="cuda")
        c2.sgemm("R", "N", "N", M, N, K, 1.0, dev(A), 0, dev(B), 0, 0.0, Cd, 0)
        err = oracle.rel_fro(Cd.cpu().numpy(), ref)
    print(f"long-k rel-Frobenius error with k_chunk={k_chunk} split={split}: {err:.3e}")
    if k_chunk == 256:  # the shipped default must hold the tolerance
        assert err <= TOL, err


def test_gemm_wave_lockstep_is_result_neutral(bof):
    """The wave lock-step only paces the TMA producers: results are bit-identical with it on, off, or dense."""
    M, N, K = 2560, 2816, 8192  # 10 x 11 = 110 tiles > 74 clusters, 256 k-blocks
    A = oracle.gen_dense((M, K), seed=41)
    B = oracle.gen_dense((K, N), seed=42)
    outs = []
    for sync in (-1, 0, 8):
        with bof.Context(device=0, gemm_wave_sync=sync) as c2:
            Cd = torch.empty((M, N), device="cuda")
            c2.sgemm("R", "N", "N", M, N, K, 1.0, dev(A), 0, dev(B), 0, 0.0, Cd, 0)
            outs.append(Cd.cpu())
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])
    ref = A[:64].astype(np.float64) @ B.astype(np.float64)
    assert oracle.rel_fro(outs[1][:64].numpy(), ref) <= TOL


def test_gemm_linearity_full_size_property(ctx):
    """Size-independent check at 8192^2 x 4096: (A1 + A2) B == A1 B + A2 B with integer-valued data (exact)."""
    M, N, K = 8192, 8192, 4096
    A1 = torch.randint(0, 3, (M, K), device="cuda").float()
    A2 = torch.randint(0, 3, (M, K), device="cuda").float()
    B = torch.randint(0, 3, (K, N), device="cuda").float()
    ws = ctx.sgemm_workspace(M, N, K)
    C1 = torch.empty((M, N), device="cuda"); C2 = torch.empty_like(C1); C3 = torch.empty_like(C1)
    ctx.sgemm("R", "N", "N", M, N, K, 1.0, A1, 0, B, 0, 0.0, C1, 0, ws=ws)
    ctx.sgemm("R", "N", "N", M, N, K, 1.0, A2, 0, B, 0, 0.0, C2, 0, ws=ws)
    ctx.sgemm("R", "N", "N", M, N, K, 1.0, A1 + A2, 0, B, 0, 0.0, C3, 0, ws=ws)
    assert torch.equal(C3, C1 + C2)
    # checksum of checksums: sum(C) == colsum(A) . rowsum(B)
    assert torch.isclose(C1.double().sum(), (A1.double().sum(0) * B.double().sum(1)).sum(), rtol=1e-12)


@pytest.mark.parametrize("ord_,ta,tb", [("R", "N", "N"), ("R", "T", "N"), ("C", "N", "T"), ("C", "T", "T")])
def test_host_gemm(bof, ord_, ta, tb):
    """flash::gemm through the host entry point; a small row block forces a multi-block pipeline."""
    rng = np.random.default_rng(6)
    M, N, K = 1300, 700, 900
    A = rng.random((M, K), dtype=np.float32)
    B = rng.random((K, N), dtype=np.float32)
    C0 = rng.random((M, N), dtype=np.float32)
    ref = oracle.gemm("R", "N", "N", M, N, K, 1.5, 0.5, A, B, C0, acc64=True)
    with bof.Context(device=0, gemm_row_block=256) as c2:
        c_h = store(C0, "N", ord_)
        c2.host_gemm(ord_, ta, tb, M, N, K, 1.5, 0.5, store(A, ta, ord_), store(B, tb, ord_), c_h)
        got = c_h if ord_ == "R" else c_h.T
        assert oracle.rel_fro(got, ref) <= TOL
        st = c2.stats()
        assert st.h2d_bytes >= (M * K + K * N + M * N) * 4 and st.d2h_bytes == M * N * 4


def test_host_devb_variants(ctx):
    """Replicated operand already in HBM (the all-gather path of multi-GPU runs): same results as the host path."""
    rng = np.random.default_rng(8)
    M, N, K = 900, 640, 500
    A = rng.random((M, K), dtype=np.float32); B = rng.random((K, N), dtype=np.float32)
    C0 = rng.random((M, N), dtype=np.float32)
    c_h = C0.copy()
    ctx.host_gemm_devb("N", "N", M, N, K, 1.5, 0.5, A, dev(B), c_h)
    assert oracle.rel_fro(c_h, oracle.gemm("R", "N", "N", M, N, K, 1.5, 0.5, A, B, C0, acc64=True)) <= TOL
    c_h = C0.copy()
    ctx.host_gemm_devb("N", "T", M, N, K, 1.0, 0.0, A, dev(np.ascontiguousarray(B.T)), c_h)
    assert oracle.rel_fro(c_h, oracle.gemm("R", "N", "N", M, N, K, 1.0, 0.0, A, B, C0, acc64=True)) <= TOL
    from gpu_util import ragged_csr
    m, n, k = 2000, 1500, 64
    a, ia, ja = ragged_csr(rng, m, n, 40)
    Bd = rng.random((n, k), dtype=np.float32)
    c2 = np.full((m, k), np.nan, np.float32)
    ctx.host_csrmm_devb(m, n, k, 1.0, 0.0, a, ia, ja, dev(Bd), c2)
    assert oracle.rel_fro(c2, oracle.csrmm("N", m, n, k, 1.0, 0.0, a, ia, ja, "R", Bd, np.zeros((m, k), np.float32), acc64=True)) <= TOL


def test_host_gemm_bad_args(ctx):
    from bof_b200 import ptr
    z = np.zeros(4, np.float32)
    for bad in ((b"X", b"N", b"N"), (b"R", b"X", b"N"), (b"R", b"N", b"X")):
        assert ctx.lib.bof_host_gemm(ctx.h, *bad, 2, 2, 2, 1.0, 0.0, ptr(z), ptr(z), ptr(z), 0, 0, 0) == -1
    assert ctx.lib.bof_host_gemm(ctx.h, b"R", b"N", b"N", 2, 2, 2, 1.0, 0.0, ptr(z), ptr(z), ptr(z), 1, 0, 0) == -1
    assert "leading dimension" in ctx.last_error()
