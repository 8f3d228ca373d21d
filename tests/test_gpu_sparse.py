"""GPU parity: SpMM / SpMV device-tile kernels and the host csrmm / csrgemv pipelines vs the oracle."""
from pathlib import Path

import numpy as np
import pytest
import torch

import oracle
from gpu_util import csr_to_device, dev, ragged_csr

pytestmark = pytest.mark.gpu
TOL = 1e-5  # relative Frobenius error, BASELINE.json north_star
G = np.load(Path(__file__).parent / "golden" / "golden_small.npz")


@pytest.mark.parametrize("k", [4, 8, 32, 64, 100, 128, 256, 260, 1000, 7, 131])
def test_spmm_rowmajor_k_sweep(ctx, k):
    rng = np.random.default_rng(k)
    m, n = 777, 513
    a, ia, ja = ragged_csr(rng, m, n, 40)
    B = rng.random((n, k), dtype=np.float32)
    C0 = rng.random((m, k), dtype=np.float32)
    vals, idx, offs = csr_to_device(a, ia, ja)
    for alpha, beta in ((1.0, 0.0), (1.5, 0.5)):
        Cd = dev(C0 if beta else np.full((m, k), np.nan, np.float32))  # beta == 0: C must not be read
        ctx.spmm("R", m, n, k, alpha, vals, idx, offs, dev(B), k, beta, Cd, k)
        ref = oracle.csrmm("N", m, n, k, alpha, beta, a, ia, ja, "R", B, C0, acc64=True)
        assert oracle.rel_fro(Cd.cpu().numpy(), ref) <= TOL


def test_spmm_golden_fixture(ctx):
    m, n, k = int(G["sp_m"]), int(G["sp_n"]), int(G["sp_k"])
    vals, idx, offs = csr_to_device(G["sp_a"], G["sp_ia"], G["sp_ja"])
    Cd = dev(G["sp_C0"])
    ctx.spmm("R", m, n, k, 1.5, vals, idx, offs, dev(G["sp_B"]), k, 0.5, Cd, k)
    assert oracle.rel_fro(Cd.cpu().numpy(), G["spmm_R_a15_b05"]) <= TOL
    # column-major B and C (SimpleCsrmmCmTask)
    Bc = np.asfortranarray(G["sp_B"]).T.copy().reshape(-1)
    Cc = dev(np.asfortranarray(G["sp_C0"]).T.copy().reshape(-1))
    ctx.spmm("C", m, n, k, 1.5, vals, idx, offs, dev(Bc), n, 0.5, Cc, m)
    assert oracle.rel_fro(Cc.cpu().numpy(), G["spmm_C_a15_b05"]) <= TOL


def test_spmm_padded_ld_and_unrebased_offsets(ctx):
    rng = np.random.default_rng(5)
    m, n, k = 300, 200, 64
    a, ia, ja = ragged_csr(rng, m, n, 20)
    B = np.zeros((n, k + 8), np.float32); B[:, :k] = rng.random((n, k), dtype=np.float32)
    Cp = np.full((m, k + 4), 7.0, np.float32)
    r0, r1 = 50, 250
    z0, z1 = ia[r0], ia[r1]
    vals, idx, offs = dev(a[z0:z1]), dev(ja[z0:z1].astype(np.int32)), dev(ia[r0:r1 + 1])  # offs[0] != 0
    Cd = dev(Cp[r0:r1])
    ctx.spmm("R", r1 - r0, n, k, 1.0, vals, idx, offs, dev(B), k + 8, 0.0, Cd, k + 4)
    got = Cd.cpu().numpy()
    ref = oracle.csrmm("N", m, n, k, 1.0, 0.0, a, ia, ja, "R", B[:, :k].copy(), np.zeros((m, k), np.float32), acc64=True)
    assert oracle.rel_fro(got[:, :k], ref[r0:r1]) <= TOL
    assert np.all(got[:, k:] == 7.0)  # padding columns untouched


def test_spmm_integer_compat_data_bit_exact(ctx):
    """Reference generator patterns (misc/sparse_create.cpp:52-55, misc/dense_create.cpp:28-32): exact in fp32."""
    m, n, k = 2048, 1024, 128
    a, ia, ja = oracle.gen_csr(m, n, 64, seed=1, val_mode=0)
    B = oracle.gen_dense((n, k), mode=0)
    vals, idx, offs = csr_to_device(a, ia, ja)
    Cd = torch.empty((m, k), dtype=torch.float32, device="cuda")
    ctx.spmm("R", m, n, k, 1.0, vals, idx, offs, dev(B), k, 0.0, Cd, k)
    ref = oracle.csrmm("N", m, n, k, 1.0, 0.0, a, ia, ja, "R", B, np.zeros((m, k), np.float32))
    assert np.array_equal(Cd.cpu().numpy(), ref)


def test_spmm_linearity_at_scale(ctx):
    """Size-independent property on a cfg-1 shaped slice: A(B1 + 2 B2) == A B1 + 2 A B2 (integer data => exact)."""
    m, n, k = 65536, 65536, 128
    a, ia, ja = oracle.gen_csr(m, n, 64, seed=3, val_mode=0)
    vals, idx, offs = csr_to_device(a, ia, ja)
    B1 = torch.randint(0, 4, (n, k), device="cuda").float()
    B2 = torch.randint(0, 4, (n, k), device="cuda").float()
    C1 = torch.empty((m, k), device="cuda"); C2 = torch.empty_like(C1); C3 = torch.empty_like(C1)
    ctx.spmm("R", m, n, k, 1.0, vals, idx, offs, B1, k, 0.0, C1, k)
    ctx.spmm("R", m, n, k, 1.0, vals, idx, offs, B2, k, 0.0, C2, k)
    ctx.spmm("R", m, n, k, 1.0, vals, idx, offs, B1 + 2 * B2, k, 0.0, C3, k)
    assert torch.equal(C3, C1 + 2 * C2)
    # and a checksum: sum of C equals sum_j colsum(A)_j * rowsum(B)_j
    colsum = torch.zeros(n, device="cuda", dtype=torch.float64).index_add_(0, idx.long(), vals.double())
    assert torch.isclose(C1.double().sum(), (colsum * B1.double().sum(dim=1)).sum(), rtol=1e-12)


@pytest.mark.parametrize("trans", ["N", "T"])
def test_spmv(ctx, trans):
    rng = np.random.default_rng(8)
    m, n = 5000, 3000
    a, ia, ja = ragged_csr(rng, m, n, 120)
    x = rng.random(n if trans == "N" else m, dtype=np.float32)
    vals, idx, offs = csr_to_device(a, ia, ja)
    y = torch.full((m if trans == "N" else n,), float("nan"), device="cuda")
    ctx.spmv(trans, m, n, vals, idx, offs, dev(x), y)
    ref = oracle.csrgemv(trans, m, n, a, ia, ja, x, acc64=True)
    assert oracle.rel_fro(y.cpu().numpy(), ref) <= TOL


def test_spmv_golden(ctx):
    m, n = int(G["sp_m"]), int(G["sp_n"])
    vals, idx, offs = csr_to_device(G["sp_a"], G["sp_ia"], G["sp_ja"])
    y = torch.empty(m, device="cuda")
    ctx.spmv("N", m, n, vals, idx, offs, dev(G["sp_x"]), y)
    assert oracle.rel_fro(y.cpu().numpy(), G["spmv_N"]) <= TOL
    yt = torch.empty(n, device="cuda")
    ctx.spmv("T", m, n, vals, idx, offs, dev(G["sp_xt"]), yt)
    assert oracle.rel_fro(yt.cpu().numpy(), G["spmv_T"]) <= TOL


def test_idx_narrow_widen_roundtrip(ctx):
    x = torch.randint(0, 2**31 - 1, (100003,), device="cuda", dtype=torch.int64)
    n32 = torch.empty(x.numel(), device="cuda", dtype=torch.int32)
    back = torch.empty_like(x)
    ctx.idx_narrow(x, n32, x.numel())
    ctx.idx_widen(n32, back, x.numel())
    assert torch.equal(back, x)


@pytest.mark.parametrize("trans,ord_b", [("N", "R"), ("N", "C"), ("T", "R"), ("T", "C")])
def test_host_csrmm(bof, trans, ord_b):
    """flash::csrmm through the host entry point; small nnz budget forces several streamed row blocks."""
    rng = np.random.default_rng(12)
    m, n, k = 3000, 2000, 96
    a, ia, ja = ragged_csr(rng, m, n, 50)
    brows, crows = (n, m) if trans == "N" else (m, n)
    B = rng.random((brows, k), dtype=np.float32)
    C0 = rng.random((crows, k), dtype=np.float32)
    lay = (lambda X: X.copy()) if ord_b == "R" else (lambda X: np.ascontiguousarray(X.T))
    with bof.Context(device=0, csrmm_max_nnz=20000) as c2:
        for alpha, beta in ((1.0, 0.0), (0.5, 2.0)):
            b_h, c_h = lay(B), lay(C0)
            c2.host_csrmm(trans, m, n, k, alpha, beta, a, ia, ja, ord_b, b_h, c_h)
            ref = oracle.csrmm(trans, m, n, k, alpha, beta, a, ia, ja, "R", B, C0, acc64=True)
            got = c_h if ord_b == "R" else c_h.T
            assert oracle.rel_fro(got, ref) <= TOL, (trans, ord_b, alpha, beta)
        if trans == "N":
            assert c2.stats().h2d_bytes > 0 and c2.stats().kernel_launches >= 2


def test_host_csrmm_bad_args_return_minus_one(ctx):
    z = np.zeros(4, np.float32)
    ia = np.zeros(2, np.int64)
    lib = ctx.lib
    from bof_b200 import ptr
    assert lib.bof_host_csrmm(ctx.h, b"X", 1, 1, 1, 1.0, 0.0, ptr(z), ptr(ia), ptr(ia), b"R", ptr(z), ptr(z)) == -1
    assert "trans_a" in ctx.last_error()
    assert lib.bof_host_csrmm(ctx.h, b"N", 1, 1, 1, 1.0, 0.0, ptr(z), ptr(ia), ptr(ia), b"Q", ptr(z), ptr(z)) == -1
    assert "ord_b" in ctx.last_error()


@pytest.mark.parametrize("trans", ["N", "T"])
def test_host_csrgemv(bof, trans):
    rng = np.random.default_rng(13)
    m, n = 4000, 2500
    a, ia, ja = ragged_csr(rng, m, n, 60)
    x = rng.random(n if trans == "N" else m, dtype=np.float32)
    y = np.full(m if trans == "N" else n, np.nan, np.float32)
    with bof.Context(device=0, csrmm_max_nnz=15000) as c2:
        c2.host_csrgemv(trans, m, n, a, ia, ja, x, y)
    assert oracle.rel_fro(y, oracle.csrgemv(trans, m, n, a, ia, ja, x, acc64=True)) <= TOL


def test_host_csrmm_empty(ctx):
    ia = np.zeros(1, np.int64)
    e = np.zeros(0, np.float32)
    ctx.host_csrmm("N", 0, 5, 3, 1.0, 0.0, e, ia, np.zeros(0, np.int64), "R", np.zeros((5, 3), np.float32), e)


@pytest.mark.parametrize("variant", ["6", "3", "8"])
def test_spmm_kernel_variants_child_process(variant):
    """BOF_SPMM_VARIANT=6 selects the TMA-staged A-stream kernel (cp.async.bulk + mbarrier); 3 is a register variant;
    8 is the 64-column-chunk kernel the default policy picks when half of B fits L2 but all of it does not."""
    import os, subprocess, sys
    script = Path(__file__).parent / "spmm_variant_check.py"
    r = subprocess.run([sys.executable, str(script)], env=dict(os.environ, BOF_SPMM_VARIANT=variant),
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "SPMM_VARIANT_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


def test_failed_call_leaves_the_context_usable(bof, ctx):
    """An entry point that fails part-way (device allocation of a 400 GB operand) returns an error, leaves nothing
    running, and the next call on the same context is correct.  Column-major B: a row-major B of that size is a
    legitimate input since round 2 (processed in column panels) and would be read from the host."""
    rng = np.random.default_rng(21)
    m, n, k = 600, 500, 32
    a, ia, ja = ragged_csr(rng, m, n, 20)
    B = rng.random((n, k), dtype=np.float32); C = np.zeros((m, k), np.float32)
    with pytest.raises(bof.BofError):
        ctx.host_csrmm("N", m, 1 << 30, 100, 1.0, 0.0, a, ia, ja, "C", B, C)   # B would be 2^30 x 100 floats
    assert "failed" in ctx.last_error().lower() or "memory" in ctx.last_error().lower()
    ctx.host_csrmm("N", m, n, k, 1.0, 0.0, a, ia, ja, "R", B, C)
    assert oracle.rel_fro(C, oracle.csrmm("N", m, n, k, 1.0, 0.0, a, ia, ja, "R", B, np.zeros((m, k), np.float32), acc64=True)) <= TOL


@pytest.mark.parametrize("k", [128, 200, 256])
def test_spmm_l2_half_policy_shape(ctx, k):
    """B of 72-143 MB (more than half of L2, a 64-column slice fits): the launcher gathers from one slice of B at a time
    (16-lane row groups, ceil(k/64) column chunks)."""
    rng = np.random.default_rng(23)
    m, n = 4096, 140_000
    a, ia, ja = ragged_csr(rng, m, n, 30)
    B = rng.random((n, k), dtype=np.float32)
    C0 = rng.random((m, k), dtype=np.float32)
    vals, idx, offs = csr_to_device(a, ia, ja)
    Cd = dev(C0)
    ctx.spmm("R", m, n, k, 1.5, vals, idx, offs, dev(B), k, 0.5, Cd, k)
    ref = oracle.csrmm("N", m, n, k, 1.5, 0.5, a, ia, ja, "R", B, C0, acc64=True)
    assert oracle.rel_fro(Cd.cpu().numpy(), ref) <= TOL


def zipf_csr(rng, m, n, nnz_target, max_row=None):
    """power-law row lengths: row of rank r gets ~ nnz_target / (r * H_m) nonzeros (capped at n), in random row order"""
    ranks = rng.permutation(m) + 1
    lens = np.minimum((nnz_target / (ranks * np.log(m))).astype(np.int64) + 1, max_row or n)
    ia = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    ja = np.empty(int(ia[-1]), np.int64)
    for r in range(m):
        L = int(lens[r])
        ja[ia[r]:ia[r + 1]] = np.sort(rng.choice(n, size=L, replace=False)) if L < n // 4 else np.sort(rng.permutation(n)[:L])
    a = rng.random(int(ia[-1]), dtype=np.float32)
    return a, ia, ja


@pytest.mark.parametrize("k", [32, 128, 256, 640])
def test_spmm_power_law_rows(ctx, k):
    """row splitting: rows above 1024 nonzeros go to the 8 warps of a block, rows above max(16384, 64 x mean) to every
    warp of the grid (deterministic combine); checked against the oracle on a Zipf row-length distribution"""
    rng = np.random.default_rng(100 + k)
    m, n = 6000, 40000
    a, ia, ja = zipf_csr(rng, m, n, 1_200_000)
    lens = np.diff(ia)
    assert lens.max() > 20000 and (lens > 1024).sum() > 10 and np.median(lens) < 100
    B = rng.random((n, k), dtype=np.float32)
    C0 = rng.random((m, k), dtype=np.float32)
    vals, idx, offs = csr_to_device(a, ia, ja)
    for alpha, beta in ((1.0, 0.0), (1.5, 0.5)):
        Cd = dev(C0 if beta else np.full((m, k), np.nan, np.float32))
        ctx.spmm("R", m, n, k, alpha, vals, idx, offs, dev(B), k, beta, Cd, k)
        got = Cd.cpu().numpy()
        ref = oracle.csrmm("N", m, n, k, alpha, beta, a, ia, ja, "R", B, C0, acc64=True)
        assert oracle.rel_fro(got, ref) <= TOL
        # per-row check of the longest rows (a global Frobenius norm would hide one bad row)
        for r in np.argsort(lens)[-4:]:
            assert oracle.rel_fro(got[r], ref[r]) <= TOL
        Cd2 = dev(C0 if beta else np.full((m, k), np.nan, np.float32))
        ctx.spmm("R", m, n, k, alpha, vals, idx, offs, dev(B), k, beta, Cd2, k)
        assert torch.equal(Cd, Cd2)   # deterministic


def test_spmm_poisoned_unused_rows_of_b(ctx):
    """rows of B that no nonzero references may hold Inf / NaN: nothing of them may leak into C (long-row path too)"""
    rng = np.random.default_rng(7)
    m, n, k = 300, 9000, 128
    lens = np.full(m, 8); lens[5] = 3000; lens[77] = 1500
    ia = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    ja = np.concatenate([np.sort(rng.choice(np.arange(1, n), size=L, replace=False)) for L in lens]).astype(np.int64)
    a = rng.random(ja.size, dtype=np.float32)
    B = rng.random((n, k), dtype=np.float32)
    B[0] = np.inf   # column 0 is never referenced
    vals, idx, offs = csr_to_device(a, ia, ja)
    Cd = dev(np.zeros((m, k), np.float32))
    ctx.spmm("R", m, n, k, 1.0, vals, idx, offs, dev(B), k, 0.0, Cd, k)
    Bz = B.copy(); Bz[0] = 0
    assert oracle.rel_fro(Cd.cpu().numpy(), oracle.csrmm("N", m, n, k, 1.0, 0.0, a, ia, ja, "R", Bz, np.zeros((m, k), np.float32), acc64=True)) <= TOL


def test_spmv_transposed_is_deterministic_and_atomic_free_by_default(bof, ctx):
    """csrgemv 'T' (SURVEY K5): products sorted by column + fixed-order column sums -> bit-identical from run to run;
    bof_config.spmv_t_atomic = 1 opts into the red.global.add scatter (tolerance-equal)"""
    rng = np.random.default_rng(21)
    m, n = 30000, 26000
    a, ia, ja = ragged_csr(rng, m, n, 60)
    x = rng.random(m, dtype=np.float32)
    ref = oracle.csrgemv("T", m, n, a, ia, ja, x, acc64=True)
    shift = 987654321   # un-rebased offsets: a row-block slice of a larger matrix
    vals, idx, offs = csr_to_device(a, ia + shift, ja)
    runs = []
    for _ in range(3):
        y = torch.full((n,), float("nan"), device="cuda")
        ctx.spmv("T", m, n, vals, idx, offs, dev(x), y)
        runs.append(y)
        assert oracle.rel_fro(y.cpu().numpy(), ref) <= TOL
    assert torch.equal(runs[0], runs[1]) and torch.equal(runs[0], runs[2])
    with bof.Context(device=0, spmv_t_atomic=1) as c2:
        y = torch.full((n,), float("nan"), device="cuda")
        c2.spmv("T", m, n, vals, idx, offs, dev(x), y)
        assert oracle.rel_fro(y.cpu().numpy(), ref) <= TOL
    # host pipeline, several row blocks accumulated in block order
    with bof.Context(device=0, csrmm_max_nnz=200000) as c3:
        y1 = np.full(n, np.nan, np.float32); y2 = np.full(n, np.nan, np.float32)
        c3.host_csrgemv("T", m, n, a, ia, ja, x, y1)
        c3.host_csrgemv("T", m, n, a, ia, ja, x, y2)
        assert oracle.rel_fro(y1, ref) <= TOL and np.array_equal(y1, y2)


def test_host_csrmm_column_panels_child_process(tmp_path):
    """B wider than the device budget: flash::csrmm cuts B and C into column panels and re-streams A per panel
    (the reference's column blocks, src/blas/csrmm.cpp:64-126).  BOF_CSRMM_KPANEL forces the path at test size."""
    import subprocess, sys, textwrap
    code = textwrap.dedent('''
        import numpy as np, sys
        sys.path.insert(0, %r)
        import __graft_entry__ as g, oracle
        bof = g.load_package()
        rng = np.random.default_rng(4)
        m, n, k = 3000, 2500, 200
        a, ia, ja = oracle.gen_csr(m, n, 17, seed=9)
        B = oracle.gen_dense((n, k), seed=10); C0 = oracle.gen_dense((m, k), seed=11)
        with bof.Context(device=0, csrmm_max_nnz=9000) as ctx:
            for alpha, beta in ((1.0, 0.0), (1.5, 0.5)):
                C = C0.copy() if beta else np.full((m, k), np.nan, np.float32)
                ctx.host_csrmm("N", m, n, k, alpha, beta, a, ia, ja, "R", B, C)
                ref = oracle.csrmm("N", m, n, k, alpha, beta, a, ia, ja, "R", B, C0, acc64=True)
                assert oracle.rel_fro(C, ref) <= 1e-5, oracle.rel_fro(C, ref)
            assert ctx.stats().h2d_bytes > 0
        print("ok")
    ''' % str(Path(__file__).resolve().parent.parent))
    import os
    r = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, BOF_CSRMM_KPANEL="64"), capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "ok" in r.stdout, r.stdout + r.stderr
