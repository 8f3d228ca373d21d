"""GPU parity: stable CSR -> CSC transpose, bit-exact in offsets, indices and values."""
from pathlib import Path

import numpy as np
import pytest
import torch

import oracle
from gpu_util import csr_to_device, dev, ragged_csr

pytestmark = pytest.mark.gpu
G = np.load(Path(__file__).parent / "golden" / "golden_small.npz")


def run_device(ctx, m, n, a, ia, ja):
    nnz = int(ia[m] - ia[0])
    vals, idx, offs = csr_to_device(a, ia, ja)
    offs_t = torch.empty(n + 1, dtype=torch.int64, device="cuda")
    idx_t = torch.empty(max(nnz, 1), dtype=torch.int32, device="cuda")
    vals_t = torch.empty(max(nnz, 1), dtype=torch.float32, device="cuda")
    ctx.csr2csc(m, n, nnz, offs, idx, vals, offs_t, idx_t, vals_t)
    return offs_t.cpu().numpy(), idx_t.cpu().numpy()[:nnz].astype(np.int64), vals_t.cpu().numpy()[:nnz]


def assert_same(got, ref):
    assert np.array_equal(got[0], ref[0]), "offsets differ"
    assert np.array_equal(got[1], ref[1]), "indices differ"
    assert np.array_equal(got[2].view(np.uint32), ref[2].view(np.uint32)), "values differ (bitwise)"


@pytest.mark.parametrize("m,n,max_nnz", [(1000, 200, 30), (500, 70000, 90), (20000, 300, 12), (3, 5, 5),
                                         (4096, 4096, 64), (300, 17000000, 40)])
def test_csr2csc_vs_oracle(ctx, m, n, max_nnz):
    """n <= 4096 -> one radix pass; up to 2^24 columns -> two; 17e6 (25 bits) -> three."""
    rng = np.random.default_rng(m + n)
    a, ia, ja = ragged_csr(rng, m, n, max_nnz, dups=True)
    assert_same(run_device(ctx, m, n, a, ia, ja), oracle.csrcsc(m, n, ia, ja, a))


def test_csr2csc_golden(ctx):
    got = run_device(ctx, int(G["tr_m"]), int(G["tr_n"]), G["tr_a"], G["tr_ia"], G["tr_ja"])
    assert_same(got, (G["tr_ia_t"], G["tr_ja_t"], G["tr_a_t"]))


def test_csr2csc_empty_and_single(ctx):
    got = run_device(ctx, 7, 9, np.zeros(0, np.float32), np.zeros(8, np.int64), np.zeros(0, np.int64))
    assert np.array_equal(got[0], np.zeros(10, np.int64)) and got[1].size == 0
    got = run_device(ctx, 1, 1, np.array([2.5], np.float32), np.array([0, 1], np.int64), np.array([0], np.int64))
    assert np.array_equal(got[0], [0, 1]) and got[1][0] == 0 and got[2][0] == 2.5


def test_csr2csc_nan_payload_bits_preserved(ctx):
    """Values are moved bit-for-bit (no FP arithmetic): NaN payloads and -0.0 survive."""
    rng = np.random.default_rng(2)
    a, ia, ja = ragged_csr(rng, 400, 300, 20)
    bits = rng.integers(0, 2**32, size=a.size, dtype=np.uint64).astype(np.uint32)
    a = bits.view(np.float32)
    assert_same(run_device(ctx, 400, 300, a, ia, ja), oracle.csrcsc(400, 300, ia, ja, a))


def test_csr2csc_involution_at_scale(ctx):
    """Size-independent property on a cfg-4 shaped slice (2^20 x 2^20, 100 nnz/row, ~105M nnz):
    transposing twice restores the matrix bit for bit; offsets are a valid scan; per-column order sorted."""
    m = n = 1 << 20
    a, ia, ja = oracle.gen_csr(m, n, 100, seed=4)
    nnz = int(ia[m])
    vals, idx, offs = csr_to_device(a, ia, ja)
    o1 = torch.empty(n + 1, dtype=torch.int64, device="cuda"); i1 = torch.empty(nnz, dtype=torch.int32, device="cuda")
    v1 = torch.empty(nnz, dtype=torch.float32, device="cuda")
    ws = ctx.csr2csc_workspace(m, n, nnz)
    ctx.csr2csc(m, n, nnz, offs, idx, vals, o1, i1, v1, ws=ws)
    assert int(o1[0]) == 0 and int(o1[-1]) == nnz and bool((o1[1:] >= o1[:-1]).all())
    # column histogram == offsets diff
    assert torch.equal(torch.bincount(idx.long(), minlength=n), o1[1:] - o1[:-1])
    o2 = torch.empty(m + 1, dtype=torch.int64, device="cuda"); i2 = torch.empty_like(i1); v2 = torch.empty_like(v1)
    ctx.csr2csc(n, m, nnz, o1, i1, v1, o2, i2, v2, ws=ws)
    assert torch.equal(o2, offs) and torch.equal(i2, idx) and torch.equal(v2.view(torch.int32), vals.view(torch.int32))


def test_host_csrcsc(ctx):
    rng = np.random.default_rng(9)
    m, n = 6000, 9000
    a, ia, ja = ragged_csr(rng, m, n, 70, dups=True)
    nnz = int(ia[m])
    ia_t = np.zeros(n + 1, np.int64); ja_t = np.zeros(nnz, np.int64); a_t = np.zeros(nnz, np.float32)
    ctx.host_csrcsc(m, n, ia, ja, a, ia_t, ja_t, a_t)
    assert_same((ia_t, ja_t, a_t), oracle.csrcsc(m, n, ia, ja, a))


def run_on(c, m, n, a, ia, ja, offs_shift=0):
    """like run_device, on context `c`; offs_shift > 0 passes un-rebased offsets (offs[0] != 0) with the value /
    index pointers advanced accordingly -- a row-block slice of a larger matrix"""
    nnz = int(ia[m] - ia[0])
    vals, idx, offs = csr_to_device(a, ia + offs_shift, ja)
    offs_t = torch.empty(n + 1, dtype=torch.int64, device="cuda")
    idx_t = torch.full((max(nnz, 1),), -1, dtype=torch.int32, device="cuda")
    vals_t = torch.full((max(nnz, 1),), float("nan"), dtype=torch.float32, device="cuda")
    c.csr2csc(m, n, nnz, offs, idx, vals, offs_t, idx_t, vals_t)
    return offs_t.cpu().numpy(), idx_t.cpu().numpy()[:nnz].astype(np.int64), vals_t.cpu().numpy()[:nnz]


def torch_transpose(m, n, offs, idx, vals):
    """stable CSR -> CSC with torch on the device (checker for sizes the CPU oracle would take minutes on)"""
    rows = torch.repeat_interleave(torch.arange(m, device="cuda", dtype=torch.int32), (offs[1:] - offs[:-1]))
    order = torch.argsort(idx.long(), stable=True)
    cnt = torch.bincount(idx.long(), minlength=n)
    o = torch.zeros(n + 1, dtype=torch.int64, device="cuda")
    o[1:] = torch.cumsum(cnt, 0)
    return o, rows[order], vals[order]


@pytest.mark.parametrize("max_bits", [12, 8, 5, 3])
@pytest.mark.parametrize("m,n,max_nnz", [(3000, 5000, 40), (700, 1 << 20, 300), (9000, 4097, 25), (50, 60, 9)])
def test_csr2csc_digit_caps(bof, max_bits, m, n, max_nnz):
    """bof_config.radix_max_bits: 12 = one / two passes with the offsets gathered from the scanned histogram,
    8 and below = three and more passes (sorted keys of the last pass + segment offsets)"""
    rng = np.random.default_rng(m * 7 + n + max_bits)
    a, ia, ja = ragged_csr(rng, m, n, max_nnz, dups=True)
    with bof.Context(device=0, radix_max_bits=max_bits) as c:
        assert_same(run_on(c, m, n, a, ia, ja), oracle.csrcsc(m, n, ia, ja, a))


def test_csr2csc_long_rows_empty_spans_and_unrebased_offsets(ctx):
    """rows longer than a tile (8192 items), thousands of consecutive empty rows inside one tile, offs[0] != 0"""
    rng = np.random.default_rng(11)
    m, n = 40000, 30000
    counts = np.zeros(m, np.int64)
    counts[5] = 20000; counts[6] = 8192; counts[7] = 1; counts[30000] = 9000; counts[39999] = 3
    counts[rng.integers(100, 29000, 300)] = rng.integers(1, 50, 300)
    ia = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    ja = np.concatenate([np.sort(rng.choice(n, size=c, replace=False)) for c in counts if c]).astype(np.int64)
    a = rng.random(ja.size, dtype=np.float32)
    ref = oracle.csrcsc(m, n, ia, ja, a)
    assert_same(run_on(ctx, m, n, a, ia, ja), ref)
    assert_same(run_on(ctx, m, n, a, ia, ja, offs_shift=123456789), ref)


def test_csr2csc_multi_tile_supertiles(ctx):
    """45 M nonzeros: supertiles of several tiles (running cursors carried from tile to tile), two aligned passes
    for the 21-bit columns, a non-power-of-two column count; checked against a stable torch sort on the device"""
    m, n, nzr = 450_000, 1_500_001, 100
    gen = torch.Generator(device="cuda"); gen.manual_seed(5)
    idx = torch.sort(torch.randint(0, n, (m, nzr), device="cuda", generator=gen, dtype=torch.int32), dim=1).values.reshape(-1)
    idx[-1] = n - 1
    vals = torch.rand(m * nzr, device="cuda", generator=gen)
    offs = torch.arange(0, (m + 1) * nzr, nzr, dtype=torch.int64, device="cuda")
    nnz = m * nzr
    o1 = torch.empty(n + 1, dtype=torch.int64, device="cuda"); i1 = torch.empty(nnz, dtype=torch.int32, device="cuda")
    v1 = torch.empty(nnz, dtype=torch.float32, device="cuda")
    ctx.csr2csc(m, n, nnz, offs, idx, vals, o1, i1, v1)
    ro, ri, rv = torch_transpose(m, n, offs, idx, vals)
    assert torch.equal(o1, ro) and torch.equal(i1, ri) and torch.equal(v1.view(torch.int32), rv.view(torch.int32))


def test_csr2csc_skewed_columns(ctx):
    """most nonzeros in a handful of columns (one pass-1 bucket holds nearly everything)"""
    rng = np.random.default_rng(13)
    m, n = 60000, 70001
    hot = np.array([0, 1, 4096, 4097, 70000])
    ja = np.where(rng.random((m, 8)) < 0.9, hot[rng.integers(0, hot.size, (m, 8))], rng.integers(0, n, (m, 8)))
    ja = np.sort(ja, axis=1).reshape(-1).astype(np.int64)
    ia = np.arange(0, (m + 1) * 8, 8, dtype=np.int64)
    a = rng.random(ja.size, dtype=np.float32)
    assert_same(run_on(ctx, m, n, a, ia, ja), oracle.csrcsc(m, n, ia, ja, a))
