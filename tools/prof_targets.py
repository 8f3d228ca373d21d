#!/usr/bin/env python
"""Launches each hot kernel a few times at a representative size; run under `ncu --set full -k regex:...`."""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402

import __graft_entry__ as g  # noqa: E402
from tools.bench_suite import gen_csr_gpu  # noqa: E402

which = set((sys.argv[1] if len(sys.argv) > 1 else "gemm,spmm,spmv,csrcsc,kmeans").split(","))
bof = g.load_package()
ctx = bof.Context(device=0)
if "gemm" in which:
    n = 8192
    A = torch.rand((n, n), device="cuda"); B = torch.rand((n, n), device="cuda"); C = torch.empty((n, n), device="cuda")
    ws = ctx.sgemm_workspace(n, n, n)
    for _ in range(2):
        ctx.sgemm("R", "N", "N", n, n, n, 1.0, A, 0, B, 0, 0.0, C, 0, ws=ws)
    torch.cuda.synchronize()
    del A, B, C, ws
if "gemm32k" in which or "gemm32k_nosync" in which or "gemm32k_hyb" in which:
    n = 32768
    c2 = bof.Context(device=0, gemm_wave_sync=-1 if "gemm32k_nosync" in which else 0,
                     gemm_split=2 if "gemm32k_hyb" in which else 1)
    A = torch.rand((n, n), device="cuda"); B = torch.rand((n, n), device="cuda"); C = torch.empty((n, n), device="cuda")
    ws = c2.sgemm_workspace(n, n, n)
    for _ in range(2):
        c2.sgemm("R", "N", "N", n, n, n, 1.0, A, 0, B, 0, 0.0, C, 0, ws=ws)
    torch.cuda.synchronize()
    del A, B, C, ws
    c2.close()
if "spmm" in which:
    m = n = 262144
    vals, idx, offs = gen_csr_gpu(m, n, 64, 1)
    for k in (128, 256):
        B = torch.rand((n, k), device="cuda"); C = torch.empty((m, k), device="cuda")
        for _ in range(2):
            ctx.spmm("R", m, n, k, 1.0, vals, idx, offs, B, k, 0.0, C, k)
    torch.cuda.synchronize()
if "spmv" in which:
    m = n = 1 << 21
    vals, idx, offs = gen_csr_gpu(m, n, 100, 2)
    x = torch.rand(n, device="cuda"); y = torch.empty(m, device="cuda")
    for tr in "NNTT":
        ctx.spmv(tr, m, n, vals, idx, offs, x, y)
    torch.cuda.synchronize()
if "csrcsc" in which:
    m = n = 1 << 20
    vals, idx, offs = gen_csr_gpu(m, n, 100, 3)
    nnz = m * 100
    o = torch.empty(n + 1, dtype=torch.int64, device="cuda"); i = torch.empty(nnz, dtype=torch.int32, device="cuda")
    v = torch.empty(nnz, device="cuda")
    ctx.csr2csc(m, n, nnz, offs, idx, vals, o, i, v)
    torch.cuda.synchronize()
if "kmeans" in which:
    P, K, d = 1 << 20, 1024, 256
    pts = torch.randn((P, d), device="cuda"); cent = pts[:K].clone()
    p2 = torch.empty(P, device="cuda"); c2 = torch.empty(K, device="cuda")
    ctx.row_sqnorm(P, d, pts, d, p2); ctx.row_sqnorm(K, d, cent, d, c2)
    asg = torch.empty(P, dtype=torch.int32, device="cuda")
    planes = ctx.kmeans_prepare_points(P, d, pts)
    for _ in range(2):
        ctx.kmeans_assign(P, K, d, pts, cent, c2, p2, asg, planes=planes)
    sums = torch.empty((K, d), device="cuda"); cnt = torch.empty(K, device="cuda")
    ctx.kmeans_reduce(P, K, d, pts, asg, sums, cnt)
    ctx.kmeans_finalize(K, d, sums, cnt, cent, c2)
    torch.cuda.synchronize()
ctx.close()
