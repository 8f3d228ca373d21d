#!/usr/bin/env python
"""Launches the kernels of every BASELINE config once (after one warm-up) at the bench sizes; run under ncu:
  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none ...   (traffic per launch)
  ncu --set full --import-source on --clock-control none -k regex:<kernel> -c 1 ...                          (one full capture)
    python tools/prof_r02.py cfg1,cfg3,cfg4,cfg5,gemm,skew [--scale 1.0]
cudaProfilerStart/Stop bracket the measured launches (use --profile-from-start off)."""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import __graft_entry__ as g  # noqa: E402
from tools.bench_configs import gen_csr_gpu  # noqa: E402

which = set((sys.argv[1] if len(sys.argv) > 1 else "cfg1,cfg3,cfg4,cfg5,gemm").split(","))
scale = float(sys.argv[sys.argv.index("--scale") + 1]) if "--scale" in sys.argv else 1.0
bof = g.load_package()
ctx = bof.Context(device=0)
dev = "cuda:0"
prof = torch.cuda.profiler


def measured(fn, warm=1):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    prof.start()
    fn()
    torch.cuda.synchronize()
    prof.stop()


if "cfg1" in which:
    m = n = 262144
    vals, idx, offs = gen_csr_gpu(m, n, 64, 1, dev)
    B = torch.rand((n, 128), device=dev); C = torch.empty((m, 128), device=dev)
    measured(lambda: ctx.spmm("R", m, n, 128, 1.0, vals, idx, offs, B, 128, 0.0, C, 128), warm=2)
    del vals, idx, offs, B, C
if "skew" in which:
    m = n = 262144
    gen = torch.Generator(device=dev); gen.manual_seed(3)
    ranks = torch.randperm(m, device=dev, generator=gen) + 1
    lens = torch.clamp((16777216 / (ranks.double() * np.log(m))).long() + 1, max=n)
    offs = torch.zeros(m + 1, dtype=torch.int64, device=dev); offs[1:] = torch.cumsum(lens, 0)
    nnz = int(offs[-1])
    idx = torch.randint(0, n, (nnz,), device=dev, generator=gen, dtype=torch.int32); vals = torch.rand(nnz, device=dev, generator=gen)
    B = torch.rand((n, 128), device=dev); C = torch.empty((m, 128), device=dev)
    measured(lambda: ctx.spmm("R", m, n, 128, 1.0, vals, idx, offs, B, 128, 0.0, C, 128), warm=2)
    del vals, idx, offs, B, C
if "cfg3" in which or "cfg4" in which:
    m = n = int((1 << 23) * scale)
    vals, idx, offs = gen_csr_gpu(m, n, 100, 3, dev)
    nnz = m * 100
    if "cfg3" in which:
        B = torch.rand((n, 256), device=dev); C = torch.empty((m, 256), device=dev)
        measured(lambda: ctx.spmm("R", m, n, 256, 1.0, vals, idx, offs, B, 256, 0.0, C, 256))
        del B, C
    if "cfg4" in which:
        x = torch.rand(n, device=dev); y = torch.empty(m, device=dev); xt = torch.rand(m, device=dev); yt = torch.empty(n, device=dev)
        measured(lambda: ctx.spmv("N", m, n, vals, idx, offs, x, y))
        measured(lambda: ctx.spmv("T", m, n, vals, idx, offs, xt, yt))
        o1 = torch.empty(n + 1, dtype=torch.int64, device=dev); i1 = torch.empty(nnz, dtype=torch.int32, device=dev); v1 = torch.empty(nnz, device=dev)
        ws = ctx.csr2csc_workspace(m, n, nnz)
        measured(lambda: ctx.csr2csc(m, n, nnz, offs, idx, vals, o1, i1, v1, ws=ws))
        del x, y, xt, yt, o1, i1, v1, ws
    del vals, idx, offs
    torch.cuda.empty_cache()
if "cfg5" in which:
    P, K, d = int(10_000_000 * scale), 1024, 256
    gen = torch.Generator(device=dev); gen.manual_seed(5)
    cent = torch.randn((K, d), device=dev, generator=gen) * 4
    pts = torch.empty((P, d), device=dev)
    for q0 in range(0, P, 1 << 20):
        q1 = min(P, q0 + (1 << 20))
        pts[q0:q1] = cent[torch.randint(0, K, (q1 - q0,), device=dev, generator=gen)] + 0.5 * torch.randn((q1 - q0, d), device=dev, generator=gen)
    c0 = (cent + 0.25 * torch.randn((K, d), device=dev, generator=gen))
    p2 = torch.empty(P, device=dev); c2 = torch.empty(K, device=dev)
    ctx.row_sqnorm(P, d, pts, d, p2); ctx.row_sqnorm(K, d, c0, d, c2)
    asg = torch.empty(P, dtype=torch.int32, device=dev)
    planes = ctx.kmeans_prepare_points(P, d, pts)
    sums = torch.empty((K, d), device=dev); cnt = torch.empty(K, device=dev)

    def it():
        ctx.kmeans_assign(P, K, d, pts, c0, c2, p2, asg, planes=planes)
        ctx.kmeans_reduce(P, K, d, pts, asg, sums, cnt)
        ctx.kmeans_finalize(K, d, sums, cnt, c0, c2)
    measured(it)
    del pts, planes, asg
    torch.cuda.empty_cache()
if "gemm" in which:
    nn = int(32768 * min(scale, 1.0))
    A = torch.rand((nn, nn), device=dev); B = torch.rand((nn, nn), device=dev); C = torch.empty((nn, nn), device=dev)
    ws = ctx.sgemm_workspace(nn, nn, nn)
    measured(lambda: ctx.sgemm("R", "N", "N", nn, nn, nn, 1.0, A, 0, B, 0, 0.0, C, 0, ws=ws))
ctx.close()
