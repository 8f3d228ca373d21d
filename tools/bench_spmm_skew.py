#!/usr/bin/env python
"""SpMM on a power-law (Zipf) row-length distribution: the row-splitting path against the plain row-group kernel
(BOF_SPMM_NO_SPLIT=1, a separate process because the switch is read once), same matrix, same B.
    python tools/bench_spmm_skew.py [--rows 262144] [--nnz 16777216] [--k 128]"""
import argparse
import json
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))


def run(args):
    import numpy as np
    import torch

    import __graft_entry__ as g
    bof = g.load_package()
    m = n = args.rows
    gen = torch.Generator(device="cuda"); gen.manual_seed(3)
    ranks = torch.randperm(m, device="cuda", generator=gen) + 1
    lens = torch.clamp((args.nnz / (ranks.double() * np.log(m))).long() + 1, max=n)
    offs = torch.zeros(m + 1, dtype=torch.int64, device="cuda"); offs[1:] = torch.cumsum(lens, 0)
    nnz = int(offs[-1])
    idx = torch.randint(0, n, (nnz,), device="cuda", generator=gen, dtype=torch.int32)   # duplicates allowed in CSR
    vals = torch.rand(nnz, device="cuda", generator=gen)
    B = torch.rand((n, args.k), device="cuda", generator=gen); C = torch.empty((m, args.k), device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    with bof.Context(device=0) as ctx:
        ts = []
        for i in range(8):
            flush.zero_()
            e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
            e0.record(); ctx.spmm("R", m, n, args.k, 1.0, vals, idx, offs, B, args.k, 0.0, C, args.k); e1.record()
            torch.cuda.synchronize()
            if i >= 3:
                ts.append(e0.elapsed_time(e1))
        # checksum of checksums (fp64) on the timed buffer
        colsum = torch.zeros(n, device="cuda", dtype=torch.float64).index_add_(0, idx.long(), vals.double())
        chk = float((C.double().sum() - (colsum * B.double().sum(1)).sum()).abs() / C.double().sum().abs())
        # the longest row exactly (fp64)
        r = int(torch.argmax(lens)); z0, z1 = int(offs[r]), int(offs[r + 1])
        ref = (vals[z0:z1].double()[:, None] * B[idx[z0:z1].long()].double()).sum(0)
        err = float((C[r].double() - ref).norm() / ref.norm())
    t = sum(ts) / len(ts)
    print(json.dumps({"split": os.environ.get("BOF_SPMM_NO_SPLIT") is None, "ms": t, "gflops": 2.0 * nnz * args.k / t / 1e6, "nnz": nnz,
                      "max_row": int(lens.max()), "rows_over_1024": int((lens > 1024).sum()), "median_row": int(lens.median()),
                      "checksum_rel_err": chk, "longest_row_rel_err": err}), flush=True)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=262144)
    ap.add_argument("--nnz", type=int, default=16777216)
    ap.add_argument("--k", type=int, default=128)
    ap.add_argument("--child", action="store_true")
    a = ap.parse_args()
    if a.child:
        run(a)
    else:
        for env in ({}, {"BOF_SPMM_NO_SPLIT": "1"}):
            subprocess.run([sys.executable, __file__, "--child", "--rows", str(a.rows), "--nnz", str(a.nnz), "--k", str(a.k)],
                           env=dict(os.environ, **env), check=False)
