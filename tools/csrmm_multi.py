#!/usr/bin/env python
"""cfg-3 (out-of-core csrmm) under torchrun: output row blocks sharded over the ranks (nnz-balanced), the
dense operand B replicated on every GPU, no collective on the data path.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29512 \
        tools/csrmm_multi.py [--rows 8388608] [--nnz-per-row 100] [--k 256]
"""
import argparse
import json
import os
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import __graft_entry__ as g  # noqa: E402
from tools.bench_suite import gen_csr_gpu, pinned_like  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=1 << 23)
    ap.add_argument("--nnz-per-row", type=int, default=100)
    ap.add_argument("--k", type=int, default=256)
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    g.load_package()
    from bof_b200 import dist as _bd
    _bd.bind_to_gpu_numa(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    bof = g.load_package()
    from bof_b200 import dist as bdist
    ctx = bof.Context(device=local)
    m = n = args.rows
    k, nzr = args.k, args.nnz_per_row
    # constant nnz per row => the nnz-balanced shard is the equal row shard; exercise the helper anyway
    ia_full = np.arange(0, (m + 1) * nzr, nzr, dtype=np.int64)
    r0, r1 = bdist.nnz_balanced_shard(ia_full, world, rank)
    rows = r1 - r0
    vals, idx, offs = gen_csr_gpu(rows, n, nzr, seed=1000 + rank)
    gen = torch.Generator(device="cuda"); gen.manual_seed(7)
    B = torch.rand((n, k), device="cuda", generator=gen)
    a_h, ja_h, ia_h, B_h = pinned_like(vals), pinned_like(idx, torch.int64), pinned_like(offs), pinned_like(B)
    C_h = torch.empty((rows, k), dtype=torch.float32, pin_memory=True)
    # device-side checksum for the shard
    Cd = torch.empty((rows, k), device="cuda")
    ctx.spmm("R", rows, n, k, 1.0, vals, idx, offs, B, k, 0.0, Cd, k)
    ref_sum = float(Cd.double().sum())
    del vals, idx, B, Cd
    torch.cuda.empty_cache()
    nnz = m * nzr
    Bdev = torch.empty((n, k), device="cuda") if world > 1 else None
    modes = ["replicated"] + (["allgather"] if world > 1 else [])
    for mode in modes:
        ts, up = [], 0
        for i in range(3):
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            if mode == "replicated":  # every rank uploads all of B over its own PCIe link
                ctx.host_csrmm("N", rows, n, k, 1.0, 0.0, a_h, ia_h, ja_h, "R", B_h, C_h)
            else:                     # every rank uploads 1/N of B, NVLink all-gather, B passed as a device pointer
                up = bdist.allgather_dense(B_h, Bdev)
                ctx.host_csrmm_devb(rows, n, k, 1.0, 0.0, a_h, ia_h, ja_h, Bdev, C_h)
            ts.append(time.perf_counter() - t0)
        st = ctx.stats()
        t = torch.tensor([min(ts[1:])], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        secs = float(t[0])
        err = abs(float(C_h.double().sum()) - ref_sum) / abs(ref_sum)
        h2d = st.h2d_bytes + up
        if rank == 0:
            print(json.dumps({"config": f"cfg3 csrmm {m}^2, {nzr} nnz/row, k={k}, {world} GPU(s), row blocks sharded, B {mode}",
                              "e2e_ms": secs * 1e3, "gflops_job": 2.0 * nnz * k / secs / 1e9,
                              "h2d_bytes_per_gpu": h2d, "d2h_bytes_per_gpu": st.d2h_bytes,
                              "pcie_h2d_bound_ms_at_55gbs": h2d / 55e9 * 1e3,
                              "ratio_to_pcie_bound": secs / (h2d / 55e9), "shard_sum_rel_err": err}), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
