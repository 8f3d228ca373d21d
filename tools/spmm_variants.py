#!/usr/bin/env python
"""Sweep BOF_SPMM_VARIANT (bytes in flight per SM) on cfg-1 and a cfg-3 shaped slice; one process per variant."""
import json, os, subprocess, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
if len(sys.argv) > 1 and sys.argv[1] == "child":
    sys.path.insert(0, str(ROOT))
    import torch
    import __graft_entry__ as g
    from tools.bench_suite import gen_csr_gpu, time_gpu
    bof = g.load_package(); ctx = bof.Context(device=0)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    out = {"variant": int(os.environ.get("BOF_SPMM_VARIANT", "0"))}
    cases = (("cfg1", 262144, 64, int(os.environ.get("SPMM_K", "128"))),) if os.environ.get("ONLY_CFG1") else (("cfg1", 262144, 64, 128), ("cfg3_slice", 1 << 21, 100, 256))
    for name, m, nzr, k in cases:
        vals, idx, offs = gen_csr_gpu(m, m, nzr, 1)
        B = torch.rand((m, k), device="cuda"); C = torch.empty((m, k), device="cuda")
        t, tmin = time_gpu(lambda: ctx.spmm("R", m, m, k, 1.0, vals, idx, offs, B, k, 0.0, C, k), flush=flush)
        out[name + "_ms"] = t * 1e3
        out[name + "_gflops"] = 2.0 * m * nzr * k / t / 1e9
        del vals, idx, offs, B, C
    print(json.dumps(out), flush=True)
else:
    for v in (list(range(8)) if len(sys.argv) < 2 else [int(x) for x in sys.argv[1].split(',')]):
        env = dict(os.environ, BOF_SPMM_VARIANT=str(v))
        r = subprocess.run([sys.executable, __file__, "child"], env=env, capture_output=True, text=True)
        print(r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-500:], flush=True)
