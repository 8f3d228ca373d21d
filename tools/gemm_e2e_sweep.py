#!/usr/bin/env python
"""End-to-end bof_host_gemm at 32768^3 from pinned host buffers for a few row-block sizes (one process per setting)."""
import json, os, subprocess, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
if len(sys.argv) > 1 and sys.argv[1] == "child":
    sys.path.insert(0, str(ROOT))
    import torch
    import __graft_entry__ as g
    bof = g.load_package()
    rb = int(os.environ["RB"]); n = 32768
    ctx = bof.Context(device=0, gemm_row_block=rb)
    gen = torch.Generator(device="cuda"); gen.manual_seed(1)
    A = torch.empty((n, n), dtype=torch.float32).pin_memory(); B = torch.empty((n, n), dtype=torch.float32).pin_memory()
    C = torch.empty((n, n), dtype=torch.float32).pin_memory()
    for H in (A, B):
        for r0 in range(0, n, 4096):
            H[r0:r0 + 4096].copy_(torch.rand((4096, n), device="cuda", generator=gen))
    ts = []
    for _ in range(4):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        ctx.host_gemm("R", "N", "N", n, n, n, 1.0, 0.0, A, B, C)
        ts.append(time.perf_counter() - t0)
    print(json.dumps({"row_block": rb, "q_panels": os.environ.get("BOF_GEMM_QPANELS", "8"), "ms": [round(t * 1e3, 1) for t in ts],
                      "best_ms": min(ts[1:]) * 1e3, "tflops": 2.0 * n ** 3 / min(ts[1:]) / 1e12}), flush=True)
else:
    for rb, qp in ((4096, 8), (2048, 8), (1024, 8), (8192, 8), (4096, 4), (2048, 4)):
        env = dict(os.environ, RB=str(rb), BOF_GEMM_QPANELS=str(qp))
        r = subprocess.run([sys.executable, __file__, "child"], env=env, capture_output=True, text=True)
        print(r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-800:], flush=True)
