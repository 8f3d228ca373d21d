#!/usr/bin/env python
"""cfg-3 csrmm through bof_host_csrmm from pinned host buffers, three calls; run with BOF_TRACE=1 for the timeline."""
import json, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import torch
import __graft_entry__ as g
from tools.bench_suite import gen_csr_gpu, pinned_like
bof = g.load_package(); ctx = bof.Context(device=0)
m = n = 1 << 23; nzr = 100; k = 256
vals, idx, offs = gen_csr_gpu(m, n, nzr, seed=3)
a_h, ja_h, ia_h = pinned_like(vals), pinned_like(idx, torch.int64), pinned_like(offs)
del vals, idx, offs; torch.cuda.empty_cache()
B_h = torch.rand((n, k), dtype=torch.float32).pin_memory(); C_h = torch.empty((m, k), dtype=torch.float32).pin_memory()
ts = []
for _ in range(3):
    t0 = time.perf_counter(); ctx.host_csrmm("N", m, n, k, 1.0, 0.0, a_h, ia_h, ja_h, "R", B_h, C_h); ts.append(time.perf_counter() - t0)
print(json.dumps({"ms": [round(t * 1e3, 1) for t in ts]}))
