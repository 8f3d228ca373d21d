#!/usr/bin/env python
"""Device-resident timing of bof_csr2csc on the cfg-4 matrix (2^23 x 2^23, 100 nnz/row) for several digit caps
(bof_config.radix_max_bits: 12 -> two passes, 8 -> three), with the involution / histogram property checks.
    python tools/bench_csrcsc.py [--rows 8388608] [--nzr 100] [--bits 12,11,8]"""
import argparse
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402

import __graft_entry__ as g  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=1 << 23)
    ap.add_argument("--nzr", type=int, default=100)
    ap.add_argument("--bits", default="12,8")
    ap.add_argument("--iters", type=int, default=3)
    args = ap.parse_args()
    bof = g.load_package()
    m = n = args.rows
    gen = torch.Generator(device="cuda"); gen.manual_seed(2)
    idx = torch.empty((m, args.nzr), dtype=torch.int32, device="cuda")
    for r0 in range(0, m, 1 << 18):
        r1 = min(m, r0 + (1 << 18))
        idx[r0:r1] = torch.sort(torch.randint(0, n, (r1 - r0, args.nzr), device="cuda", generator=gen, dtype=torch.int32), dim=1).values
    idx = idx.reshape(-1)
    nnz = m * args.nzr
    vals = torch.rand(nnz, device="cuda", generator=gen)
    offs = torch.arange(0, (m + 1) * args.nzr, args.nzr, dtype=torch.int64, device="cuda")
    o1 = torch.empty(n + 1, dtype=torch.int64, device="cuda"); i1 = torch.empty(nnz, dtype=torch.int32, device="cuda")
    v1 = torch.empty(nnz, device="cuda")
    hbm = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"] if (ROOT / "MEASURED_PEAKS.json").exists() else 6650.0
    for bits in [int(b) for b in args.bits.split(",")]:
        with bof.Context(device=0, radix_max_bits=bits) as c:
            ws = c.csr2csc_workspace(m, n, nnz)
            ts = []
            for i in range(args.iters + 1):
                e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
                e0.record(); c.csr2csc(m, n, nnz, offs, idx, vals, o1, i1, v1, ws=ws); e1.record()
                torch.cuda.synchronize()
                if i:
                    ts.append(e0.elapsed_time(e1))
            t = sum(ts) / len(ts)
            ok_hist = bool(torch.equal(torch.bincount(idx.long(), minlength=n), o1[1:] - o1[:-1]))
            o2 = torch.empty(m + 1, dtype=torch.int64, device="cuda"); i2 = torch.empty_like(i1); v2 = torch.empty_like(v1)
            c.csr2csc(n, m, nnz, o1, i1, v1, o2, i2, v2, ws=ws)
            same = bool(torch.equal(o2, offs) and torch.equal(i2, idx) and torch.equal(v2.view(torch.int32), vals.view(torch.int32)))
            ideal = nnz * 20 + (m + n + 2) * 8
            print(json.dumps({"radix_max_bits": bits, "ms": t, "ms_min": min(ts), "nnz": nnz, "gnnz_per_s": nnz / t / 1e6,
                              "ideal_gbs": ideal / t / 1e6, "frac_single_pass_ideal": ideal / t / 1e6 / hbm,
                              "histogram_matches_offsets": ok_hist, "double_transpose_bit_exact": same,
                              "workspace_gb": ws.numel() / 1e9}), flush=True)
            del ws, o2, i2, v2
            torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
