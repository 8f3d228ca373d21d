#!/usr/bin/env python
"""Out-of-core runs through the C++ drop-in layer: files in /dev/shm (page-cache resident, so the numbers
isolate our pipeline from the box's disk), mapped to flash_ptrs by the drivers built from drivers/*.cpp,
streamed through the pinned staging ring.  Prints one JSON record per run.

    python tools/driver_bench.py [--dir /dev/shm/bof] [--rows 2097152] [--gemm 16384]
"""
import argparse
import json
import os
import re
import subprocess
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from tools.bench_suite import gen_csr_gpu  # noqa: E402

BIN = ROOT / "build"


def run(exe, *args, env=None):
    t0 = time.perf_counter()
    r = subprocess.run([str(BIN / exe), *map(str, args)], capture_output=True, text=True, env=env)
    wall = time.perf_counter() - t0
    if r.returncode != 0:
        raise RuntimeError(f"{exe} failed: {r.stdout} {r.stderr}")
    m = re.search(r"took ([0-9.]+) s", r.stdout)
    return float(m.group(1)), wall, r.stdout.strip().splitlines()[-1]


def prealloc(path, nbytes):
    """Output files are pre-allocated, as misc/gemm_run.sh:17-18 does with `fallocate -l` (a sparse file would
    charge page allocation to the first write)."""
    fd = os.open(path, os.O_RDWR | os.O_CREAT, 0o666)
    try:
        os.posix_fallocate(fd, 0, nbytes)
    finally:
        os.close(fd)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dir", default="/dev/shm/bof")
    ap.add_argument("--rows", type=int, default=1 << 21)
    ap.add_argument("--nnz-per-row", type=int, default=100)
    ap.add_argument("--k", type=int, default=256)
    ap.add_argument("--gemm", type=int, default=16384)
    a = ap.parse_args()
    d = Path(a.dir); d.mkdir(parents=True, exist_ok=True)
    recs = []
    try:
        # ---- csrmm ----
        m = n = a.rows
        vals, idx, offs = gen_csr_gpu(m, n, a.nnz_per_row, seed=11)
        B = torch.rand((n, a.k), device="cuda")
        vals.cpu().numpy().tofile(d / "A.csr"); idx.cpu().numpy().astype(np.int64).tofile(d / "A.col")
        offs.cpu().numpy().tofile(d / "A.off"); B.cpu().numpy().tofile(d / "B.bin")
        prealloc(d / "C.bin", m * a.k * 4)
        colsum = torch.zeros(n, device="cuda", dtype=torch.float64).index_add_(0, idx.long(), vals.double())
        want = float((colsum * B.double().sum(1)).sum())
        nnz = m * a.nnz_per_row
        del vals, idx, B
        torch.cuda.empty_cache()
        secs, wall, line = run("csrmm", d / "A.csr", d / "A.col", d / "A.off", d / "B.bin", d / "C.bin", m, n, a.k, 1.0, 0.0, "N", "R")
        got = float(np.fromfile(d / "C.bin", dtype=np.float32).astype(np.float64).sum())
        streamed = nnz * 12 + (m + 1) * 8 + n * a.k * 4 + m * a.k * 4
        recs.append({"config": f"drivers/csrmm on /dev/shm files: {m}^2, {a.nnz_per_row} nnz/row, k={a.k} (pageable mmap -> pinned ring -> H2D)",
                     "flash_csrmm_s": secs, "process_wall_s": wall, "gflops": 2.0 * nnz * a.k / secs / 1e9,
                     "streamed_gbs": streamed / secs / 1e9, "checksum_rel_err": abs(got - want) / abs(want), "driver_says": line})
        print(json.dumps(recs[-1]), flush=True)
        # ---- csrcsc on the same matrix ----
        secs, wall, line = run("csrcsc", d / "A.csr", d / "A.col", d / "A.off", d / "T.csr", d / "T.col", d / "T.off", m, n)
        offs_t = np.fromfile(d / "T.off", dtype=np.int64)
        recs.append({"config": f"drivers/csrcsc on /dev/shm files: {m}^2, {a.nnz_per_row} nnz/row", "flash_csrcsc_s": secs,
                     "process_wall_s": wall, "file_gbs_both_ways": 2 * (nnz * 12 + (n + 1) * 8) / secs / 1e9,
                     "offsets_end_equals_nnz": bool(offs_t[-1] == nnz and offs_t[0] == 0), "driver_says": line})
        print(json.dumps(recs[-1]), flush=True)
        if os.environ.get("BOF_AB"):
            env = dict(os.environ, BOF_NO_FD="1")
            secs, wall, line = run("csrmm", d / "A.csr", d / "A.col", d / "A.off", d / "B.bin", d / "C.bin", m, n, a.k, 1.0, 0.0, "N", "R", env=env)
            recs.append({"config": "drivers/csrmm again with BOF_NO_FD=1 (memcpy through the mapping)", "flash_csrmm_s": secs, "driver_says": line})
            print(json.dumps(recs[-1]), flush=True)
            secs, wall, line = run("csrmm", d / "A.csr", d / "A.col", d / "A.off", d / "B.bin", d / "C.bin", m, n, a.k, 1.0, 0.0, "N", "R")
            recs.append({"config": "drivers/csrmm third run (fd path, C file now fully allocated)", "flash_csrmm_s": secs, "driver_says": line})
            print(json.dumps(recs[-1]), flush=True)
        for f in ("A.csr", "A.col", "A.off", "B.bin", "C.bin", "T.csr", "T.col", "T.off"):
            (d / f).unlink(missing_ok=True)
        # ---- gemm ----
        g = a.gemm
        A = torch.rand((g, g), device="cuda"); Bm = torch.rand((g, g), device="cuda")
        A.cpu().numpy().tofile(d / "GA.bin"); Bm.cpu().numpy().tofile(d / "GB.bin")
        prealloc(d / "GC.bin", g * g * 4)
        want = float((A.double().sum(0) * Bm.double().sum(1)).sum())
        del A, Bm
        torch.cuda.empty_cache()
        secs, wall, line = run("gemm", d / "GA.bin", d / "GB.bin", d / "GC.bin", g, g, g, 1.0, 0.0, "N", "N", "R", 0, 0, 0)
        got = float(np.fromfile(d / "GC.bin", dtype=np.float32).astype(np.float64).sum())
        recs.append({"config": f"drivers/gemm on /dev/shm files: {g}^3 fp32 (pageable mmap -> pinned ring -> H2D)", "flash_gemm_s": secs,
                     "process_wall_s": wall, "tflops": 2.0 * g ** 3 / secs / 1e12, "file_gbs": 3 * g * g * 4 / secs / 1e9,
                     "checksum_rel_err": abs(got - want) / abs(want), "driver_says": line})
        print(json.dumps(recs[-1]), flush=True)
    finally:
        for f in d.glob("*"):
            f.unlink()
    out = ROOT / "gpurun_out" / "driver_bench.json"
    out.parent.mkdir(exist_ok=True)
    out.write_text(json.dumps(recs, indent=1))


if __name__ == "__main__":
    main()
