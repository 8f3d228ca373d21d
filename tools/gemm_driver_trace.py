#!/usr/bin/env python
"""drivers/gemm on /dev/shm files (pageable mmap path) at n^3 with BOF_TRACE=1; prints the driver's report and timeline."""
import os, subprocess, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import torch
n = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
d = Path("/dev/shm/bof"); d.mkdir(parents=True, exist_ok=True)
g = torch.Generator(device="cuda"); g.manual_seed(3)
for name in ("A.bin", "B.bin"):
    with open(d / name, "wb") as f:
        for r0 in range(0, n, 4096):
            f.write(torch.rand((min(4096, n - r0), n), device="cuda", generator=g).cpu().numpy().tobytes())
fd = os.open(d / "C.bin", os.O_RDWR | os.O_CREAT, 0o666); os.posix_fallocate(fd, 0, n * n * 4); os.close(fd)
torch.cuda.empty_cache()
for rep in range(2):
    r = subprocess.run([str(ROOT / "build" / "gemm"), d / "A.bin", d / "B.bin", d / "C.bin", str(n), str(n), str(n), "1.0", "0.0", "N", "N", "R", "0", "0", "0"],
                       capture_output=True, text=True, env=dict(os.environ, BOF_TRACE="1"))
    print(r.stdout.strip().splitlines()[-1])
    if rep == 1:
        print(r.stderr)
for name in ("A.bin", "B.bin", "C.bin"):
    (d / name).unlink()
