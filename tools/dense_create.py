#!/usr/bin/env python
"""Writes a dense fp32 matrix file as the reference's misc/dense_create.cpp:
    dense_create.py <file> <nrows> <ncols> <r|s|z> [--seed S]
  s: c[i] = i % 10 (dense_create.cpp:28-32)   z: zeros (:33-37)
  r: the reference mixes i with a thread-shared rand_r (:21-26, not reproducible); here U[0,1) from a seeded
     generator, the distribution misc/gemm_run.sh:20-21 uses for its accuracy runs."""
import argparse

import numpy as np


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("file"); ap.add_argument("nrows", type=int); ap.add_argument("ncols", type=int)
    ap.add_argument("mode", choices=["r", "s", "z"])
    ap.add_argument("--seed", type=int, default=0x5EED0002)
    a = ap.parse_args()
    n = a.nrows * a.ncols
    rng = np.random.default_rng(a.seed)
    with open(a.file, "wb") as f:
        for i0 in range(0, n, 1 << 26):
            cnt = min(1 << 26, n - i0)
            if a.mode == "s":
                x = (np.arange(i0, i0 + cnt, dtype=np.int64) % 10).astype(np.float32)
            elif a.mode == "z":
                x = np.zeros(cnt, np.float32)
            else:
                x = rng.random(cnt, dtype=np.float32)
            x.tofile(f)
    print(f"wrote {a.file}: {a.nrows} x {a.ncols} fp32, mode {a.mode}")


if __name__ == "__main__":
    main()
