#!/usr/bin/env python
"""Writes a synthetic CSR matrix in the reference's on-disk format (misc/sparse_create.cpp:23-29,63-81):
    <name>csr  fp32 values         <name>col  int64 column indices
    <name>off  int64 row offsets   <name>info "<nrows> <ncols> <sparsity>"
CLI as the reference tool: sparse_create.py <name> <nrows> <ncols> <sparsity> [--values compat|uniform] [--seed S]
nnz per row = ceil(ncols * sparsity) (as the reference), columns sorted and unique per row.  Values:
`compat` = (i % 9) + 1 (misc/sparse_create.cpp:52-55), `uniform` = U[0,1) fp32.  Columns come from a seeded
numpy generator instead of rand_r (the reference's per-row rand_r draw is biased and unspecified in
evaluation order)."""
import argparse
import math

import numpy as np


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("name"); ap.add_argument("nrows", type=int); ap.add_argument("ncols", type=int)
    ap.add_argument("sparsity", type=float)
    ap.add_argument("--values", default="compat", choices=["compat", "uniform"])
    ap.add_argument("--seed", type=int, default=0x5EED0001)
    a = ap.parse_args()
    assert a.sparsity < 1.0
    nzr = math.ceil(a.ncols * a.sparsity)
    nnz = a.nrows * nzr
    open(a.name + "info", "w").write(f"{a.nrows} {a.ncols} {a.sparsity}\n")
    np.arange(0, nnz + 1, nzr, dtype=np.int64).tofile(a.name + "off")
    rng = np.random.default_rng(a.seed)
    chunk = max(1, (1 << 24) // max(nzr, 1))
    with open(a.name + "col", "wb") as fc, open(a.name + "csr", "wb") as fv:
        for r0 in range(0, a.nrows, chunk):
            rows = min(chunk, a.nrows - r0)
            if nzr * 4 <= a.ncols:  # sparse rows: draw, sort, re-draw the rare duplicates
                cols = np.sort(rng.integers(0, a.ncols, size=(rows, nzr), dtype=np.int64), axis=1)
                while True:
                    dup = np.zeros_like(cols, dtype=bool)
                    dup[:, 1:] = cols[:, 1:] == cols[:, :-1]
                    if not dup.any():
                        break
                    cols[dup] = rng.integers(0, a.ncols, size=int(dup.sum()), dtype=np.int64)
                    cols.sort(axis=1)
            else:  # dense-ish rows: sample without replacement row by row
                cols = np.stack([np.sort(rng.choice(a.ncols, nzr, replace=False)) for _ in range(rows)]).astype(np.int64)
            cols.tofile(fc)
            i = np.arange(r0 * nzr, (r0 + rows) * nzr, dtype=np.int64)
            vals = ((i % 9) + 1).astype(np.float32) if a.values == "compat" else rng.random(i.size, dtype=np.float32)
            vals.tofile(fv)
    print(f"wrote {a.name}{{csr,col,off,info}}: {a.nrows} x {a.ncols}, {nzr} nnz/row, {nnz} nnz")


if __name__ == "__main__":
    main()
