"""Helpers shared by the GPU parity tests (device buffers via torch; everything computed through the C ABI)."""
import numpy as np
import torch


def dev(x):
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


def csr_to_device(a, ia, ja):
    """(vals f32, idx i32, offs i64) on the device -- the layout of the device-tile kernels."""
    return dev(a.astype(np.float32)), dev(ja.astype(np.int32)), dev(ia.astype(np.int64))


def ragged_csr(rng, m, n, max_nnz, dups=False):
    counts = rng.integers(0, max_nnz + 1, size=m)
    counts[rng.random(m) < 0.15] = 0  # empty rows
    counts = np.minimum(counts, n)
    ia = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    ja = np.zeros(int(ia[-1]), np.int64)
    for r in range(m):
        cols = np.sort(rng.choice(n, size=counts[r], replace=False))
        if dups and counts[r] > 2 and r % 3 == 0:
            cols[1] = cols[0]
        ja[ia[r]:ia[r + 1]] = cols
    a = rng.random(int(ia[-1]), dtype=np.float32) + 0.01
    return a, ia, ja
