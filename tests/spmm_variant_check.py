"""Run by test_gpu_sparse.py in a child process with BOF_SPMM_VARIANT set (the library reads it once per process):
checks the selected SpMM kernel variant against the oracle on aligned, ragged, shifted and very long rows."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import __graft_entry__ as g
import oracle
from gpu_util import csr_to_device, dev, ragged_csr


def uniform_csr(rng, m, n, nzr):
    ia = np.arange(m + 1, dtype=np.int64) * nzr
    ja = np.sort(rng.integers(0, n, size=(m, nzr)), axis=1).reshape(-1).astype(np.int64)
    return rng.random(m * nzr, dtype=np.float32) + 0.01, ia, ja


def main():
    bof = g.load_package()
    ctx = bof.Context(device=0)
    rng = np.random.default_rng(11)
    cases = []
    cases.append(("uniform64_k128", *uniform_csr(rng, 1003, 900, 64), 1003, 900, 128))
    cases.append(("uniform100_k256", *uniform_csr(rng, 517, 700, 100), 517, 700, 256))
    cases.append(("ragged_k128", *ragged_csr(rng, 777, 513, 40), 777, 513, 128))
    cases.append(("ragged_k384", *ragged_csr(rng, 300, 513, 90), 300, 513, 384))
    a, ia, ja = uniform_csr(rng, 37, 5000, 3000)  # 8 rows x 3000 nnz = 24000 > staging chunk: chunk loop
    cases.append(("long_rows_k128", a, ia, ja, 37, 5000, 128))
    worst = 0.0
    for name, a, ia, ja, m, n, k in cases:
        B = rng.random((n, k), dtype=np.float32)
        C0 = rng.random((m, k), dtype=np.float32)
        for r0, r1 in ((0, m), (3, m - 2)):  # second pass: offs[0] != 0 and a base that is not 16-byte aligned
            z0, z1 = int(ia[r0]), int(ia[r1])
            vals, idx, offs = dev(a[z0:z1]), dev(ja[z0:z1].astype(np.int32)), dev(ia[r0:r1 + 1])
            Cd = dev(C0[r0:r1])
            ctx.spmm("R", r1 - r0, n, k, 1.5, vals, idx, offs, dev(B), k, 0.5, Cd, k)
            ref = oracle.csrmm("N", m, n, k, 1.5, 0.5, a, ia, ja, "R", B, C0, acc64=True)[r0:r1]
            err = oracle.rel_fro(Cd.cpu().numpy(), ref)
            worst = max(worst, err)
            assert err <= 1e-5, (name, r0, err)
    print("SPMM_VARIANT_OK worst_rel_err=%.3g" % worst)


if __name__ == "__main__":
    main()
