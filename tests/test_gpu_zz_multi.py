"""GPU parity of the multi-GPU paths (needs >= 2 GPUs; skipped on a single-GPU box): one process driving several
GPUs through bof_mgpu_* (what the C++ flash:: adapters use with BOF_GPUS > 1), and one process per GPU under
torchrun through bof_comm_init + bof_dist_* (tools/dist_check.py).  World-size-1 behaviour of the same entry points
is covered on any GPU box.  The file name sorts last on purpose: under `pytest -x` a problem on a multi-GPU box
cannot hide the single-GPU parity files."""
import os
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

import oracle

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent
TOL = 1e-5
NGPU = torch.cuda.device_count() if torch.cuda.is_available() else 0
need2 = pytest.mark.skipif(NGPU < 2, reason="needs at least 2 GPUs")


class _Done:
    def __init__(self, returncode, stdout, stderr):
        self.returncode, self.stdout, self.stderr = returncode, stdout, stderr


def run_bounded(cmd, timeout, env=None):
    """subprocess.run with a deadline that ends the WHOLE process tree, so that workers blocked in a collective
    cannot outlive the test and keep the GPUs busy for whatever runs next.  The child gets its own session; on a
    time-out its group first gets SIGTERM -- torchrun starts every worker in a session of its own and only its
    SIGTERM handler reaches them (elastic SubprocessHandler.close -> killpg per worker) -- then SIGKILL."""
    import signal
    p = subprocess.Popen(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, start_new_session=True)
    try:
        out, err = p.communicate(timeout=timeout)
    except subprocess.TimeoutExpired:
        for sig, grace in ((signal.SIGTERM, 45), (signal.SIGKILL, 10)):
            try:
                os.killpg(p.pid, sig)   # pid == pgid == sid of the child
            except ProcessLookupError:
                pass
            try:
                out, err = p.communicate(timeout=grace)
                break
            except subprocess.TimeoutExpired:
                continue
        else:
            out, err = "", ""
        return _Done(-9, out, (err or "") + "\n[run_bounded] ended the process tree after %d s" % timeout)
    return _Done(p.returncode, out, err)



def test_dist_entry_points_at_world_one(bof, ctx):
    """without a communicator bof_dist_* are the plain pipelines and bof_kmeans_allreduce is a no-op"""
    assert ctx.comm_world() == 1
    M, N, K = 700, 500, 900
    a, b = oracle.gen_dense((M, K), seed=1), oracle.gen_dense((K, N), seed=2)
    c = np.full((M, N), np.nan, np.float32)
    ctx.dist_gemm("N", "N", M, N, K, 1.0, 0.0, a, b, c)
    assert oracle.rel_fro(c, oracle.gemm("R", "N", "N", M, N, K, 1.0, 0.0, a, b, np.zeros_like(c), acc64=True)) <= TOL
    m, n, k = 3000, 2000, 64
    av, ia, ja = oracle.gen_csr(m, n, 16, seed=3)
    B = oracle.gen_dense((n, k), seed=4)
    C = np.full((m, k), np.nan, np.float32)
    ctx.dist_csrmm(m, n, k, 1.0, 0.0, av, ia, ja, B, C)
    assert oracle.rel_fro(C, oracle.csrmm("N", m, n, k, 1.0, 0.0, av, ia, ja, "R", B, np.zeros_like(C), acc64=True)) <= TOL


def test_slab_download_path_small_m(ctx):
    """M small enough that every row block rides in the panel prologue: C leaves the device slab by slab"""
    M, N, K = 1024, 40960, 2048     # B = 320 MiB -> panelled; one row block
    a, b = oracle.gen_dense((M, K), seed=5), oracle.gen_dense((K, N), seed=6)
    c0 = oracle.gen_dense((M, N), seed=7)
    c = c0.copy()
    ctx.host_gemm("R", "N", "N", M, N, K, 1.0, 0.5, a, b, c)
    jj = np.random.default_rng(1).integers(0, N, 512)
    ref = a.astype(np.float64) @ b[:, jj].astype(np.float64) + 0.5 * c0[:, jj]
    assert np.linalg.norm(c[:, jj] - ref) / np.linalg.norm(ref) <= TOL


def test_kmeans_count_split_is_exact(bof, ctx):
    """cluster sizes travel as (size & 4095, size >> 12): the centroid of a cluster of > 4096 points is exact"""
    P, K, d = 20000, 3, 8
    pts = np.zeros((P, d), np.float32); pts[:, 0] = 1.0
    pts[15000:, 0] = 100.0
    c0 = np.array([[1.0] + [0] * (d - 1), [100.0] + [0] * (d - 1), [1000.0] + [0] * (d - 1)], np.float32)
    km = bof.KMeans(ctx, P, K, d, pts, c0)
    km.lloyd(2)
    c = np.zeros((K, d), np.float32); a = np.zeros(P, np.int64)
    km.get(c, a); km.close()
    assert np.allclose(c, c0 * np.array([[1], [1], [0]], np.float32), rtol=1e-6, atol=0)   # 15000 and 5000 points; empty cluster -> 0
    assert (a[:15000] == 0).all() and (a[15000:] == 1).all()


@need2
def test_mgpu_one_process():
    """tools/mgpu_check.py in a child process with a deadline: a missed hand-shake between the per-GPU threads shows up
    as a failure here instead of stalling the whole suite"""
    r = run_bounded([sys.executable, str(ROOT / "tools" / "mgpu_check.py")], 240)
    assert r.returncode == 0 and "mgpu ok" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


@need2
def test_dist_paths_under_torchrun():
    n = min(NGPU, 4)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = run_bounded([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
                        "--master-port", "29533", str(ROOT / "tools" / "dist_check.py")], 240, env=env)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert '"comm_world": %d' % n in r.stdout


@need2
def test_cpp_drivers_with_bof_gpus(tmp_path):
    """BOF_GPUS=n: the unchanged C++ drivers (reference CLIs, file-backed flash_ptrs) spread over the GPUs"""
    import __graft_entry__ as g
    binp = ROOT / "build"
    if not (binp / "gemm").exists():
        g.build()
    env = dict(os.environ, BOF_GPUS=str(min(NGPU, 4)))
    rng = np.random.default_rng(3)
    M, N, K = 3000, 2200, 1600
    A = rng.random((M, K), dtype=np.float32); B = rng.random((K, N), dtype=np.float32); C0 = rng.random((M, N), dtype=np.float32)
    A.tofile(tmp_path / "A.bin"); B.tofile(tmp_path / "B.bin"); C0.tofile(tmp_path / "C.bin")
    r = run_bounded([str(binp / "gemm"), *map(str, (tmp_path / "A.bin", tmp_path / "B.bin", tmp_path / "C.bin", M, K, N, 1.0, 0.5,
                                                   "N", "N", "R", 0, 0, 0))], 240, env=env)
    assert r.returncode == 0 and "returned 0" in r.stdout, r.stdout + r.stderr
    got = np.fromfile(tmp_path / "C.bin", dtype=np.float32).reshape(M, N)
    assert oracle.rel_fro(got, oracle.gemm("R", "N", "N", M, N, K, 1.0, 0.5, A, B, C0, acc64=True)) <= TOL
    m, n, k = 40000, 30000, 64
    av, ia, ja = oracle.gen_csr(m, n, 18, seed=5)
    Bd = oracle.gen_dense((n, k), seed=6); Cd = oracle.gen_dense((m, k), seed=7)
    av.tofile(tmp_path / "A.csr"); ja.tofile(tmp_path / "A.col"); ia.tofile(tmp_path / "A.off")
    Bd.tofile(tmp_path / "B2.bin"); Cd.tofile(tmp_path / "C2.bin")
    r = run_bounded([str(binp / "csrmm"), *map(str, (tmp_path / "A.csr", tmp_path / "A.col", tmp_path / "A.off", tmp_path / "B2.bin",
                                                    tmp_path / "C2.bin", m, n, k, 1.0, 0.5, "N", "R"))], 240, env=env)
    assert r.returncode == 0 and "returned 0" in r.stdout, r.stdout + r.stderr
    got = np.fromfile(tmp_path / "C2.bin", dtype=np.float32).reshape(m, k)
    assert oracle.rel_fro(got, oracle.csrmm("N", m, n, k, 1.0, 0.5, av, ia, ja, "R", Bd, Cd, acc64=True)) <= TOL
