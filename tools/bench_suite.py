#!/usr/bin/env python
"""Per-config measurements for every row of SURVEY.md section 8 (cfg-1 .. cfg-5), one JSON record each.

    python tools/bench_suite.py [--only cfg1,cfg3,...] [--out gpurun_out/suite.json] [--scale 1.0]

Not the driver's headline (that is bench.py); this is the harness behind DESIGN.md's tables and the
numbers copied into profiles/.  Inputs are generated on the GPU (seeded), sizes are BASELINE.json's.
Each record carries device-resident kernel time (CUDA events, L2 flushed or inputs > L2), the
roofline fraction under the byte/flop model of SURVEY.md 8(d), and where it applies the end-to-end
time through the host entry point with pinned host buffers (H2D/D2H inside the timed region).
"""
from __future__ import annotations

import argparse
import json
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

import numpy as np  # noqa: E402
import torch  # noqa: E402

import __graft_entry__ as g  # noqa: E402

PK = json.loads((ROOT / "MEASURED_PEAKS.json").read_text()) if (ROOT / "MEASURED_PEAKS.json").exists() else {
    "hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}
HBM = PK["hbm_gbs"]


def flush_l2(buf):
    buf.zero_()


def time_gpu(fn, iters=5, warm=3, flush=None):
    ts = []
    for i in range(warm + iters):
        if flush is not None:
            flush_l2(flush)
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        if i >= warm:
            ts.append(e0.elapsed_time(e1) * 1e-3)
    return sum(ts) / len(ts), min(ts)


def gen_csr_gpu(m, n, nzr, seed, chunk=1 << 18):
    """exactly nzr sorted columns per row (duplicates possible, as a CSR may hold), U[0,1) values."""
    gen = torch.Generator(device="cuda"); gen.manual_seed(seed)
    idx = torch.empty((m, nzr), dtype=torch.int32, device="cuda")
    for r0 in range(0, m, chunk):
        r1 = min(m, r0 + chunk)
        c = torch.randint(0, n, (r1 - r0, nzr), device="cuda", generator=gen, dtype=torch.int32)
        idx[r0:r1] = torch.sort(c, dim=1).values
    vals = torch.rand((m * nzr,), device="cuda", generator=gen)
    offs = torch.arange(0, (m + 1) * nzr, nzr, dtype=torch.int64, device="cuda")
    return vals, idx.reshape(-1), offs


def pinned_like(t, dtype=None):
    h = torch.empty(t.shape, dtype=dtype or t.dtype, pin_memory=True)
    h.copy_(t if dtype is None else t.to(dtype))
    return h


def spmm_record(ctx, name, m, n, nzr, k, flush, e2e=True, seed=1):
    vals, idx, offs = gen_csr_gpu(m, n, nzr, seed)
    nnz = m * nzr
    B = torch.rand((n, k), device="cuda"); Cm = torch.empty((m, k), device="cuda")
    t, tmin = time_gpu(lambda: ctx.spmm("R", m, n, k, 1.0, vals, idx, offs, B, k, 0.0, Cm, k), flush=flush)
    bytes_gather = nnz * 8 + (m + 1) * 8 + nnz * k * 4 + m * k * 4
    bytes_min = nnz * 8 + (m + 1) * 8 + n * k * 4 + m * k * 4
    # property checks at full size: checksum of checksums in fp64
    colsum = torch.zeros(n, device="cuda", dtype=torch.float64).index_add_(0, idx.long(), vals.double())
    chk = float((Cm.double().sum() - (colsum * B.double().sum(1)).sum()).abs() / Cm.double().sum().abs())
    rec = {"config": name, "kernel": "spmm_csr_rm_vec_kernel", "m": m, "n": n, "nnz": nnz, "k": k, "ms": t * 1e3,
           "ms_min": tmin * 1e3, "gflops": 2.0 * nnz * k / t / 1e9,
           "roofline": {"bound": "hbm", "achieved": bytes_gather / t / 1e9, "peak": HBM, "unit": "GB/s",
                        "frac": bytes_gather / t / 1e9 / HBM, "model": "bytes_gather", "bytes_min_gbs": bytes_min / t / 1e9},
           "checksum_rel_err": chk}
    if e2e:
        a_h, ja_h, ia_h = pinned_like(vals), pinned_like(idx, torch.int64), pinned_like(offs)
        B_h, C_h = pinned_like(B), torch.empty((m, k), dtype=torch.float32, pin_memory=True)
        ref_sum = float(Cm.double().sum())
        del vals, idx, B, Cm
        torch.cuda.empty_cache()
        ts = []
        for i in range(3):
            t0 = time.perf_counter()
            ctx.host_csrmm("N", m, n, k, 1.0, 0.0, a_h, ia_h, ja_h, "R", B_h, C_h)
            ts.append(time.perf_counter() - t0)
        st = ctx.stats()
        te = min(ts[1:])
        streamed = a_h.numel() * 4 + ja_h.numel() * 8 + ia_h.numel() * 8 + B_h.numel() * 4 + C_h.numel() * 4
        rec["e2e"] = {"ms": te * 1e3, "gflops": 2.0 * nnz * k / te / 1e9, "h2d_bytes": st.h2d_bytes,
                      "d2h_bytes": st.d2h_bytes, "streamed_gbs": streamed / te / 1e9,
                      "h2d_gbs": st.h2d_bytes / te / 1e9, "api": "bof_host_csrmm, pinned host buffers",
                      "sum_rel_err": abs(float(C_h.double().sum()) - ref_sum) / abs(ref_sum)}
    return rec


def spmv_csrcsc_record(ctx, m, n, nzr, flush, seed=2):
    out = []
    vals, idx, offs = gen_csr_gpu(m, n, nzr, seed)
    nnz = m * nzr
    x = torch.rand(n, device="cuda"); y = torch.empty(m, device="cuda")
    for trans in "NT":
        xv, yv = (x, y) if trans == "N" else (torch.rand(m, device="cuda"), torch.empty(n, device="cuda"))
        t, tmin = time_gpu(lambda: ctx.spmv(trans, m, n, vals, idx, offs, xv, yv), flush=flush)
        byts = nnz * 8 + (m + 1) * 8 + n * 4 + m * 4
        out.append({"config": f"cfg4 csrgemv '{trans}'", "kernel": f"spmv_csr_{trans.lower()}_kernel", "m": m, "nnz": nnz,
                    "ms": t * 1e3, "gflops": 2.0 * nnz / t / 1e9,
                    "roofline": {"bound": "hbm", "achieved": byts / t / 1e9, "peak": HBM, "unit": "GB/s",
                                 "frac": byts / t / 1e9 / HBM, "model": "nnz*(4+4) + offsets + x + y"}})
    o1 = torch.empty(n + 1, dtype=torch.int64, device="cuda"); i1 = torch.empty(nnz, dtype=torch.int32, device="cuda")
    v1 = torch.empty(nnz, device="cuda")
    ws = ctx.csr2csc_workspace(m, n, nnz)
    t, tmin = time_gpu(lambda: ctx.csr2csc(m, n, nnz, offs, idx, vals, o1, i1, v1, ws=ws), iters=3, warm=1)
    ideal = nnz * 4 + nnz * 8 + nnz * 8 + (m + n + 2) * 8  # SURVEY 8d with I = 4: 20 B/nnz
    ok_hist = bool(torch.equal(torch.bincount(idx.long(), minlength=n), o1[1:] - o1[:-1]))
    # involution: transpose back and compare bit for bit
    o2 = torch.empty(m + 1, dtype=torch.int64, device="cuda"); i2 = torch.empty_like(i1); v2 = torch.empty_like(v1)
    ctx.csr2csc(n, m, nnz, o1, i1, v1, o2, i2, v2, ws=ws)
    same = bool(torch.equal(o2, offs) and torch.equal(v2.view(torch.int32), vals.view(torch.int32)))
    # duplicates inside a row may swap places only if unstable; with a stable sort i2 == idx exactly
    same = same and bool(torch.equal(i2, idx))
    out.append({"config": "cfg4 csrcsc", "kernel": "radix_hist/scan/scatter x passes", "m": m, "n": n, "nnz": nnz,
                "ms": t * 1e3, "roofline": {"bound": "hbm", "achieved": ideal / t / 1e9, "peak": HBM, "unit": "GB/s",
                                            "frac": ideal / t / 1e9 / HBM, "model": "single-pass ideal 20 B/nnz (I=4)"},
                "histogram_matches_offsets": ok_hist, "double_transpose_bit_exact": same})
    del ws, o1, i1, v1, o2, i2, v2
    torch.cuda.empty_cache()
    # e2e csrcsc through the host entry point (int64 indices on the host side)
    a_h, ja_h, ia_h = pinned_like(vals), pinned_like(idx, torch.int64), pinned_like(offs)
    at_h = torch.empty(nnz, dtype=torch.float32, pin_memory=True); jat_h = torch.empty(nnz, dtype=torch.int64, pin_memory=True)
    iat_h = torch.empty(n + 1, dtype=torch.int64, pin_memory=True)
    del vals, idx
    torch.cuda.empty_cache()
    ts = []
    for i in range(2):
        t0 = time.perf_counter()
        ctx.host_csrcsc(m, n, ia_h, ja_h, a_h, iat_h, jat_h, at_h)
        ts.append(time.perf_counter() - t0)
    st = ctx.stats()
    out[-1]["e2e"] = {"ms": ts[-1] * 1e3, "h2d_bytes": st.h2d_bytes, "d2h_bytes": st.d2h_bytes,
                      "pcie_gbs": (st.h2d_bytes + st.d2h_bytes) / ts[-1] / 1e9, "api": "bof_host_csrcsc, pinned"}
    # e2e csrgemv 'N'
    x_h = pinned_like(x); y_h = torch.empty(m, dtype=torch.float32, pin_memory=True)
    ts = []
    for i in range(2):
        t0 = time.perf_counter()
        ctx.host_csrgemv("N", m, n, a_h, ia_h, ja_h, x_h, y_h)
        ts.append(time.perf_counter() - t0)
    st = ctx.stats()
    out[0]["e2e"] = {"ms": ts[-1] * 1e3, "gflops": 2.0 * nnz / ts[-1] / 1e9, "h2d_bytes": st.h2d_bytes,
                     "h2d_gbs": st.h2d_bytes / ts[-1] / 1e9, "api": "bof_host_csrgemv, pinned"}
    return out


def kmeans_record(bof, ctx, P, K, d, iters, seed=5):
    gen = torch.Generator(device="cuda"); gen.manual_seed(seed)
    cent_true = torch.randn((K, d), device="cuda", generator=gen) * 4
    pts = torch.empty((P, d), dtype=torch.float32, pin_memory=True)
    for r0 in range(0, P, 1 << 20):
        r1 = min(P, r0 + (1 << 20))
        lab = torch.randint(0, K, (r1 - r0,), device="cuda", generator=gen)
        pts[r0:r1].copy_(cent_true[lab] + 0.5 * torch.randn((r1 - r0, d), device="cuda", generator=gen))
    cent0 = pts[:K].clone()
    t0 = time.perf_counter()
    km = bof.KMeans(ctx, P, K, d, pts, cent0)
    t_open = time.perf_counter() - t0
    torch.cuda.synchronize()
    stream = torch.cuda.ExternalStream(km.stream())
    step_ms = []
    with torch.cuda.stream(stream):
        for it in range(iters):
            e0, e1, e2 = torch.cuda.Event(True), torch.cuda.Event(True), torch.cuda.Event(True)
            e0.record(stream)
            km.local_step()
            e1.record(stream)
            km.update()
            e2.record(stream)
            stream.synchronize()
            step_ms.append((e0.elapsed_time(e2), e0.elapsed_time(e1)))
    assign = np.zeros(P, np.int64); cent = np.zeros((K, d), np.float32)
    km.get(cent, assign)
    km.close()
    it_ms = sum(s[0] for s in step_ms[1:]) / max(1, len(step_ms) - 1)
    flops = 2.0 * P * K * d
    counts = np.bincount(assign, minlength=K)
    tf32_third = PK["bf16_tflops_sustained"] / 2 / 3
    return {"config": f"cfg5 kmeans {P}x{d}, k={K}, {iters} iterations", "kernel": "gemm3xtf32_kernel<2,ARGMIN> + segment sums",
            "upload_and_prepare_s": t_open, "ms_per_iter": it_ms, "ms_first_iter": step_ms[0][0],
            "distance_tflops": flops / (it_ms * 1e-3) / 1e12,
            "roofline": {"bound": "tensor", "achieved": flops / (it_ms * 1e-3) / 1e12, "peak": tf32_third, "unit": "TFLOP/s",
                         "frac": flops / (it_ms * 1e-3) / 1e12 / tf32_third, "model": "2*P*K*d useful flops per iteration over the whole iteration time"},
            "total_s_20_iters": sum(s[0] for s in step_ms) * 1e-3, "empty_clusters": int((counts == 0).sum()),
            "min_cluster": int(counts.min()), "max_cluster": int(counts.max())}


def gemm_chunk_sweep(bof, sizes=(32768,), chunks=(128, 256, 512, 1024)):
    out = []
    for n in sizes:
        A = torch.rand((n, n), device="cuda"); B = torch.rand((n, n), device="cuda"); Cm = torch.empty((n, n), device="cuda")
        ii = torch.randint(0, n, (256,), device="cuda"); jj = torch.randint(0, n, (256,), device="cuda")
        ref = (A[ii].double() * B[:, jj].t().double()).sum(1)
        for ch in chunks:
            with bof.Context(device=0, gemm_k_chunk=ch) as c2:
                ws = c2.sgemm_workspace(n, n, n)
                t, tmin = time_gpu(lambda: c2.sgemm("R", "N", "N", n, n, n, 1.0, A, 0, B, 0, 0.0, Cm, 0, ws=ws), iters=3, warm=1)
                kern_ms = c2.stats().kernel_ms
                err = float(((Cm[ii, jj].double() - ref).norm() / ref.norm()))
                out.append({"config": f"gemm {n}^3 k_chunk={ch}", "ms": t * 1e3, "kernel_ms": kern_ms,
                            "tflops": 2.0 * n ** 3 / t / 1e12, "sampled_rel_fro_err": err})
                del ws
        del A, B, Cm
        torch.cuda.empty_cache()
    return out


def gemm_split_sweep(bof, n=32768):
    """Pure 3xTF32 vs hybrid (TF32 + 2 x BF16 cross terms) on the headline shape, with sampled accuracy."""
    out = []
    A = torch.rand((n, n), device="cuda"); B = torch.rand((n, n), device="cuda"); Cm = torch.empty((n, n), device="cuda")
    ii = torch.randint(0, n, (256,), device="cuda"); jj = torch.randint(0, n, (256,), device="cuda")
    ref = (A[ii].double() * B[:, jj].t().double()).sum(1)
    for split in (1, 2):
        with bof.Context(device=0, gemm_split=split) as c2:
            ws = c2.sgemm_workspace(n, n, n)
            t, tmin = time_gpu(lambda: c2.sgemm("R", "N", "N", n, n, n, 1.0, A, 0, B, 0, 0.0, Cm, 0, ws=ws), iters=3, warm=2)
            err = float(((Cm[ii, jj].double() - ref).norm() / ref.norm()))
            out.append({"config": f"gemm {n}^3 gemm_split={split}", "ms": t * 1e3, "kernel_ms": c2.stats().kernel_ms,
                        "tflops": 2.0 * n ** 3 / t / 1e12, "sampled_rel_fro_err": err})
            del ws
    return out


def gemm_sync_sweep(bof, n=32768, syncs=(-1, 16, 64, 256)):
    """A/B of the wave lock-step (bof_config.gemm_wave_sync) on the headline shape."""
    out = []
    A = torch.rand((n, n), device="cuda"); B = torch.rand((n, n), device="cuda"); Cm = torch.empty((n, n), device="cuda")
    for sy in syncs:
        with bof.Context(device=0, gemm_wave_sync=sy) as c2:
            ws = c2.sgemm_workspace(n, n, n)
            t, tmin = time_gpu(lambda: c2.sgemm("R", "N", "N", n, n, n, 1.0, A, 0, B, 0, 0.0, Cm, 0, ws=ws), iters=3, warm=2)
            out.append({"config": f"gemm {n}^3 wave_sync={sy}", "ms": t * 1e3, "ms_min": tmin * 1e3, "kernel_ms": c2.stats().kernel_ms,
                        "tflops": 2.0 * n ** 3 / t / 1e12})
            del ws
    return out


def pcie_record():
    """Out-of-core roofline denominators: pinned cudaMemcpyAsync H2D, D2H, and both at once (1 GiB each)."""
    n = 1 << 28
    h_in = torch.empty(n, dtype=torch.float32, pin_memory=True); h_out = torch.empty(n, dtype=torch.float32, pin_memory=True)
    d_in = torch.empty(n, dtype=torch.float32, device="cuda"); d_out = torch.rand(n, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    res = {}
    for name in ("h2d", "d2h", "both"):
        best = 1e9
        for _ in range(4):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            if name in ("h2d", "both"):
                with torch.cuda.stream(s1):
                    d_in.copy_(h_in, non_blocking=True)
            if name in ("d2h", "both"):
                with torch.cuda.stream(s2):
                    h_out.copy_(d_out, non_blocking=True)
            torch.cuda.synchronize()
            best = min(best, time.perf_counter() - t0)
        res[name + "_gbs"] = (2 if name == "both" else 1) * n * 4 / best / 1e9
    return {"config": "PCIe pinned copy bandwidth (1 GiB transfers)", **res}


def tf32_cublas_peak():
    """Denominator only: cuBLAS TF32 GEMM (library call, not on the product path)."""
    torch.backends.cuda.matmul.allow_tf32 = True
    n = 8192
    a = torch.rand((n, n), device="cuda"); b = torch.rand((n, n), device="cuda")
    for _ in range(3):
        a @ b
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(10):
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record(); a @ b; e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    t_end = time.time() + 4.0
    reps = 0
    e0.record()
    while time.time() < t_end:
        for _ in range(20):
            a @ b
        reps += 20
        torch.cuda.synchronize()
    e1.record(); torch.cuda.synchronize()
    sustained = 2.0 * n ** 3 * reps / (e0.elapsed_time(e1) * 1e-3) / 1e12
    torch.backends.cuda.matmul.allow_tf32 = False
    return {"config": "cuBLAS TF32 8192^3 (denominator)", "tf32_tflops_burst": 2.0 * n ** 3 / (best * 1e-3) / 1e12,
            "tf32_tflops_sustained": sustained}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="tf32,cfg1,chunks,cfg3,cfg4,cfg5")
    ap.add_argument("--out", default=str(ROOT / "gpurun_out" / "suite.json"))
    ap.add_argument("--scale", type=float, default=1.0, help="shrink the big configs (debug)")
    args = ap.parse_args()
    only = set(args.only.split(","))
    bof = g.load_package()
    ctx = bof.Context(device=0)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    recs = []

    def add(r):
        rs = r if isinstance(r, list) else [r]
        for x in rs:
            print(json.dumps(x), flush=True)
        recs.extend(rs)
        Path(args.out).parent.mkdir(exist_ok=True)
        Path(args.out).write_text(json.dumps(recs, indent=1))

    if "tf32" in only:
        add(tf32_cublas_peak())
    if "pcie" in only:
        add(pcie_record())
    if "cfg1" in only:
        add(spmm_record(ctx, "cfg1 csrmm 262144^2, 64 nnz/row, k=128", 262144, 262144, 64, 128, flush))
    if "chunks" in only:
        add(gemm_chunk_sweep(bof))
    if "sync" in only:
        add(gemm_sync_sweep(bof))
    if "split" in only:
        add(gemm_split_sweep(bof))
    big = int((1 << 23) * args.scale)
    if "cfg3" in only:
        add(spmm_record(ctx, f"cfg3 csrmm {big}^2, 100 nnz/row, k=256", big, big, 100, 256, None, seed=3))
    if "cfg4" in only:
        add(spmv_csrcsc_record(ctx, big, big, 100, flush))
    if "cfg5" in only:
        add(kmeans_record(bof, ctx, int(10_000_000 * args.scale), 1024, 256, 20))
    ctx.close()


if __name__ == "__main__":
    main()
