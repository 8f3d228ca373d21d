"""Multi-GPU host logic: one process per GPU, output row blocks sharded with no data-path collective;
the k-means step is the only place with an exchange (allreduce of centroid sums and counts).

Reference context: BLAS-on-Flash is single-process; the shard rule reuses the idea of
its nnz-budgeted row blocks (include/blas_utils.h:72-97) at the granularity of a rank.
"""
from __future__ import annotations

import numpy as np


def row_shard(n_rows: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous rows [r0, r1) of rank `rank`; sizes differ by at most one row."""
    base, rem = divmod(n_rows, world)
    r0 = rank * base + min(rank, rem)
    return r0, r0 + base + (1 if rank < rem else 0)


def nnz_balanced_shard(ia, world: int, rank: int) -> tuple[int, int]:
    """Contiguous rows [r0, r1) holding ~nnz/world nonzeros each (CSR offsets `ia`, m+1 entries).
    Cut points are the first rows whose offset reaches g * nnz / world, so the shards tile [0, m)."""
    ia = np.asarray(ia)
    m = ia.shape[0] - 1
    nnz = int(ia[m] - ia[0])
    targets = ia[0] + (np.arange(world + 1, dtype=np.int64) * nnz) // world
    cuts = np.searchsorted(ia, targets, side="left").astype(np.int64)
    cuts[0], cuts[-1] = 0, m
    cuts = np.maximum.accumulate(np.minimum(cuts, m))
    return int(cuts[rank]), int(cuts[rank + 1])


class DeviceView:
    """Zero-copy torch view of a raw device buffer returned by the C ABI (``__cuda_array_interface__``)."""

    def __init__(self, ptr: int, count: int, typestr: str = "<f4"):
        self.__cuda_array_interface__ = {"shape": (count,), "typestr": typestr, "data": (ptr, False), "version": 2}


def as_tensor(ptr: int, count: int, device: int):
    import torch

    return torch.as_tensor(DeviceView(ptr, count), device=f"cuda:{device}")


def init_comm(ctx, group=None):
    """Give `ctx` a communicator rank of its own (bof_comm_init): rank 0 creates the NCCL id, torch.distributed
    carries the 128 bytes, the library then runs its collectives (panel broadcasts of the replicated operand, the
    k-means allreduce) on its own streams -- no host synchronisation between the exchange and the kernels."""
    import torch
    import torch.distributed as dist

    from . import comm_unique_id

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return False
    if ctx.comm_world() > 1:
        return True
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    raw = comm_unique_id() if rank == 0 else bytes(128)
    t = torch.tensor(list(raw), dtype=torch.uint8)
    if dist.get_backend(group) == "nccl":
        t = t.cuda()
    dist.broadcast(t, src=0, group=group)
    ctx.comm_init(world, rank, bytes(t.cpu().tolist()))
    return True


def lloyd(km, iters: int, group=None):
    """`iters` Lloyd iterations on this rank's resident shard; with a process group the partial
    [K*dim sums | K counts] buffer is summed in place over NVLink by NCCL on the library's stream."""
    import torch
    import torch.distributed as dist

    if km.ctx.comm_world() > 1:   # the context has its own communicator: the library allreduces on its stream
        km.lloyd(iters)
        return
    use_dist = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
    stream = torch.cuda.ExternalStream(km.stream(), device=km.ctx.device) if use_dist else None
    for _ in range(iters):
        ptr, count = km.local_step()
        if use_dist:
            with torch.cuda.stream(stream):
                dist.all_reduce(as_tensor(ptr, count, km.ctx.device), op=dist.ReduceOp.SUM, group=group)
        km.update()


def allgather_dense(host_full, out_dev, group=None):
    """Assemble a replicated row-major dense operand in every rank's HBM with 1/G of the PCIe traffic:
    rank r uploads rows row_shard(r) of `host_full` (a host tensor every rank can read, e.g. the mmap of
    the same file) into its slice of `out_dev`, then the slices are exchanged over NVLink.
    Returns the bytes this rank moved over PCIe."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    rows = host_full.shape[0]
    r0, r1 = row_shard(rows, world, rank)
    out_dev[r0:r1].copy_(host_full[r0:r1], non_blocking=True)
    if world > 1:
        if rows % world == 0:
            dist.all_gather_into_tensor(out_dev, out_dev[r0:r1], group=group)
        else:  # uneven shards: one broadcast per owner
            for src in range(world):
                s0, s1 = row_shard(rows, world, src)
                dist.broadcast(out_dev[s0:s1], src=src, group=group)
    torch.cuda.current_stream().synchronize()  # the library's streams do not know about this one
    return (r1 - r0) * host_full[0].numel() * host_full.element_size()


def sum_partials_fixed_order(y_local, group=None):
    """csrgemv 'T' over row shards: every rank holds a full-length partial y = A_shard^T x_shard;
    the result is their sum taken on the host in rank order, so it is bitwise reproducible and identical on
    every rank (the reference adds task results in completion order under a mutex,
    include/tasks/csrgemv_task.h:170-176).  The partials travel with all_gather (exact copies), not a reduction."""
    import torch
    import torch.distributed as dist

    y_local = np.ascontiguousarray(y_local, dtype=np.float32)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return y_local.copy()
    world = dist.get_world_size(group)
    t = torch.from_numpy(y_local)
    if dist.get_backend(group) == "nccl":
        t = t.cuda()
    parts = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(parts, t, group=group)
    out = parts[0].cpu().numpy().copy()
    for p in parts[1:]:
        out += p.cpu().numpy()
    return out


def bind_to_gpu_numa(device_index: int) -> int:
    """Pin the calling process to the CPUs NVML reports as local to CUDA device `device_index`, so that the pinned
    host buffers it allocates afterwards sit on the NUMA node next to that GPU's PCIe root (one process per GPU:
    otherwise every rank's DMA may cross the socket interconnect).  Returns the number of CPUs in the new
    affinity mask, 0 if NVML or the affinity call is unavailable (nothing changed)."""
    import os

    try:
        import pynvml
        import torch

        pynvml.nvmlInit()
        p = torch.cuda.get_device_properties(device_index)
        bus = "%08X:%02X:%02X.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
        try:
            h = pynvml.nvmlDeviceGetHandleByPciBusId(bus)
        except Exception:
            h = pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode())
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return 0
        os.sched_setaffinity(0, cpus)
        return len(cpus)
    except Exception:
        return 0
