"""Multi-GPU plumbing: one process per GPU, positions sharded, one NCCL gather of result records.

The reference parallelises over genes with OS processes and appends to one table under a file
lock (``xpore/scripts/diffmod.py:109-125,68-69``; ``helper.py:96-116``).  Positions are
independent, so here each rank fits its own shard with no data-path collective; the only
exchange is the final gather of the fixed-width records of fitted positions to rank 0 (the
single writer).  Works with ``nccl`` (GPU) and ``gloo`` (CPU tests).
"""
from __future__ import annotations

import numpy as np


def balanced_partition(cost: np.ndarray, world: int) -> list:
    """Contiguous ranges of positions with ~equal total cost (cost = reads per position — the
    ``data.index`` byte span is the same proxy at gene granularity).  Returns ``world`` index
    arrays that cover ``range(len(cost))`` in order (contiguity keeps genes together so the
    per-gene log lines of ``--resume`` stay meaningful)."""
    cost = np.asarray(cost, dtype=np.float64)
    n = len(cost)
    if n == 0:
        return [np.zeros(0, np.int64) for _ in range(world)]
    csum = np.cumsum(cost)
    targets = csum[-1] * (np.arange(1, world) / world)
    cuts = np.searchsorted(csum, targets, side="left") + 1
    cuts = np.minimum(np.maximum.accumulate(np.concatenate([[0], cuts, [n]])), n)
    return [np.arange(cuts[r], cuts[r + 1], dtype=np.int64) for r in range(world)]


def bind_to_gpu_cpus(cuda_index: int) -> int:
    """Pin this process to the CPUs NVML reports as local to the GPU (its NUMA node), so that pinned host
    buffers allocated afterwards sit next to the GPU's PCIe root.  With one process per GPU and all ranks
    copying at once, buffers on the far socket halve the host-to-device rate.  Best effort: returns the
    number of CPUs bound to, 0 when NVML, the device or an allowed local CPU is missing (nothing changed)."""
    import os
    try:
        import pynvml
        import torch

        pr = torch.cuda.get_device_properties(cuda_index)
        bus = "%08x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode())
        words = pynvml.nvmlDeviceGetCpuAffinity(h, ((os.cpu_count() or 64) + 63) // 64)
        allowed = os.sched_getaffinity(0)
        cpus = {64 * i + b for i, w in enumerate(words) for b in range(64) if (int(w) >> b) & 1} & allowed
        if not cpus or cpus == allowed:
            return 0
        os.sched_setaffinity(0, cpus)
        return len(cpus)
    except Exception:  # no NVML / no permission: leave the affinity alone
        return 0


def pack_records(status, n_iter, pval, fit, stats, index_offset: int = 0):
    """Compact float64 record matrix of the selected positions (torch tensors in, any device):
    columns = [global position index, status, n_iter, pval, fit..., stats...]."""
    import torch

    sel = torch.nonzero(status.to(torch.int64) & 1, as_tuple=False).squeeze(1)
    cols = [(sel + index_offset).to(torch.float64)[:, None],
            status[sel].to(torch.int64).bitwise_and(0xFFFFFFFF).to(torch.float64)[:, None],
            n_iter[sel].to(torch.float64)[:, None], pval[sel][:, None], fit[sel], stats[sel]]
    return torch.cat(cols, dim=1)


def unpack_records(rec: np.ndarray, fit_w: int, stats_w: int):
    idx = rec[:, 0].astype(np.int64)
    status = rec[:, 1].astype(np.int64).astype(np.uint32)
    n_iter = rec[:, 2].astype(np.int32)
    pval = rec[:, 3]
    fit = rec[:, 4:4 + fit_w]
    stats = rec[:, 4 + fit_w:4 + fit_w + stats_w]
    return idx, status, n_iter, pval, fit, stats


def gather_records(rec, rank: int, world: int, dst: int = 0):
    """Gather per-rank record matrices [n_r, W] on ``dst`` (rows ordered by rank).  One tiny
    all_gather of the counts, then one padded gather — the only collective on the path."""
    import torch
    import torch.distributed as dist

    if world == 1:
        return rec
    n = torch.tensor([rec.shape[0]], dtype=torch.int64, device=rec.device)
    counts = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(counts, n)
    counts = [int(c.item()) for c in counts]
    m = max(max(counts), 1)
    pad = torch.zeros((m, rec.shape[1]), dtype=rec.dtype, device=rec.device)
    pad[:rec.shape[0]] = rec
    if rank == dst:
        bufs = [torch.empty_like(pad) for _ in range(world)]
        dist.gather(pad, bufs, dst=dst)
        return torch.cat([b[:c] for b, c in zip(bufs, counts)], dim=0)
    dist.gather(pad, None, dst=dst)
    return None
