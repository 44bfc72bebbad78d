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
    """Gather per-rank record matrices [n_r, W] on ``dst`` (rows ordered by rank): one all-gather of the counts
    (a single host read), then one padded ``gather`` into a buffer allocated once per call - the only collective on
    the path.  ``gather_records.last_counts`` holds the per-rank row counts of the last call."""
    import torch
    import torch.distributed as dist

    if world == 1:
        gather_records.last_counts = [int(rec.shape[0])]
        return rec
    n = torch.tensor([rec.shape[0]], dtype=torch.int64, device=rec.device)
    counts = torch.empty(world, dtype=torch.int64, device=rec.device)
    dist.all_gather_into_tensor(counts, n)
    counts = counts.tolist()
    gather_records.last_counts = counts
    m = max(max(counts), 1)
    if rec.shape[0] == m:
        pad = rec.contiguous()
    else:
        pad = torch.zeros((m, rec.shape[1]), dtype=rec.dtype, device=rec.device)
        pad[:rec.shape[0]] = rec
    if rank == dst:
        buf = torch.empty((world, m, rec.shape[1]), dtype=rec.dtype, device=rec.device)
        dist.gather(pad, list(buf.unbind(0)), dst=dst)
        return torch.cat([buf[r, :c] for r, c in enumerate(counts)], dim=0)
    dist.gather(pad, None, dst=dst)
    return None


gather_records.last_counts = None


class RowGatherer:
    """The per-step result gather of a multi-GPU run without any host synchronisation on the sending ranks: every
    rank holds its rows in a fixed-capacity device buffer [cap, W] with the counts {rows, fitted} beside it (what
    ``xpb200_fit_host_rows_async`` / ``xpb200_pack_rows`` leave behind); ``gather`` ships both to ``dst`` with two
    NCCL gathers into buffers allocated once; ``fetch`` (dst only) reads the counts - its single host
    synchronisation - and copies exactly the rows to page-locked host memory."""

    def __init__(self, cap: int, width: int, rank: int, world: int, device, dst: int = 0):
        import torch
        self.torch, self.rank, self.world, self.dst, self.cap, self.W = torch, rank, world, dst, int(cap), int(width)
        self.rows = torch.zeros((self.cap, self.W), dtype=torch.float64, device=device)
        self.counts = torch.zeros(2, dtype=torch.int32, device=device)
        if rank == dst:
            self.all_rows = torch.empty((world, self.cap, self.W), dtype=torch.float64, device=device)
            self.all_counts = torch.zeros((world, 2), dtype=torch.int32, device=device)
            self.host = torch.empty((world * self.cap, self.W), dtype=torch.float64).pin_memory() \
                if torch.cuda.is_available() else torch.empty((world * self.cap, self.W), dtype=torch.float64)
            self.host_counts = torch.zeros((world, 2), dtype=torch.int32)
            if torch.cuda.is_available():
                self.host_counts = self.host_counts.pin_memory()

    def gather(self):
        import torch.distributed as dist
        if self.world == 1:
            return
        if self.rank == self.dst:
            dist.gather(self.counts, list(self.all_counts.unbind(0)), dst=self.dst)
            dist.gather(self.rows, list(self.all_rows.unbind(0)), dst=self.dst)
        else:
            dist.gather(self.counts, None, dst=self.dst)
            dist.gather(self.rows, None, dst=self.dst)

    def fetch(self):
        """dst only: (host rows [n, W] view, per-rank counts, bytes copied device -> host)."""
        torch = self.torch
        if self.world == 1:
            self.host_counts[0].copy_(self.counts)
            if self.rows.is_cuda:
                torch.cuda.current_stream().synchronize()
            counts = self.host_counts[:1].tolist()
            src = [self.rows]
        else:
            self.host_counts.copy_(self.all_counts, non_blocking=True)
            if self.all_counts.is_cuda:
                torch.cuda.current_stream().synchronize()
            counts = self.host_counts.tolist()
            src = list(self.all_rows.unbind(0))
        o = 0
        for r, (n, _) in enumerate(counts):
            if n > self.cap:
                raise RuntimeError("rank %d produced %d rows, capacity %d" % (r, n, self.cap))
            self.host[o:o + n].copy_(src[r][:n], non_blocking=True)
            o += n
        if self.rows.is_cuda:
            torch.cuda.current_stream().synchronize()
        return self.host[:o], counts, o * self.W * 8 + 8 * len(counts)
