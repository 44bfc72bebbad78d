"""``torch.ops.xpore_b200.fit`` - the diffmod hot path as a PyTorch custom operator.

A thin layer over ``xpb200_fit_device`` (C ABI, ``include/xpore_b200.h``): CUDA tensors in, CUDA tensors out, enqueued on
the current torch stream without any synchronisation.  It replaces what ``execute()`` does per gene between
``io.load_data`` and ``io.generate_result_table`` (``xpore/scripts/diffmod.py:20-58``: priors, prefilter gate,
``GMM.fit``, and the statistics of ``io.py:255-367``) for a CSR batch of positions.

    import xpore_b200.torch_op  # registers the operator
    status, n_iter, pval, fit, stats, n_fitted = torch.ops.xpore_b200.fit(
        y, read_off, grp_len, kmer_mean, kmer_std, grp_cond, n_conditions, max_reads,
        max_iters, stopping_criteria, prefilter, prefilter_threshold, dirichlet_conc)

``y``: float64 reads, int32 codes (reads x 100) or int16 carrying uint16 codes (reads x 100) - the three read formats of the
C ABI; the other arguments are the arrays of ``xpb200_batch`` and the fields of ``xpb200_method``.  There is no CPU
implementation: CPU tensors raise.
"""
from __future__ import annotations

import ctypes as C
import math

import torch

from . import _lib
from .fit import fit_width, stats_width

_Y_FORMAT = {torch.float64: (0, 1.0), torch.int32: (1, 100.0), torch.int16: (2, 100.0)}


@torch.library.custom_op("xpore_b200::fit", mutates_args=(), device_types="cuda")
def fit(y: torch.Tensor, read_off: torch.Tensor, grp_len: torch.Tensor, kmer_mean: torch.Tensor, kmer_std: torch.Tensor,
        grp_cond: torch.Tensor, n_conditions: int, max_reads: int, max_iters: int, stopping_criteria: float,
        prefilter: int, prefilter_threshold: float, dirichlet_conc: float
        ) -> tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
    if y.dtype not in _Y_FORMAT:
        raise TypeError("y must be float64, int32 (reads x 100) or int16 (uint16 codes, reads x 100)")
    yfmt, yscale = _Y_FORMAT[y.dtype]
    dev = y.device
    P, G = int(kmer_mean.numel()), int(grp_cond.numel())
    assert read_off.dtype == torch.int64 and read_off.numel() == P + 1
    assert grp_len.dtype == torch.int32 and grp_len.numel() == P * G and grp_cond.dtype == torch.int32
    assert kmer_mean.dtype == torch.float64 and kmer_std.dtype == torch.float64
    for t in (y, read_off, grp_len, kmer_mean, kmer_std, grp_cond):
        assert t.is_cuda and t.device == dev and t.is_contiguous()
    L = _lib.lib()
    status = torch.zeros(P, dtype=torch.int32, device=dev)
    n_iter = torch.zeros(P, dtype=torch.int32, device=dev)
    pval = torch.full((P,), math.nan, dtype=torch.float64, device=dev)
    fit_rec = torch.full((P, fit_width(G)), math.nan, dtype=torch.float64, device=dev)
    stats = torch.full((P, stats_width(G, int(n_conditions))), math.nan, dtype=torch.float64, device=dev)
    n_fitted = torch.zeros(1, dtype=torch.int32, device=dev)
    ws = torch.empty(int(L.xpb200_workspace_bytes(P)), dtype=torch.uint8, device=dev)
    n_reads = int(y.numel())  # the format's trailing pad (an even number of doubles / 8 codes) is never addressed past read_off[P]
    b = _lib.XpBatch(P, G, int(n_conditions), int(max_reads), n_reads, y.data_ptr(), read_off.data_ptr(), grp_len.data_ptr(),
                     kmer_mean.data_ptr(), kmer_std.data_ptr(), grp_cond.data_ptr(), yfmt, 0, yscale)
    m = _lib.XpMethod(int(max_iters), int(prefilter), float(stopping_criteria), float(prefilter_threshold),
                      float(dirichlet_conc))
    r = _lib.XpResult(status.data_ptr(), n_iter.data_ptr(), pval.data_ptr(), fit_rec.data_ptr(), stats.data_ptr(), None)
    with torch.cuda.device(dev):
        stream = torch.cuda.current_stream(dev).cuda_stream
        _lib.check(L.xpb200_fit_device(C.byref(b), C.byref(m), C.byref(r), ws.data_ptr(), ws.numel(), n_fitted.data_ptr(),
                                       C.c_void_p(stream)))
    ws.record_stream(torch.cuda.current_stream(dev))  # scratch of the enqueued kernels: not reused before they ran
    return status, n_iter, pval, fit_rec, stats, n_fitted


@fit.register_fake
def _(y, read_off, grp_len, kmer_mean, kmer_std, grp_cond, n_conditions, max_reads, max_iters, stopping_criteria,
      prefilter, prefilter_threshold, dirichlet_conc):
    P, G = kmer_mean.numel(), grp_cond.numel()
    i32 = dict(dtype=torch.int32, device=y.device)
    f64 = dict(dtype=torch.float64, device=y.device)
    return (torch.empty(P, **i32), torch.empty(P, **i32), torch.empty(P, **f64), torch.empty((P, fit_width(G)), **f64),
            torch.empty((P, stats_width(G, int(n_conditions))), **f64), torch.empty(1, **i32))
