"""Batched replacement of the reference's per-position fit loop.

Reference seam (``xpore/scripts/diffmod.py:21-58``): for every position of a gene build priors,
optionally ``StatsTest(data).fit('t-test')``, then ``GMM(method, data, priors, kmer_signal).fit()``
and later ``io.generate_result_table``.  Here a whole :class:`~xpore_b200.batch.Batch` goes
through one C-ABI call; results come back as fixed-width records (``include/xpore_b200.h``).

Two entry points:

* :func:`fit_batch`   — host arrays in, host arrays out (``xpb200_fit_host``): the drop-in call;
* :func:`fit_device`  — tensors already resident in HBM, enqueued on the current torch stream
  without synchronising (``xpb200_fit_device``); used by the benchmark and the multi-GPU driver.

PyTorch only provides device memory and streams here.  No CPU fallback exists.
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass

import numpy as np

from . import _lib
from .batch import Batch

SELECTED, CONVERGED, ELBO_DECREASED, HIT_MAX, KEEP_ROW, MOD_IS_1, HIGHER, TTEST_VALID = 1, 2, 4, 8, 16, 32, 64, 128
NOT_FITTED, NEAR_THRESHOLD = 256, 512
FIT_MU, FIT_LAM, FIT_ALPHA, FIT_BETA, FIT_ELBO, FIT_N = 0, 2, 4, 6, 8, 9
ST_MU, ST_S2, ST_CONF, ST_POV, ST_CDF, ST_RATE = 0, 2, 4, 6, 7, 11


@dataclass
class Method:
    """The ``method`` section of the YAML as the path reads it (``configurator.py:46-60``)."""
    max_iters: int = 500
    stopping_criteria: float = 1e-5
    pooling: bool = False
    prefilter_method: str | None = None
    prefilter_threshold: float = float("nan")
    dirichlet_conc: float = 0.001

    @classmethod
    def from_dict(cls, method: dict) -> "Method":
        pre = method.get("prefiltering", False)
        return cls(max_iters=int(method.get("max_iters", 500)),
                   stopping_criteria=float(method.get("stopping_criteria", 1e-5)),
                   pooling=bool(method.get("pooling", False)),
                   prefilter_method=(pre["method"] if pre else None),
                   prefilter_threshold=(float(pre["threshold"]) if pre else float("nan")))

    def c_struct(self) -> _lib.XpMethod:
        # prefilter: 0 off; 1 t-test; 2 = configured with another method => p = NaN => fit all
        # (statstest.py:11-18 only knows 't-test')
        code = 0 if self.prefilter_method is None else (1 if self.prefilter_method == "t-test" else 2)
        return _lib.XpMethod(self.max_iters, code, self.stopping_criteria, self.prefilter_threshold,
                             self.dirichlet_conc)


def fit_width(G: int) -> int:
    return FIT_N + 2 * G


def stats_width(G: int, C_: int) -> int:
    return ST_RATE + 2 * G + 3 * (C_ * (C_ - 1) // 2) + (3 * C_ if C_ > 2 else 0)


@dataclass
class FitResult:
    """Host-side view of the result records of one batch."""
    n_groups: int
    n_conditions: int
    status: np.ndarray
    n_iter: np.ndarray
    pval: np.ndarray
    fit: np.ndarray
    stats: np.ndarray
    n_fitted: int

    @property
    def selected(self):
        return (self.status & SELECTED) != 0

    @property
    def keep_row(self):
        return (self.status & (SELECTED | KEEP_ROW)) == (SELECTED | KEEP_ROW)

    @property
    def mu(self):
        return self.fit[:, FIT_MU:FIT_MU + 2]

    @property
    def lam(self):
        return self.fit[:, FIT_LAM:FIT_LAM + 2]

    @property
    def alpha(self):
        return self.fit[:, FIT_ALPHA:FIT_ALPHA + 2]

    @property
    def beta(self):
        return self.fit[:, FIT_BETA:FIT_BETA + 2]

    @property
    def elbo(self):
        return self.fit[:, FIT_ELBO]

    @property
    def N(self):
        return self.fit[:, FIT_N:].reshape(-1, self.n_groups, 2)


def _max_reads(batch: Batch) -> int:
    if getattr(batch, "max_reads", None) is not None:
        return int(batch.max_reads)
    return int(np.diff(batch.read_off).max()) if batch.n_positions else 0


def _padded_y(y: np.ndarray) -> np.ndarray:
    """The kernels copy 16-byte aligned spans: give y an even number of elements."""
    if y.shape[0] % 2 == 0 and y.shape[0] > 0:
        return y
    out = np.zeros(y.shape[0] + 1 + (y.shape[0] + 1) % 2, dtype=np.float64)
    out[:y.shape[0]] = y
    return out


def _y_buffer(batch: Batch):
    """(array, format, scale) of the read buffer the library should see."""
    if batch.y_enc is not None and batch.y_format != 0:
        return np.ascontiguousarray(batch.y_enc), int(batch.y_format), float(batch.y_scale)
    return _padded_y(np.ascontiguousarray(batch.y, dtype=np.float64)), 0, 1.0


def fit_batch(batch: Batch, method: Method, device: int = 0) -> FitResult:
    """Fit every position of ``batch`` on GPU ``device`` through ``xpb200_fit_host``."""
    L = _lib.lib()
    P, G, Cn = batch.n_positions, batch.layout.n_groups, batch.layout.n_conditions
    y, yfmt, yscale = _y_buffer(batch)
    read_off = np.ascontiguousarray(batch.read_off, dtype=np.int64)
    grp_len = np.ascontiguousarray(batch.grp_len, dtype=np.int32)
    km = np.ascontiguousarray(batch.kmer_mean, dtype=np.float64)
    ks = np.ascontiguousarray(batch.kmer_std, dtype=np.float64)
    gc = np.ascontiguousarray(batch.layout.grp_cond, dtype=np.int32)
    res = FitResult(G, Cn, np.zeros(P, np.uint32), np.zeros(P, np.int32), np.full(P, np.nan),
                    np.full((P, fit_width(G)), np.nan), np.full((P, stats_width(G, Cn)), np.nan), 0)
    ptr = lambda a: a.ctypes.data_as(C.c_void_p)
    b = _lib.XpBatch(P, G, Cn, _max_reads(batch), batch.n_reads, ptr(y), ptr(read_off), ptr(grp_len), ptr(km),
                     ptr(ks), ptr(gc), yfmt, 0, yscale)
    r = _lib.XpResult(ptr(res.status), ptr(res.n_iter), ptr(res.pval), ptr(res.fit), ptr(res.stats))
    m = method.c_struct()
    nf = C.c_uint32(0)
    _lib.check(L.xpb200_fit_host(int(device), C.byref(b), C.byref(m), C.byref(r), C.byref(nf)))
    res.n_fitted = int(nf.value)
    return res


def fit_rows(batch: Batch, method: Method, device: int = 0, capacity_rows: int | None = None, staging=None,
             max_reads: int | None = None):
    """Rows of ``diffmod.table`` only (``xpb200_fit_host_rows``): ``(records, n_fitted)`` with ``records`` the
    ``[n_rows, 4+FW+SW]`` matrix ``[position, status, n_iter, pval, fit, stats]`` of the positions that were
    fitted and pass the reference's row filter, ascending by position (``dist.unpack_records`` splits it).
    ``staging`` (:class:`xpore_b200.ingest.Staging`): take the record buffer from its page-locked memory - the
    returned matrix is then a view that the next call with the same staging overwrites.  ``max_reads`` overrides the
    batch's own deepest position as the caller's claim (tests: the library checks it against the offsets)."""
    L = _lib.lib()
    P, G, Cn = batch.n_positions, batch.layout.n_groups, batch.layout.n_conditions
    y, yfmt, yscale = _y_buffer(batch)
    read_off = np.ascontiguousarray(batch.read_off, dtype=np.int64)
    grp_len = np.ascontiguousarray(batch.grp_len, dtype=np.int32)
    km = np.ascontiguousarray(batch.kmer_mean, dtype=np.float64)
    ks = np.ascontiguousarray(batch.kmer_std, dtype=np.float64)
    gc = np.ascontiguousarray(batch.layout.grp_cond, dtype=np.int32)
    W = 4 + fit_width(G) + stats_width(G, Cn)
    cap = P if capacity_rows is None else int(capacity_rows)
    rows = staging.array("rows", np.float64, (max(cap, 1), W)) if staging is not None else \
        np.empty((max(cap, 1), W), dtype=np.float64)
    ptr = lambda a: a.ctypes.data_as(C.c_void_p)
    b = _lib.XpBatch(P, G, Cn, _max_reads(batch) if max_reads is None else int(max_reads), batch.n_reads, ptr(y),
                     ptr(read_off), ptr(grp_len), ptr(km), ptr(ks), ptr(gc), yfmt, 0, yscale)
    m = method.c_struct()
    nr, nf = C.c_int64(0), C.c_uint32(0)
    _lib.check(L.xpb200_fit_host_rows(int(device), C.byref(b), C.byref(m), ptr(rows), cap, C.byref(nr), C.byref(nf)))
    return rows[:int(nr.value)], int(nf.value)


class DeviceBatch:
    """A batch resident in HBM (torch tensors own the memory)."""

    def __init__(self, batch: Batch, device="cuda:0", pin: bool = False):
        import torch
        self.torch = torch
        self.device = torch.device(device)
        self.P, self.G, self.C = batch.n_positions, batch.layout.n_groups, batch.layout.n_conditions
        self.R = batch.n_reads
        self.max_reads = _max_reads(batch)
        ybuf, self.y_format, self.y_scale = _y_buffer(batch)
        if ybuf.dtype == np.uint16:  # torch has no arithmetic on uint16; it is only a byte container here
            ybuf = ybuf.view(np.int16)
        self.host = {
            "y": torch.from_numpy(ybuf),
            "read_off": torch.from_numpy(np.ascontiguousarray(batch.read_off, dtype=np.int64)),
            "grp_len": torch.from_numpy(np.ascontiguousarray(batch.grp_len, dtype=np.int32).reshape(-1)),
            "kmer_mean": torch.from_numpy(np.ascontiguousarray(batch.kmer_mean, dtype=np.float64)),
            "kmer_std": torch.from_numpy(np.ascontiguousarray(batch.kmer_std, dtype=np.float64)),
            "grp_cond": torch.from_numpy(np.ascontiguousarray(batch.layout.grp_cond, dtype=np.int32)),
        }
        if pin:
            self.host = {k: v.pin_memory() for k, v in self.host.items()}
        self.dev = {k: torch.empty_like(v, device=self.device) for k, v in self.host.items()}
        self.upload()

    def upload(self, non_blocking: bool = True):
        for k, v in self.host.items():
            self.dev[k].copy_(v, non_blocking=non_blocking)

    @property
    def h2d_bytes(self) -> int:
        return sum(v.numel() * v.element_size() for v in self.host.values())

    def c_struct(self) -> _lib.XpBatch:
        d = self.dev
        return _lib.XpBatch(self.P, self.G, self.C, self.max_reads, self.R, d["y"].data_ptr(),
                            d["read_off"].data_ptr(), d["grp_len"].data_ptr(), d["kmer_mean"].data_ptr(),
                            d["kmer_std"].data_ptr(), d["grp_cond"].data_ptr(), self.y_format, 0, self.y_scale)


class DeviceResult:
    def __init__(self, P: int, G: int, Cn: int, device, with_coef: bool = False):
        import torch
        self.torch = torch
        self.P, self.G, self.C = P, G, Cn
        self.status = torch.zeros(P, dtype=torch.int32, device=device)  # bit field (uint32 on the C side)
        self.n_iter = torch.zeros(P, dtype=torch.int32, device=device)
        self.pval = torch.full((P,), math.nan, dtype=torch.float64, device=device)
        self.fit = torch.full((P, fit_width(G)), math.nan, dtype=torch.float64, device=device)
        self.stats = torch.full((P, stats_width(G, Cn)), math.nan, dtype=torch.float64, device=device)
        # coefficients of the last E-step, only kept for --save_models (xpb200_posterior_device)
        self.coef = torch.full((P, G + 2), math.nan, dtype=torch.float64, device=device) if with_coef else None
        self.n_fitted = torch.zeros(1, dtype=torch.int32, device=device)
        self.workspace = torch.empty(int(_lib.lib().xpb200_workspace_bytes(P)), dtype=torch.uint8, device=device)

    def c_struct(self) -> _lib.XpResult:
        return _lib.XpResult(self.status.data_ptr(), self.n_iter.data_ptr(), self.pval.data_ptr(),
                             self.fit.data_ptr(), self.stats.data_ptr(),
                             self.coef.data_ptr() if self.coef is not None else None)

    def to_host(self) -> FitResult:
        return FitResult(self.G, self.C, self.status.cpu().numpy().view(np.uint32), self.n_iter.cpu().numpy(),
                         self.pval.cpu().numpy(), self.fit.cpu().numpy(), self.stats.cpu().numpy(),
                         int(self.n_fitted.item()))


def fit_device(dbatch: DeviceBatch, method: Method, out: DeviceResult | None = None) -> DeviceResult:
    """Enqueue the whole path on the current torch stream of ``dbatch.device``; no synchronisation."""
    torch = dbatch.torch
    L = _lib.lib()
    if out is None:
        out = DeviceResult(dbatch.P, dbatch.G, dbatch.C, dbatch.device)
    b, m, r = dbatch.c_struct(), method.c_struct(), out.c_struct()
    with torch.cuda.device(dbatch.device):
        stream = torch.cuda.current_stream(dbatch.device).cuda_stream
        _lib.check(L.xpb200_fit_device(C.byref(b), C.byref(m), C.byref(r), out.workspace.data_ptr(),
                                       out.workspace.numel(), out.n_fitted.data_ptr(), C.c_void_p(stream)))
    return out


def posterior_device(dbatch: DeviceBatch, out: DeviceResult):
    """The z node of every fitted position (``xpb200_posterior_device``): ``(prob0[R], ln_prob[R, 2])`` device
    tensors, NaN for reads of positions that were not fitted.  ``out`` must come from a fit with
    ``DeviceResult(..., with_coef=True)``."""
    torch = dbatch.torch
    assert out.coef is not None, "fit with DeviceResult(with_coef=True) first"
    R = int(dbatch.R)
    prob0 = torch.full((max(R, 1),), math.nan, dtype=torch.float64, device=dbatch.device)
    ln_prob = torch.full((max(R, 1), 2), math.nan, dtype=torch.float64, device=dbatch.device)
    b, r = dbatch.c_struct(), out.c_struct()
    with torch.cuda.device(dbatch.device):
        stream = torch.cuda.current_stream().cuda_stream
        _lib.check(_lib.lib().xpb200_posterior_device(C.byref(b), C.byref(r), prob0.data_ptr(), ln_prob.data_ptr(),
                                                      C.c_void_p(stream)))
    return prob0[:R], ln_prob[:R]


def pack_records_device(out: DeviceResult, index_offset: int = 0, buf=None):
    """Compact [n_fitted, 4+FW+SW] record matrix of the fitted positions on the device (see
    ``xpb200_pack_records``); synchronises only to read the fitted count."""
    torch = out.torch
    L = _lib.lib()
    W = 4 + fit_width(out.G) + stats_width(out.G, out.C)
    if buf is None or buf.shape[0] < out.P or buf.shape[1] != W:
        buf = torch.empty((max(out.P, 1), W), dtype=torch.float64, device=out.status.device)
    r = out.c_struct()
    with torch.cuda.device(out.status.device):
        stream = torch.cuda.current_stream(out.status.device).cuda_stream
        _lib.check(L.xpb200_pack_records(C.byref(r), out.P, out.G, out.C, out.workspace.data_ptr(), int(index_offset),
                                         buf.data_ptr(), buf.shape[0], C.c_void_p(stream)))
    n = int(out.n_fitted.item())
    return buf[:n], buf
