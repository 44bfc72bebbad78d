"""Native ingest: ``data.json`` byte spans -> :class:`~xpore_b200.batch.Batch` through the C++ reader
of ``libxpore_b200.so`` (``csrc/xp_ingest.cpp``; C ABI in ``include/xpore_b200.h``).

This is the batched, multi-threaded replacement of the reference's per-gene
``seek/read/ujson.loads`` (``xpore/scripts/diffmod.py:156-169``) + ``io.load_data``
(``xpore/diffmod/io.py:21-90``).  :class:`xpore_b200.batch.BatchBuilder` is the Python statement
of the same selection rules; ``tests/test_ingest.py`` checks the two against each other.

Two ways to take the result of :meth:`NativeIngest.read`:

* default: a self-contained :class:`Batch` with the reads as float64 (tests, ``--save_models``, the cache writer);
* ``staging=``: the arrays land in page-locked buffers (:class:`Staging`) that the host entry points of the
  library copy from at full PCIe rate, the reads as the u16 codes the parser took from the text itself
  (2 bytes per read, no float64 copy is ever made), and the (id, position, kmer) strings stay one blob that is
  decoded only for the rows that reach ``diffmod.table``.  This is the path of the ``diffmod`` driver.
"""
from __future__ import annotations

import ctypes as C
import os
from collections.abc import Sequence

import numpy as np

from . import _lib
from .batch import Batch, Layout


class PinnedBuffer:
    """A growable page-locked byte buffer (``xpb200_host_alloc``); pageable numpy memory when no CUDA device is
    usable (CPU-only tests of the driver logic)."""

    def __init__(self):
        self.ptr, self.cap, self._np = None, 0, None

    def view(self, nbytes: int, dtype, shape):
        nbytes = int(nbytes)
        if nbytes > self.cap:
            self.free()
            cap = max(nbytes + nbytes // 4, 1 << 16)
            L = _lib.lib()
            ptr = L.xpb200_host_alloc(cap)
            if ptr:
                self.ptr = ptr
                self._np = np.frombuffer((C.c_uint8 * cap).from_address(ptr), dtype=np.uint8)
            else:
                self._np = np.empty(cap, dtype=np.uint8)
            self.cap = cap
        return self._np[:nbytes].view(dtype).reshape(shape)

    def free(self):
        if self.ptr:
            _lib.lib().xpb200_host_free(self.ptr)
        self.ptr, self.cap, self._np = None, 0, None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Staging:
    """Page-locked buffers for the arrays of one batch in flight (reused chunk after chunk)."""

    NAMES = ("y", "read_off", "grp_len", "kmer_mean", "kmer_std", "rows")

    def __init__(self):
        self.buf = {n: PinnedBuffer() for n in self.NAMES}

    def array(self, name, dtype, shape):
        n = int(np.prod(shape)) * np.dtype(dtype).itemsize
        return self.buf[name].view(n, dtype, shape)

    def free(self):
        for b in self.buf.values():
            b.free()


class KeyStrings(Sequence):
    """Column ``which`` (0 id, 1 position, 2 kmer) of the keys of a batch, decoded on demand from the reader's
    string blob (``<position>\\0<kmer>\\0`` per position)."""

    def __init__(self, table, which):
        self.t, self.which = table, which

    def __len__(self):
        return self.t.n

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self[j] for j in range(*i.indices(len(self)))]
        return self.t.key(int(i))[self.which]

    def __eq__(self, other):
        return list(self) == list(other)


class KeyTable:
    def __init__(self, ids, gene, blob: bytes):
        self.ids, self.gene, self.blob = ids, gene, blob
        self.n = len(gene)
        z = np.flatnonzero(np.frombuffer(blob, dtype=np.uint8) == 0) if blob else np.zeros(0, np.int64)
        self.end = z.astype(np.int64)                      # end[2p] closes the position, end[2p+1] the kmer
        self.start = np.concatenate([[0], z[:-1] + 1]).astype(np.int64) if len(z) else self.end

    def key(self, p: int):
        if p < 0:
            p += self.n
        a, b = 2 * p, 2 * p + 1
        return (self.ids[self.gene[p]], self.blob[self.start[a]:self.end[a]].decode(),
                self.blob[self.start[b]:self.end[b]].decode())


class NativeIngest:
    def __init__(self, data_info, layout: Layout, kmer_model, min_count: int, max_count: int):
        self.L = _lib.lib()
        self.layout = layout
        self.km = kmer_model
        self.runs = [(cond, run, d) for cond, rd in data_info.items() for run, d in rd.items()]
        cond_rank = {c: i for i, c in enumerate(layout.condition_names)}
        run_slot = {r: i for i, r in enumerate(layout.run_names)}
        n = len(self.runs)
        paths = (C.c_char_p * n)(*[os.path.join(d, "data.json").encode() for _, _, d in self.runs])
        slot = np.asarray([run_slot[r] for _, r, _ in self.runs], dtype=np.int32)
        cond = np.asarray([cond_rank[c] for c, _, _ in self.runs], dtype=np.int32)
        kmers = (C.c_char_p * len(kmer_model.kmers))(*[k.encode() for k in kmer_model.kmers])
        self.handle = self.L.xpb200_ingest_open(n, paths, slot.ctypes.data, cond.ctypes.data, layout.n_groups,
                                                layout.n_conditions, int(layout.pooling), int(min_count),
                                                int(max_count), len(kmer_model.kmers), kmers)
        if not self.handle:
            raise RuntimeError(self.L.xpb200_ingest_last_error().decode())

    def close(self):
        if self.handle:
            self.L.xpb200_ingest_close(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def read(self, ids, f_index, n_threads: int = 1, staging: Staging | None = None, before_alloc=None) -> Batch:
        """Selected positions of ``ids`` (in that order) as one batch.  ``f_index``: {run: {idx: (start, end)}}.
        ``before_alloc`` is called between the parse and the first allocation of result buffers (the driver waits for
        its CUDA context there: page-locked staging needs it, the parse does not)."""
        import time
        t0 = time.perf_counter()
        n, R = len(ids), len(self.runs)
        starts = np.full((n, R), -1, dtype=np.int64)
        ends = np.zeros((n, R), dtype=np.int64)
        for r, (_, run, _) in enumerate(self.runs):
            fi = f_index[run]
            for g, idx in enumerate(ids):
                span = fi.get(idx)
                if span is not None:
                    starts[g, r], ends[g, r] = span
        P, Rd, S = C.c_int64(0), C.c_int64(0), C.c_int64(0)
        rc = self.L.xpb200_ingest_genes(self.handle, n, starts.ctypes.data, ends.ctypes.data, int(max(1, n_threads)),
                                        C.byref(P), C.byref(Rd), C.byref(S))
        if rc != 0:
            msg = self.L.xpb200_ingest_last_error().decode()
            if msg.startswith("k-mer not in the model"):
                raise KeyError(msg)
            raise RuntimeError(msg)
        P, Rd, S = P.value, Rd.value, S.value
        t1 = time.perf_counter()
        if before_alloc is not None:
            before_alloc()
        G = self.layout.n_groups
        max_reads = C.c_int32(0)
        enc = self.L.xpb200_ingest_encoding(self.handle, C.byref(max_reads))
        new = (lambda name, dtype, shape: staging.array(name, dtype, shape)) if staging is not None else \
              (lambda name, dtype, shape: np.empty(shape, dtype=dtype))
        read_off = new("read_off", np.int64, (P + 1,))
        grp_len = new("grp_len", np.int32, (P, G))
        row = np.zeros(P, dtype=np.int32)
        gene = np.zeros(P, dtype=np.int32)
        strings = C.create_string_buffer(max(S, 1))
        u16 = staging is not None and enc == 2 and Rd > 0
        y = None if u16 else np.empty(Rd, dtype=np.float64)
        rc = self.L.xpb200_ingest_copy(self.handle, None if u16 else y.ctypes.data, read_off.ctypes.data, grp_len.ctypes.data,
                                       row.ctypes.data, gene.ctypes.data, C.addressof(strings))
        if rc != 0:
            raise RuntimeError(self.L.xpb200_ingest_last_error().decode())
        kmean, kstd = new("kmer_mean", np.float64, (P,)), new("kmer_std", np.float64, (P,))
        np.take(self.km.mean, row, out=kmean)
        np.take(self.km.stdv, row, out=kstd)
        keys = KeyTable(list(ids), gene, strings.raw[:S])
        b = Batch(self.layout, y, read_off, grp_len, kmean, kstd, KeyStrings(keys, 0), KeyStrings(keys, 1),
                  KeyStrings(keys, 2))
        b.keys, b.kmer_row, b.max_reads = keys, row, int(max_reads.value)
        if u16:
            yq = staging.array("y", np.uint16, (Rd + 8 + (-Rd) % 8,))
            yq[Rd:] = 0
            if self.L.xpb200_ingest_copy_u16(self.handle, yq.ctypes.data) != 0:
                raise RuntimeError(self.L.xpb200_ingest_last_error().decode())
            b.y_enc, b.y_format, b.y_scale = yq, 2, 100.0
        self.last_times = {"parse_s": t1 - t0, "copy_s": time.perf_counter() - t1}
        return b.validate()
