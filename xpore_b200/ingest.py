"""Native ingest: ``data.json`` byte spans -> :class:`~xpore_b200.batch.Batch` through the C++ reader
of ``libxpore_b200.so`` (``csrc/xp_ingest.cpp``; C ABI in ``include/xpore_b200.h``).

This is the batched, multi-threaded replacement of the reference's per-gene
``seek/read/ujson.loads`` (``xpore/scripts/diffmod.py:156-169``) + ``io.load_data``
(``xpore/diffmod/io.py:21-90``).  :class:`xpore_b200.batch.BatchBuilder` is the Python statement
of the same selection rules; ``tests/test_ingest.py`` checks the two against each other.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _lib
from .batch import Batch, Layout


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

    def read(self, ids, f_index, n_threads: int = 1) -> Batch:
        """Selected positions of ``ids`` (in that order) as one batch.  ``f_index``: {run: {idx: (start, end)}}."""
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
        G = self.layout.n_groups
        y = np.empty(Rd, dtype=np.float64)
        read_off = np.zeros(P + 1, dtype=np.int64)
        grp_len = np.zeros((P, G), dtype=np.int32)
        row = np.zeros(P, dtype=np.int32)
        gene = np.zeros(P, dtype=np.int32)
        strings = C.create_string_buffer(max(S, 1))
        rc = self.L.xpb200_ingest_copy(self.handle, y.ctypes.data, read_off.ctypes.data, grp_len.ctypes.data,
                                       row.ctypes.data, gene.ctypes.data, C.addressof(strings))
        if rc != 0:
            raise RuntimeError(self.L.xpb200_ingest_last_error().decode())
        parts = strings.raw[:S].split(b"\0")
        positions = [p.decode() for p in parts[0:2 * P:2]]
        kmers = [k.decode() for k in parts[1:2 * P:2]]
        return Batch(self.layout, y, read_off, grp_len, self.km.mean[row], self.km.stdv[row],
                     [ids[g] for g in gene], positions, kmers).validate()
