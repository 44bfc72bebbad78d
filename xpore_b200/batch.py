"""CSR-packed ragged batch of positions — the host side of the device boundary.

The reference assembles one Python dict per position (``xpore/diffmod/io.py:21-90``,
``load_data``): ``y`` (reads, concatenated run by run), one-hot ``x`` (conditions) and ``r``
(runs).  Here the same selection rules produce flat arrays the kernels read directly:

``y``          f64[R]      reads of all positions; per position contiguous per *group slot*
``read_off``   i64[P+1]    first read of each position
``grp_len``    i32[P, G]   reads of each group slot at the position, 0 = slot absent
``kmer_mean``  f64[P]      unmodified k-mer mean  (prior location, conf_mu)
``kmer_std``   f64[P]      unmodified k-mer stdv  (alpha0 = 1/std^2, beta0 = std^2/2)

A *group slot* is a run (default) or a condition (``pooling``); slots are global to the
batch (``Layout``) so that a missing/filtered run is simply a zero-length slot.  Conditions
are indexed by their rank among the **sorted** condition names because every consumer in
the reference sorts them (``io.py:12-18`` get_dummies, ``io.py:182-187``).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from itertools import combinations

import numpy as np


@dataclass
class Layout:
    """Group slots of a batch, derived from the YAML ``data`` section.

    ``data_info``: ``{condition: {run_name: dir}}`` in config order
    (``configurator.py:29-34``).
    """

    condition_names: list  # sorted (io.py:184)
    run_names: list  # config order (io.py:185) — the table's column order
    run_cond: list  # condition index (sorted rank) of each run, config order
    pooling: bool

    @classmethod
    def from_data_info(cls, data_info, pooling: bool) -> "Layout":
        conds = sorted(data_info.keys())
        rank = {c: i for i, c in enumerate(conds)}
        runs, run_cond = [], []
        for c, rd in data_info.items():
            for r in rd.keys():
                runs.append(r)
                run_cond.append(rank[c])
        return cls(conds, runs, run_cond, bool(pooling))

    @property
    def n_conditions(self) -> int:
        return len(self.condition_names)

    @property
    def group_names(self) -> list:
        return self.condition_names if self.pooling else self.run_names

    @property
    def n_groups(self) -> int:
        return len(self.group_names)

    @property
    def grp_cond(self) -> np.ndarray:
        if self.pooling:
            return np.arange(self.n_conditions, dtype=np.int32)
        return np.asarray(self.run_cond, dtype=np.int32)

    @property
    def pairs(self) -> list:
        """Pairwise tests, ``itertools.combinations`` of the sorted names (``io.py:293``)."""
        return list(combinations(range(self.n_conditions), 2))


@dataclass
class Batch:
    layout: Layout
    y: np.ndarray
    read_off: np.ndarray
    grp_len: np.ndarray
    kmer_mean: np.ndarray
    kmer_std: np.ndarray
    # host-only bookkeeping for the table (id, position, kmer are strings; io.py:348)
    ids: list = field(default_factory=list)
    positions: list = field(default_factory=list)
    kmers: list = field(default_factory=list)
    # optional compact encoding of y (see include/xpore_b200.h: XPB200_Y_*)
    y_enc: np.ndarray | None = None
    y_format: int = 0
    y_scale: float = 1.0
    # set by the native reader: lazily decoded keys, row of every position's k-mer in the model, deepest position
    keys: object = None
    kmer_row: np.ndarray | None = None
    max_reads: int | None = None

    @property
    def n_positions(self) -> int:
        return int(self.read_off.shape[0] - 1)

    @property
    def n_reads(self) -> int:
        return int(self.read_off[-1])

    def validate(self):
        P, G = self.grp_len.shape
        assert G == self.layout.n_groups
        assert self.read_off.shape == (P + 1,) and self.read_off.dtype == np.int64
        assert self.grp_len.dtype == np.int32
        assert (np.diff(self.read_off) == self.grp_len.sum(axis=1)).all()
        if self.y is None:  # staged batch: only the scaled-integer codes exist (NativeIngest.read(staging=...))
            assert self.y_enc is not None and self.y_format != 0 and self.y_enc.shape[0] >= self.read_off[-1]
        else:
            assert self.y.dtype == np.float64 and self.y.shape[0] >= self.read_off[-1]
        assert self.kmer_mean.shape == (P,) and self.kmer_std.shape == (P,)
        return self

    def encode_reads(self, scale: float = 100.0) -> int:
        """Try to store the reads as scaled integers q = rint(y*scale): dataprep rounds event means
        to 2 decimals (``dataprep.py:638``), so q / scale reproduces y bit for bit and the batch
        needs 2 (u16) or 4 (i32) bytes per read instead of 8.  Falls back to f64 when any value
        does not round-trip.  Returns the chosen format code."""
        Y_F64, Y_I32, Y_U16 = 0, 1, 2
        self.y_enc, self.y_format, self.y_scale = None, Y_F64, 1.0
        y = self.y[:self.n_reads]
        if y.size == 0:
            return Y_F64
        q = np.rint(y * scale)
        if not (np.abs(q) < 2 ** 31 - 1).all() or not np.array_equal(q / scale, y):
            return Y_F64
        if q.min() >= 0 and q.max() < 65536:
            enc, fmt, per16 = q.astype(np.uint16), Y_U16, 8
        else:
            enc, fmt, per16 = q.astype(np.int32), Y_I32, 4
        pad = (-len(enc)) % per16 + per16
        self.y_enc = np.concatenate([enc, np.zeros(pad, enc.dtype)])
        self.y_format, self.y_scale = fmt, float(scale)
        return fmt

    def subset(self, index) -> "Batch":
        """Positions ``index`` (ascending array) re-packed — used for sharding across ranks."""
        index = np.asarray(index, dtype=np.int64)
        n = np.diff(self.read_off)[index]
        off = np.zeros(len(index) + 1, dtype=np.int64)
        np.cumsum(n, out=off[1:])
        y = np.empty(int(off[-1]), dtype=np.float64)
        src0 = self.read_off[index]
        # gather ragged rows without a Python loop over reads
        if len(index):
            rep = np.repeat(np.arange(len(index)), n)
            within = np.arange(int(off[-1]), dtype=np.int64) - off[rep]
            y[:] = self.y[src0[rep] + within]
        pick = lambda lst: [lst[i] for i in index] if lst else []
        return Batch(self.layout, y, off, np.ascontiguousarray(self.grp_len[index]),
                     np.ascontiguousarray(self.kmer_mean[index]), np.ascontiguousarray(self.kmer_std[index]),
                     pick(self.ids), pick(self.positions), pick(self.kmers))


def _pos_sort_key(pk):
    pos, kmer = pk
    try:
        return (0, int(pos), kmer)
    except ValueError:
        return (1, 0, str(pos) + kmer)


class BatchBuilder:
    """Accumulates selected positions gene by gene (reference seam: ``diffmod.py:execute``
    -> ``io.load_data``) and emits one :class:`Batch`."""

    def __init__(self, layout: Layout, kmer_model, min_count: int, max_count: int):
        self.layout = layout
        self.kmer_model = kmer_model
        self.min_count = min_count
        self.max_count = max_count
        self._y, self._glen = [], []
        self._km, self._ks = [], []
        self.ids, self.positions, self.kmers = [], [], []

    def add_gene(self, idx: str, data_dict) -> int:
        """``data_dict``: ``{(condition, run_name): parsed-json-or-None}`` in config order
        (``diffmod.py:156-169``).  Applies the selection rules of ``io.py:31-80`` and returns the
        number of positions kept.  Positions are emitted in (int(position), kmer) order — the
        reference iterates a Python ``set`` (hash-seed dependent order, SURVEY.md P-order)."""
        lay = self.layout
        pooling = lay.pooling
        lo, hi = self.min_count, self.max_count
        cond_rank = {c: i for i, c in enumerate(lay.condition_names)}
        run_slot = {r: i for i, r in enumerate(lay.run_names)}

        runs = []  # (cond_idx, slot, per-position dict)
        pair_sets = []
        for (cond, run), d in data_dict.items():
            if d is None:
                continue
            g = d[idx]
            runs.append((cond_rank[cond], run_slot[run], g))
            pair_sets.append({(pos, kmer) for pos, km in g.items() for kmer in km.keys()})
        if not pair_sets:
            return 0
        common = set.intersection(*pair_sets)
        kept = 0
        C = lay.n_conditions
        G = lay.n_groups
        for pos, kmer in sorted(common, key=_pos_sort_key):
            per_slot = [None] * G if not pooling else [[] for _ in range(G)]
            glen = np.zeros(G, dtype=np.int32)
            n_cond_sum = [0] * C
            cond_has_run = [False] * C
            total = 0
            for c, slot, g in runs:
                reads = g[pos][kmer]
                n = len(reads)
                if (not pooling) and (n < lo or n > hi):
                    continue
                n_cond_sum[c] += n
                cond_has_run[c] = True
                total += n
                if pooling:
                    per_slot[c].append(reads)
                    glen[c] += n
                else:
                    per_slot[slot] = reads
                    glen[slot] = n
            if total == 0:
                continue
            if pooling:
                n_incl = sum(1 for c in range(C) if lo <= n_cond_sum[c] <= hi)
            else:
                n_incl = sum(1 for c in range(C) if cond_has_run[c])
            if n_incl < 2:
                continue
            m, s = self.kmer_model.signal(kmer)  # KeyError for an unknown k-mer, as .loc would
            if pooling:
                for lst in per_slot:
                    for reads in lst:
                        self._y.append(np.asarray(reads, dtype=np.float64))
            else:
                for reads in per_slot:
                    if reads is not None and len(reads):
                        self._y.append(np.asarray(reads, dtype=np.float64))
            self._glen.append(glen)
            self._km.append(m)
            self._ks.append(s)
            self.ids.append(idx)
            self.positions.append(pos)
            self.kmers.append(kmer)
            kept += 1
        return kept

    @property
    def n_positions(self) -> int:
        return len(self._glen)

    def build(self) -> Batch:
        G = self.layout.n_groups
        P = len(self._glen)
        grp_len = np.stack(self._glen).astype(np.int32) if P else np.zeros((0, G), np.int32)
        read_off = np.zeros(P + 1, dtype=np.int64)
        np.cumsum(grp_len.sum(axis=1, dtype=np.int64), out=read_off[1:])
        y = np.concatenate(self._y) if self._y else np.zeros(0, np.float64)
        return Batch(self.layout, np.ascontiguousarray(y, dtype=np.float64), read_off, grp_len,
                     np.asarray(self._km, dtype=np.float64), np.asarray(self._ks, dtype=np.float64),
                     list(self.ids), list(self.positions), list(self.kmers)).validate()
