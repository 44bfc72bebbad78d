"""Unmodified-RNA k-mer signal model (mean / stdv per 5-mer).

The reference reads ``model_kmer.csv`` with pandas and looks rows up by k-mer string
(``xpore/scripts/diffmod.py:95`` and ``:23``); the YAML key ``prior:`` may point at a
replacement CSV (``xpore/diffmod/configurator.py:19-22``).  Here the shipped table lives in
``data/model_kmer.npz`` (made by ``tools/make_kmer_model.py``) and a CSV override is parsed
with the ``csv`` module.  Lookups are vectorised: k-mer string -> row id -> (mean, stdv).
"""
from __future__ import annotations

import csv
import os

import numpy as np

_DEFAULT = os.path.join(os.path.dirname(__file__), "data", "model_kmer.npz")


class KmerModel:
    def __init__(self, kmers, mean, stdv, path):
        self.kmers = [k.decode() if isinstance(k, bytes) else str(k) for k in kmers]
        self.mean = np.ascontiguousarray(mean, dtype=np.float64)
        self.stdv = np.ascontiguousarray(stdv, dtype=np.float64)
        self.path = path
        self._row = {k: i for i, k in enumerate(self.kmers)}

    @classmethod
    def load(cls, path: str | None = None) -> "KmerModel":
        if path is None:
            z = np.load(_DEFAULT)
            return cls(z["kmer"], z["mean"], z["stdv"], _DEFAULT)
        kmers, mean, stdv = [], [], []
        with open(path) as f:
            rd = csv.DictReader(f)
            for row in rd:
                kmers.append(row["model_kmer"])
                mean.append(float(row["model_mean"]))
                stdv.append(float(row["model_stdv"]))
        return cls(kmers, mean, stdv, os.path.abspath(path))

    def row(self, kmer: str) -> int:
        """Row id of a k-mer; KeyError for an unknown k-mer like ``model_kmer.loc[kmer]``."""
        return self._row[kmer]

    def rows(self, kmers) -> np.ndarray:
        return np.fromiter((self._row[k] for k in kmers), dtype=np.int64, count=len(kmers))

    def signal(self, kmer: str):
        i = self._row[kmer]
        return float(self.mean[i]), float(self.stdv[i])

    def usable_rows(self) -> np.ndarray:
        """Rows whose k-mer has no ``N`` (dataprep never emits the others, ``dataprep.py:473``)."""
        return np.array([i for i, k in enumerate(self.kmers) if "N" not in k], dtype=np.int64)
