"""Binary hand-off of the diffmod input: the CSR batch of *every* id of a run set in one file.

``data.json`` is text (one JSON object per id, ``dataprep.py:485-495,642-651``); parsing it is the
end-to-end bottleneck once the fit runs on a GPU.  A producer - the
synthetic generator here, a dataprep-compatible writer elsewhere - can run :func:`build_cache` once
next to ``data.json``/``data.index``; ``xpore diffmod --csr_cache FILE`` then maps the arrays
straight from the file (``numpy.memmap``, no parsing, no selection pass) and still writes the same
``diffmod.table``.  ``data.json`` stays the drop-in input; the cache is optional and self-checking.

File layout (little endian)::

    8 bytes   magic  b"XPB200C1"
    8 bytes   header length H (uint64)
    H bytes   JSON header: layout (conditions, runs, grp_cond, pooling), criteria, sources
              (path, size, mtime_ns of every data.json / data.index), k-mer model digest,
              ids, and for every array its dtype / shape / byte offset
    ...       arrays, each 64-byte aligned: y (u16 | i32 scaled by 100, or f64), read_off i64[P+1],
              grp_len i32[P,G], kmer_mean f64[P], kmer_std f64[P], gene i32[P] (index into ids),
              strings (position \\0 kmer \\0 ... as in xpb200_ingest_copy)

A cache is used only if its sources (size and mtime of every input file), criteria, pooling mode and
k-mer model digest match the current configuration; otherwise :func:`load_cache` returns ``None``
and the caller falls back to ``data.json``.
"""
from __future__ import annotations

import hashlib
import json
import os
import struct
from collections import OrderedDict

import numpy as np

from .batch import Batch, Layout

MAGIC = b"XPB200C1"
ALIGN = 64


def _sources(data_info):
    out = []
    for cond, runs in data_info.items():
        for run, d in runs.items():
            for name in ("data.json", "data.index"):
                p = os.path.join(d, name)
                st = os.stat(p)
                out.append([cond, run, name, int(st.st_size), int(st.st_mtime_ns)])
    return out


def _km_digest(km) -> str:
    h = hashlib.sha256()
    h.update("\0".join(km.kmers).encode())
    h.update(np.ascontiguousarray(km.mean, dtype=np.float64).tobytes())
    h.update(np.ascontiguousarray(km.stdv, dtype=np.float64).tobytes())
    return h.hexdigest()


def _key(data_info, layout: Layout, criteria, km):
    return {"sources": _sources(data_info),
            "conditions": list(layout.condition_names), "runs": list(layout.run_names),
            "group_names": list(layout.group_names), "grp_cond": [int(c) for c in layout.grp_cond],
            "pooling": bool(layout.pooling),
            "readcount_min": int(criteria["readcount_min"]), "readcount_max": int(criteria["readcount_max"]),
            "kmer_model": _km_digest(km)}


def write_cache(path: str, batch: Batch, ids, gene: np.ndarray, key: dict) -> None:
    """Write ``batch`` (positions of ``ids``, ``gene[p]`` = index of position p's id) to ``path``."""
    fmt = batch.encode_reads()
    y = batch.y_enc[:batch.n_reads] if fmt != 0 else batch.y[:batch.n_reads]
    strings = ("\0".join(x for pair in zip(batch.positions, batch.kmers) for x in pair) + "\0").encode() \
        if batch.n_positions else b""
    arrays = OrderedDict([
        ("y", np.ascontiguousarray(y)),
        ("read_off", np.ascontiguousarray(batch.read_off, dtype=np.int64)),
        ("grp_len", np.ascontiguousarray(batch.grp_len, dtype=np.int32)),
        ("kmer_mean", np.ascontiguousarray(batch.kmer_mean, dtype=np.float64)),
        ("kmer_std", np.ascontiguousarray(batch.kmer_std, dtype=np.float64)),
        ("gene", np.ascontiguousarray(gene, dtype=np.int32)),
        ("strings", np.frombuffer(strings, dtype=np.uint8)),
    ])
    header = {"key": key, "ids": list(ids), "y_format": int(fmt), "y_scale": float(batch.y_scale), "arrays": {}}
    # two passes: the header's own length shifts the offsets
    for _ in range(3):
        blob = json.dumps(header).encode()
        off = (len(MAGIC) + 8 + len(blob) + ALIGN - 1) // ALIGN * ALIGN
        for name, a in arrays.items():
            header["arrays"][name] = {"dtype": a.dtype.str, "shape": list(a.shape), "offset": off}
            off = (off + a.nbytes + ALIGN - 1) // ALIGN * ALIGN
    blob = json.dumps(header).encode()
    tmp = path + ".tmp"
    with open(tmp, "wb") as f:
        f.write(MAGIC)
        f.write(struct.pack("<Q", len(blob)))
        f.write(blob)
        for name, a in arrays.items():
            f.seek(header["arrays"][name]["offset"])
            f.write(a.tobytes())
        f.truncate(max(f.tell(), off))
    os.replace(tmp, path)


def build_cache(config_path: str, out_path: str, n_threads: int = 1, ids=None) -> dict:
    """Ingest every id of the configuration (native reader) and write the cache.  Returns counts."""
    from .config import Configurator
    from .diffmod import get_ids, read_index
    from .ingest import NativeIngest
    from .kmer_model import KmerModel

    config = Configurator(config_path)
    data_info, method, criteria = config.get_data_info(), config.get_method(), config.get_criteria()
    km = KmerModel.load(config.get_paths()["model_kmer"])
    layout = Layout.from_data_info(data_info, method["pooling"])
    f_index = OrderedDict()
    for cond, runs in data_info.items():
        for run, d in runs.items():
            f_index[run] = read_index(d)
    ids = list(ids) if ids else get_ids(f_index, data_info)
    reader = NativeIngest(data_info, layout, km, criteria["readcount_min"], criteria["readcount_max"])
    batch = reader.read(ids, f_index, n_threads=n_threads)
    reader.close()
    index_of = {idx: i for i, idx in enumerate(ids)}
    gene = np.asarray([index_of[i] for i in batch.ids], dtype=np.int32)
    write_cache(out_path, batch, ids, gene, _key(data_info, layout, criteria, km))
    return {"ids": len(ids), "positions": batch.n_positions, "reads": batch.n_reads, "y_format": batch.y_format}


class Cache:
    """An opened cache: arrays are memory-mapped; :meth:`batch` cuts the positions of a run of ids."""

    def __init__(self, path: str, header: dict):
        self.path, self.header = path, header
        self.ids = header["ids"]
        self._a = {}
        for name, d in header["arrays"].items():
            n = int(np.prod(d["shape"])) if d["shape"] else 1
            self._a[name] = (np.memmap(path, mode="r", dtype=np.dtype(d["dtype"]), offset=d["offset"], shape=tuple(d["shape"]))
                             if n else np.zeros(tuple(d["shape"]), dtype=np.dtype(d["dtype"])))
        gene = np.asarray(self._a["gene"])
        # positions of id i are the contiguous range [first[i], first[i+1])
        self.first = np.searchsorted(gene, np.arange(len(self.ids) + 1), side="left")
        s = bytes(self._a["strings"])
        parts = s.split(b"\0")[:-1] if s else []
        self.positions = [p.decode() for p in parts[0::2]]
        self.kmers = [k.decode() for k in parts[1::2]]
        self.index_of = {idx: i for i, idx in enumerate(self.ids)}

    def batch(self, ids, layout: Layout) -> Batch:
        """Batch of the positions of ``ids`` (each must be in the cache), in that order."""
        fmt, scale = self.header["y_format"], self.header["y_scale"]
        read_off_all = self._a["read_off"]
        spans = [(int(self.first[self.index_of[i]]), int(self.first[self.index_of[i] + 1])) for i in ids]
        P = sum(b - a for a, b in spans)
        G = layout.n_groups
        if P == 0:
            return Batch(layout, np.zeros(0), np.zeros(1, dtype=np.int64), np.zeros((0, G), dtype=np.int32),
                         np.zeros(0), np.zeros(0), [], [], []).validate()
        ys, lens, gl, kmn, ksd, bid, pos, kmr = [], [], [], [], [], [], [], []
        for idx, (a, b) in zip(ids, spans):
            if b == a:
                continue
            r0, r1 = int(read_off_all[a]), int(read_off_all[b])
            ys.append(np.asarray(self._a["y"][r0:r1]))
            lens.append(np.diff(np.asarray(read_off_all[a:b + 1])))
            gl.append(np.asarray(self._a["grp_len"][a:b]))
            kmn.append(np.asarray(self._a["kmer_mean"][a:b]))
            ksd.append(np.asarray(self._a["kmer_std"][a:b]))
            bid.extend([idx] * (b - a))
            pos.extend(self.positions[a:b])
            kmr.extend(self.kmers[a:b])
        y_raw = np.concatenate(ys)
        read_off = np.zeros(P + 1, dtype=np.int64)
        np.cumsum(np.concatenate(lens), out=read_off[1:])
        # doubles exactly as the parser would have produced them: q / scale is the correctly rounded literal
        y = y_raw.astype(np.float64) / scale if fmt != 0 else np.ascontiguousarray(y_raw, dtype=np.float64)
        bt = Batch(layout, y, read_off, np.ascontiguousarray(np.concatenate(gl)), np.concatenate(kmn),
                   np.concatenate(ksd), bid, pos, kmr).validate()
        if fmt != 0:  # hand the compact encoding through (padded like Batch.encode_reads): no second pass
            per16 = 16 // y_raw.dtype.itemsize
            pad = (-len(y_raw)) % per16 + per16
            bt.y_enc = np.concatenate([y_raw, np.zeros(pad, y_raw.dtype)])
            bt.y_format, bt.y_scale = fmt, scale
        return bt


def load_cache(path: str, data_info, layout: Layout, criteria, km):
    """Open ``path`` if it exists and matches the current inputs; ``None`` otherwise."""
    if not path or not os.path.exists(path):
        return None
    with open(path, "rb") as f:
        if f.read(len(MAGIC)) != MAGIC:
            return None
        (hlen,) = struct.unpack("<Q", f.read(8))
        header = json.loads(f.read(hlen).decode())
    try:
        if header["key"] != json.loads(json.dumps(_key(data_info, layout, criteria, km))):
            return None
    except OSError:
        return None
    return Cache(path, header)


def main(argv=None):
    import argparse
    ap = argparse.ArgumentParser(prog="python -m xpore_b200.csrcache",
                                 description="write the binary CSR cache of a diffmod configuration")
    ap.add_argument("--config", required=True)
    ap.add_argument("--out", required=True)
    ap.add_argument("--n_processes", type=int, default=1)
    a = ap.parse_args(argv)
    print(build_cache(a.config, a.out, a.n_processes))


if __name__ == "__main__":
    main()
