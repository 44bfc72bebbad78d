"""Synthetic ``xpore dataprep`` output of the BASELINE.json shapes.

Two producers share one statistical model:

* :func:`write_dataset` — real ``data.json`` / ``data.index`` per run plus ``config.yml``
  (format of ``xpore/scripts/dataprep.py:485-495,642-651``), for the CLI, the golden fixtures
  and the reference runs;
* :func:`make_batch` — the same model written straight into a CSR :class:`~xpore_b200.batch.Batch`
  (vectorised; used by ``bench.py`` at 10^6 positions where a text round trip is pointless).

Model: k-mer per position uniform over the 1024 non-``N`` 5-mers; unmodified reads
~ N(mu_kmer, sigma_kmer); a fraction ``f_mod`` of positions is modifiable, there modified reads
~ N(mu_kmer +/- U(3,8), 1.3 sigma_kmer) with rate 0.6 in normal conditions and 0.05 in
"KO-like" ones; values rounded to 2 decimals (``dataprep.py:638``).
"""
from __future__ import annotations

import json
import math
import os
from collections import OrderedDict
from dataclasses import dataclass, field

import numpy as np
import yaml

from .batch import Batch, Layout
from .kmer_model import KmerModel


@dataclass
class Spec:
    conditions: dict  # {condition: n_replicates}, config order
    ko_like: tuple = ("KO",)
    n_genes: int = 1
    pos_per_gene: int = 100
    reads_lo: int = 50
    reads_hi: int = 200  # per run, inclusive
    f_mod: float = 0.4
    seed: int = 0
    loguniform_total: tuple | None = None  # (lo, hi): total reads/position log-uniform, split evenly
    method: dict = field(default_factory=dict)
    criteria: dict | None = None

    def data_info(self, root="."):
        info = OrderedDict()
        for c, nrep in self.conditions.items():
            info[c] = OrderedDict()
            for r in range(1, nrep + 1):
                info[c]["%s-rep%d" % (c, r)] = os.path.join(root, "%s_rep%d" % (c, r))
        return info


# The five BASELINE.json configs.
def config_spec(name: str) -> Spec:
    if name == "c1":
        return Spec({"KO": 2, "WT": 2}, n_genes=1, pos_per_gene=100, reads_lo=50, reads_hi=200, f_mod=0.4,
                    seed=0, method={"prefiltering": {"method": "t-test", "threshold": 0.1}})
    if name == "c2":
        return Spec({"KO": 3, "WT": 3}, n_genes=10000, pos_per_gene=100, reads_lo=20, reads_hi=100, f_mod=0.1,
                    seed=1, method={"prefiltering": {"method": "t-test", "threshold": 0.1}})
    if name == "c3":
        return Spec({"KO": 3, "WT": 3}, n_genes=10000, pos_per_gene=100, reads_lo=20, reads_hi=100, f_mod=0.1,
                    seed=1, method={"pooling": True, "max_iters": 500})
    if name == "c4":
        return Spec({"KO": 2, "WT": 2}, n_genes=1000, pos_per_gene=100, f_mod=0.3, seed=2,
                    loguniform_total=(20, 5000), criteria={"readcount_min": 5, "readcount_max": 5000})
    if name == "c5":
        conds = OrderedDict((c, 3) for c in ["A549", "HEK293T", "HEPG2", "K562", "KO", "MCF7"])
        return Spec(conds, n_genes=20000, pos_per_gene=100, reads_lo=20, reads_hi=60, f_mod=0.1, seed=3,
                    method={"prefiltering": {"method": "t-test", "threshold": 0.1}})
    raise KeyError(name)


def _counts(spec: Spec, rng, P: int, G: int) -> np.ndarray:
    if spec.loguniform_total is not None:
        lo, hi = spec.loguniform_total
        tot = np.exp(rng.uniform(np.log(lo), np.log(hi), size=P))
        return np.maximum(1, np.round(tot / G)).astype(np.int64)[:, None].repeat(G, axis=1)
    return rng.integers(spec.reads_lo, spec.reads_hi + 1, size=(P, G))


def _draw(spec: Spec, rng, km: KmerModel, P: int):
    """Per-position parameters and per-run read counts for P positions."""
    info = spec.data_info()
    lay = Layout.from_data_info(info, pooling=False)
    G = lay.n_groups
    rows = km.usable_rows()
    krow = rows[rng.integers(0, len(rows), size=P)]
    mu, sd = km.mean[krow], km.stdv[krow]
    is_mod = rng.random(P) < spec.f_mod
    shift = rng.uniform(3.0, 8.0, size=P) * np.where(rng.random(P) < 0.5, -1.0, 1.0)
    counts = _counts(spec, rng, P, G)
    ko = np.array([lay.condition_names[c] in spec.ko_like for c in lay.run_cond])
    rate = np.where(ko, 0.05, 0.6)[None, :] * is_mod[:, None]  # [P, G]
    return lay, krow, mu, sd, shift, counts, rate


def _reads(rng, mu, sd, shift, rate, n):
    """n reads of one (position, run)."""
    mod = rng.random(n) < rate
    y = rng.normal(mu, sd, size=n)
    if mod.any():
        y[mod] = rng.normal(mu + shift, 1.3 * sd, size=int(mod.sum()))
    return np.round(y, 2)


def write_dataset(root: str, spec: Spec, km: KmerModel | None = None, first_pos: int = 100) -> str:
    """Write ``<root>/<cond>_rep<i>/data.{json,index}`` and ``<root>/config.yml``; returns the
    config path.  ``out`` in the YAML is ``<root>/out``."""
    km = km or KmerModel.load()
    rng = np.random.default_rng(spec.seed)
    info = spec.data_info(root)
    P = spec.n_genes * spec.pos_per_gene
    lay, krow, mu, sd, shift, counts, rate = _draw(spec, rng, km, P)
    files = []
    for r, (c, run) in enumerate((c, run) for c, rd in info.items() for run in rd):
        d = info[c][run]
        os.makedirs(d, exist_ok=True)
        files.append((open(os.path.join(d, "data.json"), "w"), open(os.path.join(d, "data.index"), "w")))
        files[-1][1].write("idx,start,end\n")
    for gi in range(spec.n_genes):
        idx = "ENST%011d" % (gi + 1)
        for r, (fj, fi) in enumerate(files):
            data = OrderedDict()
            for j in range(spec.pos_per_gene):
                p = gi * spec.pos_per_gene + j
                y = _reads(rng, mu[p], sd[p], shift[p], rate[p, r], int(counts[p, r]))
                data[str(first_pos + j)] = {km.kmers[krow[p]]: [float(v) for v in y]}
            start = fj.tell()
            fj.write("{")
            fj.write('"%s":' % idx)
            json.dump(data, fj, separators=(",", ":"))
            fj.write("}\n")
            end = fj.tell()
            fi.write("%s,%d,%d\n" % (idx, start, end))
    for fj, fi in files:
        fj.close()
        fi.close()
    cfg = {"data": {c: {run.split("-", 1)[1]: d for run, d in rd.items()} for c, rd in info.items()},
           "out": os.path.join(root, "out")}
    if spec.method:
        cfg["method"] = dict(spec.method)
    if spec.criteria:
        cfg["criteria"] = dict(spec.criteria)
    path = os.path.join(root, "config.yml")
    with open(path, "w") as f:
        yaml.safe_dump(cfg, f, sort_keys=False)
    return path


def write_dataset_from_batch(root: str, spec: Spec, batch: Batch, km: KmerModel | None = None, first_pos: int = 100,
                             n_workers: int = 1) -> str:
    """``data.json`` / ``data.index`` per run and ``config.yml`` holding exactly the reads of a run-grouped ``batch``
    (``make_batch(..., keys=True)`` of a spec without pooling): the fast producer for the CLI benchmark - the text of
    10^5 positions in seconds, where :func:`write_dataset` formats every float in Python.  Every read must be a
    non-negative 2-decimal value below 655.36 (``dataprep.py:638`` rounds to 2 decimals); numbers are written as
    ``json.dump`` / ``repr`` would write them."""
    km = km or KmerModel.load()
    assert not batch.layout.pooling and batch.ids and batch.positions and batch.kmers
    info = spec.data_info(root)
    q = np.rint(batch.y[:batch.n_reads] * 100.0).astype(np.int64)
    assert q.min() >= 0 and q.max() < 65536 and np.array_equal(q / 100.0, batch.y[:batch.n_reads])
    text = np.empty(65536, dtype=object)
    for code in np.unique(q):
        text[int(code)] = repr(int(code) / 100.0)
    words = text[q]  # one python str per read
    G = batch.layout.n_groups
    starts = batch.read_off[:-1, None] + np.concatenate([np.zeros((batch.n_positions, 1), np.int64),
                                                         np.cumsum(batch.grp_len, axis=1)[:, :-1]], axis=1)  # [P, G]
    gene_first = [0] + [p for p in range(1, batch.n_positions) if batch.ids[p] != batch.ids[p - 1]] + [batch.n_positions]
    runs = [(c, run) for c, rd in info.items() for run in rd]
    assert len(runs) == G

    def write_run(g):
        c, run = runs[g]
        d = info[c][run]
        os.makedirs(d, exist_ok=True)
        with open(os.path.join(d, "data.json"), "w") as fj, open(os.path.join(d, "data.index"), "w") as fi:
            fi.write("idx,start,end\n")
            off = 0
            for a, b in zip(gene_first[:-1], gene_first[1:]):
                idx = batch.ids[a]
                parts = ['{"%s":{' % idx]
                for p in range(a, b):
                    s0 = int(starts[p, g])
                    parts.append('%s"%s":{"%s":[%s]}' % ("," if p > a else "", batch.positions[p], batch.kmers[p],
                                                          ",".join(words[s0:s0 + int(batch.grp_len[p, g])])))
                parts.append("}}\n")
                line = "".join(parts)
                fj.write(line)
                fi.write("%s,%d,%d\n" % (idx, off, off + len(line)))
                off += len(line)

    # one forked writer per run (the arrays are shared copy-on-write); in-process when fork is not an option
    workers = min(G, int(n_workers) if n_workers else 1)
    if workers > 1:
        import multiprocessing as mp
        procs = [mp.get_context("fork").Process(target=write_run, args=(g,)) for g in range(G)]
        live = []
        for pr in procs:
            pr.start()
            live.append(pr)
            if len(live) >= workers:
                live.pop(0).join()
        for pr in live:
            pr.join()
        assert all(pr.exitcode == 0 for pr in procs), [pr.exitcode for pr in procs]
    else:
        for g in range(G):
            write_run(g)
    cfg = {"data": {c: {run.split("-", 1)[1]: d for run, d in rd.items()} for c, rd in info.items()},
           "out": os.path.join(root, "out")}
    if spec.method:
        cfg["method"] = dict(spec.method)
    if spec.criteria:
        cfg["criteria"] = dict(spec.criteria)
    path = os.path.join(root, "config.yml")
    with open(path, "w") as f:
        yaml.safe_dump(cfg, f, sort_keys=False)
    return path


def write_dataset_from_batch(root: str, spec: Spec, batch: Batch, km: KmerModel | None = None, first_pos: int = 100,
                             n_workers: int = 1) -> str:
    """``data.json`` / ``data.index`` per run and ``config.yml`` holding exactly the reads of a run-grouped ``batch``
    (``make_batch(..., keys=True)`` of a spec without pooling): the fast producer for the CLI benchmark - the text of
    10^5 positions in seconds, where :func:`write_dataset` formats every float in Python.  Every read must be a
    non-negative 2-decimal value below 655.36 (``dataprep.py:638`` rounds to 2 decimals); numbers are written as
    ``json.dump`` / ``repr`` would write them."""
    km = km or KmerModel.load()
    assert not batch.layout.pooling and batch.ids and batch.positions and batch.kmers
    info = spec.data_info(root)
    q = np.rint(batch.y[:batch.n_reads] * 100.0).astype(np.int64)
    assert q.min() >= 0 and q.max() < 65536 and np.array_equal(q / 100.0, batch.y[:batch.n_reads])
    text = np.empty(65536, dtype=object)
    for code in np.unique(q):
        text[int(code)] = repr(int(code) / 100.0)
    words = text[q]  # one python str per read
    G = batch.layout.n_groups
    starts = batch.read_off[:-1, None] + np.concatenate([np.zeros((batch.n_positions, 1), np.int64),
                                                         np.cumsum(batch.grp_len, axis=1)[:, :-1]], axis=1)  # [P, G]
    gene_first = [0] + [p for p in range(1, batch.n_positions) if batch.ids[p] != batch.ids[p - 1]] + [batch.n_positions]
    runs = [(c, run) for c, rd in info.items() for run in rd]
    assert len(runs) == G
    for g, (c, run) in enumerate(runs):
        d = info[c][run]
        os.makedirs(d, exist_ok=True)
        with open(os.path.join(d, "data.json"), "w") as fj, open(os.path.join(d, "data.index"), "w") as fi:
            fi.write("idx,start,end\n")
            off = 0
            for a, b in zip(gene_first[:-1], gene_first[1:]):
                idx = batch.ids[a]
                parts = ['{"%s":{' % idx]
                for p in range(a, b):
                    s0 = int(starts[p, g])
                    parts.append('%s"%s":{"%s":[%s]}' % ("," if p > a else "", batch.positions[p], batch.kmers[p],
                                                          ",".join(words[s0:s0 + int(batch.grp_len[p, g])])))
                parts.append("}}\n")
                line = "".join(parts)
                fj.write(line)
                fi.write("%s,%d,%d\n" % (idx, off, off + len(line)))
                off += len(line)
    cfg = {"data": {c: {run.split("-", 1)[1]: d for run, d in rd.items()} for c, rd in info.items()},
           "out": os.path.join(root, "out")}
    if spec.method:
        cfg["method"] = dict(spec.method)
    if spec.criteria:
        cfg["criteria"] = dict(spec.criteria)
    path = os.path.join(root, "config.yml")
    with open(path, "w") as f:
        yaml.safe_dump(cfg, f, sort_keys=False)
    return path


def make_batch(spec: Spec, km: KmerModel | None = None, n_positions: int | None = None,
               chunk: int = 50000, seed_offset: int = 0, keys: bool = False) -> Batch:
    """The dataset of ``spec`` as a CSR batch, drawn chunk by chunk (vectorised).

    The read-count criteria of the config are applied like ``io.py:51-52,70-80`` so that the
    batch holds exactly the positions the packer would keep.  ``seed_offset`` decorrelates the
    shards of different ranks."""
    km = km or KmerModel.load()
    rng = np.random.default_rng([spec.seed, seed_offset])
    pooling = bool(spec.method.get("pooling", False))
    info = spec.data_info()
    lay_runs = Layout.from_data_info(info, pooling=False)
    lay = Layout.from_data_info(info, pooling=pooling)
    crit = spec.criteria or {"readcount_min": 15, "readcount_max": 1000}
    lo, hi = crit["readcount_min"], crit["readcount_max"]
    P_all = n_positions if n_positions is not None else spec.n_genes * spec.pos_per_gene
    run_cond = np.asarray(lay_runs.run_cond)
    C = lay.n_conditions
    ys, glens, kms, kss, krows = [], [], [], [], []
    for c0 in range(0, P_all, chunk):
        P = min(chunk, P_all - c0)
        _, krow, mu, sd, shift, counts, rate = _draw(spec, rng, km, P)
        if pooling:
            csum = np.stack([counts[:, run_cond == c].sum(axis=1) for c in range(C)], axis=1)
            keep = ((csum >= lo) & (csum <= hi)).sum(axis=1) >= 2
        else:
            ok = (counts >= lo) & (counts <= hi)
            counts = np.where(ok, counts, 0)
            has = np.stack([ok[:, run_cond == c].any(axis=1) for c in range(C)], axis=1)
            keep = has.sum(axis=1) >= 2
        krow, mu, sd, shift, counts, rate = (a[keep] for a in (krow, mu, sd, shift, counts, rate))
        P = len(krow)
        n_pr = counts.reshape(-1)  # [P*Gruns], run-major within position
        R = int(n_pr.sum())
        cell = np.repeat(np.arange(P * counts.shape[1]), n_pr)
        pos = cell // counts.shape[1]
        mod = rng.random(R) < rate.reshape(-1)[cell]
        y = rng.normal(mu[pos], sd[pos])
        ym = rng.normal(mu[pos] + shift[pos], 1.3 * sd[pos])
        y = np.round(np.where(mod, ym, y), 2)
        if pooling:
            # reads must be contiguous per condition slot (sorted condition order)
            order = np.argsort(run_cond, kind="stable")
            if not (order == np.arange(len(order))).all():
                cell_rank = np.empty_like(order)
                cell_rank[order] = np.arange(len(order))
                key = pos * counts.shape[1] + cell_rank[cell % counts.shape[1]]
                y = y[np.argsort(key, kind="stable")]
            glen = np.stack([counts[:, run_cond == c].sum(axis=1) for c in range(C)], axis=1)
        else:
            glen = counts
        ys.append(y)
        glens.append(glen.astype(np.int32))
        kms.append(mu)
        kss.append(sd)
        krows.append(krow)
    grp_len = np.concatenate(glens)
    read_off = np.zeros(len(grp_len) + 1, dtype=np.int64)
    np.cumsum(grp_len.sum(axis=1, dtype=np.int64), out=read_off[1:])
    b = Batch(lay, np.concatenate(ys), read_off, np.ascontiguousarray(grp_len),
              np.concatenate(kms), np.concatenate(kss))
    if keys:
        krow = np.concatenate(krows)
        b.ids = ["ENST%011d" % (i // spec.pos_per_gene + 1) for i in range(b.n_positions)]
        b.positions = [str(100 + i % spec.pos_per_gene) for i in range(b.n_positions)]
        b.kmers = [km.kmers[k] for k in krow]
    return b.validate()


def make_batch_device(spec: Spec, device, km: KmerModel | None = None, n_positions: int | None = None,
                      seed_offset: int = 0, chunk: int = 250000) -> Batch:
    """Same statistical model as :func:`make_batch`, drawn with torch on ``device`` (the numpy
    generator needs ~40 s for the 3.6e8 reads of config c2; this takes about a second) and
    returned as a host :class:`Batch`.  Not bit-identical to :func:`make_batch`."""
    import torch

    km = km or KmerModel.load()
    gen = torch.Generator(device=device)
    gen.manual_seed(1000003 * spec.seed + seed_offset)
    pooling = bool(spec.method.get("pooling", False))
    info = spec.data_info()
    lay_runs = Layout.from_data_info(info, pooling=False)
    lay = Layout.from_data_info(info, pooling=pooling)
    crit = spec.criteria or {"readcount_min": 15, "readcount_max": 1000}
    lo, hi = crit["readcount_min"], crit["readcount_max"]
    P_all = n_positions if n_positions is not None else spec.n_genes * spec.pos_per_gene
    Gr = lay_runs.n_groups
    run_cond = torch.tensor(lay_runs.run_cond, device=device)
    C = lay.n_conditions
    if pooling:
        assert lay_runs.run_cond == sorted(lay_runs.run_cond), "pooled synthetic data needs sorted conditions"
    rows = torch.tensor(km.usable_rows(), device=device)
    kmean = torch.tensor(km.mean, device=device)
    kstd = torch.tensor(km.stdv, device=device)
    ko = torch.tensor([lay_runs.condition_names[c] in spec.ko_like for c in lay_runs.run_cond], device=device)
    hundred = torch.tensor(100.0, dtype=torch.float64, device=device)
    ys, glens, kms, kss = [], [], [], []
    for c0 in range(0, P_all, chunk):
        P = min(chunk, P_all - c0)
        krow = rows[torch.randint(0, len(rows), (P,), device=device, generator=gen)]
        mu, sd = kmean[krow], kstd[krow]
        is_mod = torch.rand(P, device=device, generator=gen) < spec.f_mod
        shift = (3.0 + 5.0 * torch.rand(P, device=device, generator=gen, dtype=torch.float64)) * \
            torch.where(torch.rand(P, device=device, generator=gen) < 0.5, -1.0, 1.0)
        if spec.loguniform_total is not None:
            l0, l1 = spec.loguniform_total
            tot = torch.exp(math.log(l0) + (math.log(l1) - math.log(l0)) *
                            torch.rand(P, device=device, generator=gen, dtype=torch.float64))
            counts = torch.clamp(torch.round(tot / Gr), min=1).to(torch.int64)[:, None].repeat(1, Gr)
        else:
            counts = torch.randint(spec.reads_lo, spec.reads_hi + 1, (P, Gr), device=device, generator=gen)
        rate = torch.where(ko, 0.05, 0.6)[None, :] * is_mod[:, None]
        if pooling:
            csum = torch.stack([counts[:, run_cond == c].sum(dim=1) for c in range(C)], dim=1)
            keep = ((csum >= lo) & (csum <= hi)).sum(dim=1) >= 2
        else:
            ok = (counts >= lo) & (counts <= hi)
            counts = torch.where(ok, counts, 0)
            has = torch.stack([ok[:, run_cond == c].any(dim=1) for c in range(C)], dim=1)
            keep = has.sum(dim=1) >= 2
        mu, sd, shift, counts, rate = mu[keep], sd[keep], shift[keep], counts[keep], rate[keep]
        n_cell = counts.reshape(-1)
        R = int(n_cell.sum())
        cell = torch.repeat_interleave(torch.arange(n_cell.numel(), device=device, dtype=torch.int32), n_cell,
                                       output_size=R)
        pos = (cell // Gr).long()
        mod = torch.rand(R, device=device, generator=gen) < rate.reshape(-1)[cell.long()]
        z = torch.randn(R, device=device, generator=gen, dtype=torch.float64)
        y = torch.where(mod, mu[pos] + shift[pos] + 1.3 * sd[pos] * z, mu[pos] + sd[pos] * z)
        # true division by a device tensor: torch turns `/ python_scalar` into a multiplication by the
        # reciprocal on CUDA, which would not give the correctly rounded 2-decimal doubles
        y = torch.round(y * 100.0) / hundred
        glen = torch.stack([counts[:, run_cond == c].sum(dim=1) for c in range(C)], dim=1) if pooling else counts
        ys.append(y.cpu())
        glens.append(glen.to(torch.int32).cpu())
        kms.append(mu.cpu())
        kss.append(sd.cpu())
        del cell, pos, mod, z, y
    grp_len = torch.cat(glens).numpy()
    read_off = np.zeros(len(grp_len) + 1, dtype=np.int64)
    np.cumsum(grp_len.sum(axis=1, dtype=np.int64), out=read_off[1:])
    return Batch(lay, torch.cat(ys).numpy(), read_off, np.ascontiguousarray(grp_len), torch.cat(kms).numpy(),
                 torch.cat(kss).numpy()).validate()
