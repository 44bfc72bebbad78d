"""``diffmod.table`` layout: header and rows from the result records.

Mirrors ``xpore/diffmod/io.py:189-228`` (``get_result_table_header``) and the row assembly of
``generate_result_table`` (``io.py:346-367``).  All numbers were computed on the device
(``xp_stats_kernel``); this module only orders columns, decides which cells are empty
(``None`` in the reference -> empty CSV field) and applies the row filter flag.
"""
from __future__ import annotations

import csv
from itertools import combinations

import numpy as np

from .batch import Batch, Layout
from .fit import HIGHER, KEEP_ROW, SELECTED, ST_CONF, ST_MU, ST_RATE, ST_S2, FitResult


def get_result_table_header(layout: Layout, prefilter_method: str | None):
    conds = layout.condition_names
    header = ["id", "position", "kmer"]
    for a, b in combinations(conds, 2):
        pair = "_vs_".join((a, b))
        header += ["diff_mod_rate_%s" % pair, "pval_%s" % pair, "z_score_%s" % pair]
    if len(conds) > 2:
        for c in conds:
            pair = "%s_vs_all" % c
            header += ["diff_mod_rate_%s" % pair, "pval_%s" % pair, "z_score_%s" % pair]
    names = layout.group_names
    header += ["mod_rate_%s" % n for n in names]
    header += ["coverage_%s" % n for n in names]
    header += ["mu_unmod", "mu_mod", "sigma2_unmod", "sigma2_mod", "conf_mu_unmod", "conf_mu_mod", "mod_assignment"]
    if prefilter_method:
        header += [prefilter_method]
    return header


def _cell(v):
    """csv.writer prints floats with repr() and None as '' (diffmod.py:68-69)."""
    if v is None:
        return ""
    if isinstance(v, str):
        return v
    return repr(float(v))


def _row(key, st, status, pval, present, cond_present, layout: Layout, pairs, prefilter_method):
    """One table row (python values, ``None`` for empty cells) from a statistics record (io.py:346-361)."""
    G, Cn = layout.n_groups, layout.n_conditions
    o_pair = ST_RATE + 2 * G
    o_ova = o_pair + 3 * len(pairs)
    row = [key[0], key[1], key[2]]
    for j, (a, b) in enumerate(pairs):
        if cond_present[a] and cond_present[b]:
            row += [float(v) for v in st[o_pair + 3 * j:o_pair + 3 * j + 3]]
        else:
            row += [None, None, None]
    if Cn > 2:
        for c in range(Cn):
            if cond_present[c]:
                row += [float(v) for v in st[o_ova + 3 * c:o_ova + 3 * c + 3]]
            else:
                row += [None, None, None]
    row += [float(st[ST_RATE + g]) if present[g] else None for g in range(G)]
    row += [float(st[ST_RATE + G + g]) if present[g] else None for g in range(G)]
    row += [float(v) for v in st[ST_MU:ST_MU + 2]] + [float(v) for v in st[ST_S2:ST_S2 + 2]] + \
           [float(v) for v in st[ST_CONF:ST_CONF + 2]]
    row += ["higher" if (int(status) & HIGHER) else "lower"]
    if prefilter_method:
        row += [float(pval)]
    return row


def result_rows(batch: Batch, res: FitResult, prefilter_method: str | None, only_kept: bool = True):
    """Rows (lists of python values, ``None`` for empty cells) of the positions that pass the
    row filter — or of every fitted position with ``only_kept=False`` (the reference can only
    produce the filtered table)."""
    lay = batch.layout
    Cn = lay.n_conditions
    gc = lay.grp_cond
    present = batch.grp_len > 0  # [P, G]
    cond_present = np.stack([present[:, gc == c].any(axis=1) for c in range(Cn)], axis=1)  # [P, C]
    pairs = list(combinations(range(Cn), 2))
    mask = res.keep_row if only_kept else res.selected
    return [_row((batch.ids[p], batch.positions[p], batch.kmers[p]), res.stats[p], res.status[p], res.pval[p], present[p],
                 cond_present[p], lay, pairs, prefilter_method) for p in np.nonzero(mask)[0]]


def rows_from_records(records: np.ndarray, keys, present: np.ndarray, layout: Layout, prefilter_method: str | None):
    """Rows from the compact records of ``xpb200_fit_host_rows`` (``[position, status, n_iter, pval, fit, stats]``;
    every record is a row of the table): ``keys[i]`` = (id, position, kmer) and ``present[i]`` = the groups present
    at the position of record i."""
    from .fit import fit_width
    G, Cn = layout.n_groups, layout.n_conditions
    gc = layout.grp_cond
    o_st = 4 + fit_width(G)
    pairs = list(combinations(range(Cn), 2))
    present = np.asarray(present, dtype=bool).reshape(len(records), G)
    cond_present = np.stack([present[:, gc == c].any(axis=1) for c in range(Cn)], axis=1) if len(records) else present
    return [_row(keys[i], rec[o_st:], int(rec[1]), rec[3], present[i], cond_present[i], layout, pairs, prefilter_method)
            for i, rec in enumerate(records)]


def write_rows(path: str, rows, mode: str = "a"):
    with open(path, mode, newline="") as f:
        write_rows_to(f, rows)


def write_rows_to(f, rows):
    w = csv.writer(f, delimiter=",")  # default dialect (CRLF), as the reference writes it
    for row in rows:
        w.writerow([_cell(v) for v in row])
