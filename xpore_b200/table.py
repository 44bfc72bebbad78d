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


def result_rows(batch: Batch, res: FitResult, prefilter_method: str | None, only_kept: bool = True):
    """Rows (lists of python values, ``None`` for empty cells) of the positions that pass the
    row filter — or of every fitted position with ``only_kept=False`` (the reference can only
    produce the filtered table)."""
    lay = batch.layout
    G, Cn = lay.n_groups, lay.n_conditions
    gc = lay.grp_cond
    present = batch.grp_len > 0  # [P, G]
    cond_present = np.stack([present[:, gc == c].any(axis=1) for c in range(Cn)], axis=1)  # [P, C]
    pairs = list(combinations(range(Cn), 2))
    mask = res.keep_row if only_kept else res.selected
    rows = []
    st = res.stats
    o_pair = ST_RATE + 2 * G
    o_ova = o_pair + 3 * len(pairs)
    for p in np.nonzero(mask)[0]:
        row = [batch.ids[p], batch.positions[p], batch.kmers[p]]
        for j, (a, b) in enumerate(pairs):
            if cond_present[p, a] and cond_present[p, b]:
                row += [float(v) for v in st[p, o_pair + 3 * j:o_pair + 3 * j + 3]]
            else:
                row += [None, None, None]
        if Cn > 2:
            for c in range(Cn):
                if cond_present[p, c]:
                    row += [float(v) for v in st[p, o_ova + 3 * c:o_ova + 3 * c + 3]]
                else:
                    row += [None, None, None]
        row += [float(st[p, ST_RATE + g]) if present[p, g] else None for g in range(G)]
        row += [float(st[p, ST_RATE + G + g]) if present[p, g] else None for g in range(G)]
        row += [float(v) for v in st[p, ST_MU:ST_MU + 2]] + [float(v) for v in st[p, ST_S2:ST_S2 + 2]] + \
               [float(v) for v in st[p, ST_CONF:ST_CONF + 2]]
        row += ["higher" if (res.status[p] & HIGHER) else "lower"]
        if prefilter_method:
            row += [float(res.pval[p])]
        rows.append(row)
    return rows


def write_rows(path: str, rows, mode: str = "a"):
    with open(path, mode, newline="") as f:
        w = csv.writer(f, delimiter=",")  # default dialect (CRLF), as the reference writes it
        for row in rows:
            w.writerow([_cell(v) for v in row])
