"""Shared test plumbing: open a golden case the way the diffmod driver does."""
import json
import os
from collections import OrderedDict

from xpore_b200.config import Configurator


def open_case(config_path):
    cfg = Configurator(config_path)
    data_info = cfg.get_data_info()
    method = cfg.get_method()
    criteria = cfg.get_criteria()
    index, handles = {}, {}
    for cond, runs in data_info.items():
        for run, d in runs.items():
            idx = OrderedDict()
            with open(os.path.join(d, "data.index")) as f:
                next(f)
                for line in f:
                    k, s, e = line.rstrip("\n").rsplit(",", 2)
                    idx[k] = (int(s), int(e))
            index[run] = idx
            handles[run] = open(os.path.join(d, "data.json"))
    return cfg, data_info, method, criteria, index, handles


def gene_dict(idx, data_info, index, handles):
    dd = OrderedDict()
    for cond, runs in data_info.items():
        for run in runs:
            if idx in index[run]:
                s, e = index[run][idx]
                handles[run].seek(s)
                dd[(cond, run)] = json.loads(handles[run].read(e - s))
            else:
                dd[(cond, run)] = None
    return dd


# ---------------------------------------------------------------------------------------------
# packing a golden case into a CSR batch and comparing result records with the reference
# ---------------------------------------------------------------------------------------------
def pack_case(config_path, ids):
    """(batch, method dict, data_info) for a golden case via the product packer."""
    from xpore_b200.batch import BatchBuilder, Layout
    from xpore_b200.kmer_model import KmerModel

    cfg, data_info, method, criteria, index, handles = open_case(config_path)
    lay = Layout.from_data_info(data_info, method["pooling"])
    bb = BatchBuilder(lay, KmerModel.load(), criteria["readcount_min"], criteria["readcount_max"])
    for idx in ids:
        bb.add_gene(idx, gene_dict(idx, data_info, index, handles))
    for h in handles.values():
        h.close()
    return bb.build(), method, data_info


def _num(v):
    return float(v) if isinstance(v, str) else v


def compare_with_golden(batch, res, method, expected, rtol_param=1e-5, rtol_p=1e-4, iters_exact=True):
    """Assert a FitResult agrees with the reference outputs of a golden case:
    selection exact; fitted set exact; parameters rel 1e-5; p-values rel 1e-4; rows exact set."""
    import numpy as np

    from xpore_b200 import fit as xf
    from xpore_b200.table import get_result_table_header, result_rows

    lay = batch.layout
    keys = list(zip(batch.ids, batch.positions, batch.kmers))
    # selection (io.py:21-90)
    exp_sel = {tuple(s["key"]): s for s in expected["selected"]}
    assert set(keys) == set(exp_sel)
    n = np.diff(batch.read_off)
    for i, k in enumerate(keys):
        assert n[i] == exp_sel[k]["n_reads"]
        present = [lay.group_names[g] for g in range(lay.n_groups) if batch.grp_len[i, g] > 0]
        assert sorted(present) == sorted(exp_sel[k]["group_names"])
    # fitted set (diffmod.py:52-58)
    exp_fit = {tuple(p["key"]): p for p in expected["positions"]}
    got_fit = {keys[i] for i in np.nonzero(res.selected)[0]}
    assert got_fit == set(exp_fit), (sorted(got_fit ^ set(exp_fit)))
    assert res.n_fitted == len(exp_fit)
    iters_diff = 0
    for i, k in enumerate(keys):
        if k not in exp_fit:
            continue
        e = exp_fit[k]
        # the reference orders model groups by sorted name; ours are layout slots
        gmap = [lay.group_names.index(nm) for nm in e["group_names"]]
        if iters_exact:
            assert res.n_iter[i] == e["n_iterations"], (k, res.n_iter[i], e["n_iterations"])
        iters_diff += int(res.n_iter[i] != e["n_iterations"])
        assert bool(res.status[i] & xf.CONVERGED) == e["converged"], k
        np.testing.assert_allclose(res.mu[i], [_num(v) for v in e["location"]], rtol=rtol_param, err_msg=str(k))
        np.testing.assert_allclose(res.lam[i], [_num(v) for v in e["lambda"]], rtol=rtol_param, err_msg=str(k))
        np.testing.assert_allclose(res.alpha[i], [_num(v) for v in e["alpha"]], rtol=rtol_param, err_msg=str(k))
        np.testing.assert_allclose(res.beta[i], [_num(v) for v in e["beta"]], rtol=rtol_param, err_msg=str(k))
        N_ref = np.array([[_num(v) for v in r] for r in e["N"]])
        np.testing.assert_allclose(res.N[i][gmap], N_ref, rtol=rtol_param, atol=1e-9, err_msg=str(k))
        if e["prefilter"] is not None:
            pv = _num(list(e["prefilter"].values())[0])
            if np.isnan(pv):
                assert np.isnan(res.pval[i])
            else:
                np.testing.assert_allclose(res.pval[i], pv, rtol=rtol_p, err_msg=str(k))
    # emitted rows (io.py:346-367)
    pre = method["prefiltering"]
    pm = pre["method"] if pre else None
    assert get_result_table_header(lay, pm) == expected["header"]
    rows = {tuple(r[:3]): r for r in result_rows(batch, res, pm)}
    exp_rows = {tuple(r[:3]): r for r in expected["rows"]}
    assert set(rows) == set(exp_rows)
    hdr = expected["header"]
    knife, bound = 0, 0
    index_of = {k: i for i, k in enumerate(keys)}
    for k, r in rows.items():
        # z-test columns: where the reference's n = round(mean coverage) sits on x.5 its rounding direction is sub-ulp
        # noise (see xp_core.cuh cond_summary) - the other direction is accepted there, counted, and bounded by the
        # number of cells for which that can legitimately happen
        knife += compare_row_cells(k, hdr, r, exp_rows[k], lay, pm, rtol_param, rtol_p)
        bound += half_integer_tests(batch, index_of[k])
    assert knife <= bound, (knife, bound)
    print("knife-edge z-test cells accepted: %d (cells on a half-integer mean read count: %d, rows: %d)" % (knife, bound, len(rows)))
    return iters_diff, knife


def _knife_edge_alternatives(hdr, row, lay):
    """For every z-test triple of a reference row: the values the reference's own formula
    (stats.py:7-15) gives from the row's mod_rate_*/coverage_* columns when a mean coverage that
    lies within 1e-6 of x.5 is rounded the other way."""
    import math
    from itertools import combinations, product

    import numpy as np
    import scipy.stats

    col = {h: (None if v is None else _num(v)) for h, v in zip(hdr[3:], row[3:]) if h != "mod_assignment"}
    names = lay.group_names
    gc = list(lay.grp_cond)
    present = [g for g, nm in enumerate(names) if col.get("mod_rate_%s" % nm) is not None]

    def summary(groups):
        w = np.array([col["mod_rate_%s" % names[g]] for g in groups])
        c = np.array([col["coverage_%s" % names[g]] for g in groups])
        m = c.mean()
        ns = [np.round(m)]
        if abs(m - math.floor(m) - 0.5) < 1e-6:
            ns = [math.floor(m), math.ceil(m)]
        return w.mean(), ns

    def triples(g1, g2):
        p1, n1s = summary(g1)
        p2, n2s = summary(g2)
        out = []
        for n1, n2 in product(n1s, n2s):
            se = np.sqrt(p1 * (1 - p1) / n1 + p2 * (1 - p2) / n2)
            z = (p1 - p2) / se
            out.append((p1 - p2, scipy.stats.norm.sf(abs(z)) * 2, z))
        return out

    alt = {}
    conds = lay.condition_names
    tests = [("%s_vs_%s" % (conds[a], conds[b]), [g for g in present if gc[g] == a], [g for g in present if gc[g] == b])
             for a, b in combinations(range(len(conds)), 2)]
    if len(conds) > 2:
        tests += [("%s_vs_all" % conds[a], [g for g in present if gc[g] == a], [g for g in present if gc[g] != a])
                  for a in range(len(conds))]
    for tag, g1, g2 in tests:
        if not g1 or not g2:
            continue
        tr = triples(g1, g2)
        alt["diff_mod_rate_" + tag] = [t[0] for t in tr]
        alt["pval_" + tag] = [t[1] for t in tr]
        alt["z_score_" + tag] = [t[2] for t in tr]
    return alt


# ---------------------------------------------------------------------------------------------
# host emulator of the kernel arithmetic (tests/emul/xp_emul.cpp) - checker only
# ---------------------------------------------------------------------------------------------
_EMUL = None


def load_emulator():
    import ctypes as C
    import subprocess
    global _EMUL
    if _EMUL is not None:
        return _EMUL
    here = os.path.dirname(os.path.abspath(__file__))
    src = os.path.join(here, "emul", "xp_emul.cpp")
    so = os.path.join(here, "emul", "libxp_emul.so")
    deps = [src] + [os.path.join(here, "..", "xpore_b200", "csrc", f) for f in ("xp_math.cuh", "xp_core.cuh")]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-shared", "-fPIC", "-x", "c++",
                               src, "-o", so])
    L = C.CDLL(so)
    L.emul_prefilter.restype = C.c_double
    L.emul_prefilter.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_double]
    L.emul_fit.restype = None
    L.emul_fit.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_double, C.c_double, C.c_int, C.c_double,
                           C.c_double, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.emul_stats.restype = C.c_uint
    L.emul_stats.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_double,
                             C.c_void_p, C.c_void_p]
    L.emul_special.restype = None
    L.emul_special.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong]
    _EMUL = L
    return L


def emul_fit_batch(L, batch, method):
    """Run the emulator position by position and fill the same records the C ABI returns."""
    import ctypes as C

    import numpy as np

    from xpore_b200 import fit as xf
    P, G, Cn = batch.n_positions, batch.layout.n_groups, batch.layout.n_conditions
    res = xf.FitResult(G, Cn, np.zeros(P, np.uint32), np.zeros(P, np.int32), np.full(P, np.nan),
                       np.full((P, xf.fit_width(G)), np.nan), np.full((P, xf.stats_width(G, Cn)), np.nan), 0)
    gc = np.ascontiguousarray(batch.layout.grp_cond, dtype=np.int32)
    for p in range(P):
        y = np.ascontiguousarray(batch.y[batch.read_off[p]:batch.read_off[p + 1]])
        gl = np.ascontiguousarray(batch.grp_len[p])
        sel = True
        if method.prefilter_method is not None:
            pv = np.nan
            if method.prefilter_method == "t-test":
                pv = L.emul_prefilter(len(y), y.ctypes.data, G, gl.ctypes.data, gc.ctypes.data, batch.kmer_mean[p])
            res.pval[p] = pv
            sel = bool(np.isnan(pv) or pv < method.prefilter_threshold)
        if not sel:
            continue
        nit, st = C.c_int(0), C.c_uint(0)
        L.emul_fit(len(y), y.ctypes.data, G, gl.ctypes.data, batch.kmer_mean[p], batch.kmer_std[p],
                   method.max_iters, method.stopping_criteria, method.dirichlet_conc, res.fit[p].ctypes.data,
                   C.byref(nit), C.byref(st), None)
        fl = L.emul_stats(G, Cn, gc.ctypes.data, gl.ctypes.data, method.dirichlet_conc, batch.kmer_mean[p],
                          batch.kmer_std[p], res.fit[p].ctypes.data, res.stats[p].ctypes.data)
        res.n_iter[p] = nit.value
        res.status[p] = st.value | fl
        res.n_fitted += 1
    return res


def read_table(path):
    """diffmod.table -> (header, {key: row}) with numeric cells parsed ('' -> None)."""
    import csv
    with open(path, newline="") as f:
        rd = list(csv.reader(f))
    header = rd[0]
    rows = {}
    for r in rd[1:]:
        vals = list(r[:3])
        for name, v in zip(header[3:], r[3:]):
            if name == "mod_assignment":
                vals.append(v)
            else:
                vals.append(None if v == "" else float(v))
        rows[tuple(r[:3])] = vals
    return header, rows


def compare_row_cells(key, header, got, exp, layout, prefilter_method, rtol_param=1e-5, rtol_p=1e-4):
    """One table row against the reference's / the oracle's: north-star tolerances cell by cell.  Returns the
    number of z-test cells that only agree with the OTHER rounding of a mean coverage lying on x.5 (SURVEY.md Q8:
    there the reference's own direction is decided by sub-ulp summation noise)."""
    import numpy as np
    assert len(got) == len(exp) == len(header), (key, len(got), len(exp), len(header))
    alt = None
    knife = 0
    for name, a, b in zip(header[3:], got[3:], exp[3:]):
        if b is None or name == "mod_assignment":
            assert a == b, (key, name, a, b)
            continue
        b = _num(b)
        assert a is not None, (key, name, b)
        if np.isnan(b):
            assert np.isnan(a), (key, name, a)
            continue
        tol = rtol_p if (name.startswith("pval") or name == prefilter_method) else rtol_param
        if np.isclose(a, b, rtol=tol, atol=1e-300):
            continue
        if alt is None:
            alt = _knife_edge_alternatives(header, exp, layout)
        assert any(np.isclose(a, c, rtol=tol, atol=1e-300) for c in alt.get(name, [])), (key, name, a, b, alt.get(name))
        knife += 1
    return knife


def half_integer_tests(batch, p):
    """Number of z-test cells of position p for which a knife-edge acceptance is legitimate: tests with a side whose
    mean EXACT read count over its present groups is a half-integer (3 cells - diff, pval, z - per such test; the
    diff_mod_rate cell never depends on n, so at most 2 can differ)."""
    import numpy as np
    lay = batch.layout
    gc = np.asarray(lay.grp_cond)
    gl = batch.grp_len[p]
    Cn = lay.n_conditions

    def half(groups):
        g = [x for x in groups if gl[x] > 0]
        return bool(g) and (2 * int(gl[g].sum())) % (2 * len(g)) == len(g)

    sides = [[g for g in range(lay.n_groups) if gc[g] == c] for c in range(Cn)]
    rest = [[g for g in range(lay.n_groups) if gc[g] != c] for c in range(Cn)]
    n = 0
    for a in range(Cn):
        for b in range(a + 1, Cn):
            n += 2 * int(half(sides[a]) or half(sides[b]))
    if Cn > 2:
        for a in range(Cn):
            n += 2 * int(half(sides[a]) or half(rest[a]))
    return n


def compare_tables(got_rows, exp_rows, header, layout, prefilter_method, rtol_param=1e-5, rtol_p=1e-4):
    assert set(got_rows) == set(exp_rows)
    knife = 0
    for k, r in got_rows.items():
        knife += compare_row_cells(k, header, r, exp_rows[k], layout, prefilter_method, rtol_param, rtol_p)
    return knife


def compare_stats_with_oracle(batch, res, p, fit, present, data_info, prefilter_method, table_row, header):
    """EVERY statistic of fitted position p (kept by the row filter or not) against the oracle's result_row() evaluated on
    `fit` (posterior parameters: the oracle's own fit, or the reference's golden parameters): the table cells
    (z-test triples, mod rates, coverages, mu / sigma2 / conf, assignment), the overlap of the two Gaussians, the
    four cdfs and the keep decision (io.py:255-367).  Returns the knife-edge count of compare_row_cells."""
    import numpy as np

    from oracle import xpore_oracle as orc
    from xpore_b200 import fit as xf
    lay = batch.layout
    names = [lay.group_names[g] for g in present]
    pv = None
    if prefilter_method:
        pv = float(res.pval[p])
    key = (table_row[0], table_row[1], table_row[2])
    row, keep, extra = orc.result_row(key, fit, names, batch.kmer_mean[p], batch.kmer_std[p], data_info, lay.pooling, pv)
    knife = compare_row_cells(key, header, table_row, row, lay, prefilter_method)
    st = res.stats[p]
    if np.isnan(extra["p_overlap"]):
        assert np.isnan(st[xf.ST_POV]), key
        assert not keep
    else:
        # 1 - (|F_Y(x1) - F_X(x1)| + |F_Y(x2) - F_X(x2)|): differences of cdfs, absolute accuracy is what exists
        np.testing.assert_allclose(st[xf.ST_POV], extra["p_overlap"], rtol=1e-5, atol=1e-9, err_msg=str(key))
        np.testing.assert_allclose(st[xf.ST_CDF:xf.ST_CDF + 4], extra["cdfs"], rtol=1e-5, atol=1e-9, err_msg=str(key))
    # keep decision: identical unless the reference's own inputs sit on a threshold of the filter (never seen)
    got_keep = bool(res.status[p] & xf.KEEP_ROW)
    if got_keep != keep:
        c = np.asarray(extra["cdfs"])
        near = abs(extra["p_overlap"] - 0.5) < 1e-6 or (np.abs(c - 0.1) < 1e-6).any() or (np.abs(c - 0.9) < 1e-6).any()
        assert near, (key, got_keep, keep, extra["p_overlap"], extra["cdfs"])
    assert bool(res.status[p] & xf.MOD_IS_1) == (extra["mod"] == 1), key
    return knife


def assert_same_rows(rows1, rows2):
    """Two parsed tables hold the very same rows (NaN cells - e.g. the prefilter column of a position with more than two
    conditions - compare equal)."""
    assert set(rows1) == set(rows2)
    for k, a in rows1.items():
        b = rows2[k]
        assert len(a) == len(b), k
        for x, y in zip(a, b):
            assert x == y or (isinstance(x, float) and isinstance(y, float) and x != x and y != y), (k, x, y)
