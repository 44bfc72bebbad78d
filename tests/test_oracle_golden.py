"""Pin the CPU oracle against outputs of the unmodified reference (tests/golden/*)."""
import numpy as np
import pytest

from oracle import xpore_oracle as orc
from xpore_b200.kmer_model import KmerModel

from conftest import fnum
from helpers import gene_dict, open_case


def _run(case):
    cfg, data_info, method, criteria, index, handles = open_case(case["config_path"])
    km = KmerModel.load()
    sel, fits, rows = {}, {}, []
    rng = np.random.RandomState(0)
    for idx in case["expected"]["ids"]:
        dd = gene_dict(idx, data_info, index, handles)
        d, f, r = orc.run_gene(idx, dd, data_info, method, criteria, km.signal, rng=rng)
        sel.update(d)
        fits.update(f)
        rows += r
    return data_info, method, sel, fits, rows


def test_oracle_matches_reference(golden_case):
    exp = golden_case["expected"]
    data_info, method, sel, fits, rows = _run(golden_case)

    # selection: exact key sets, read counts and group names (io.py:21-90)
    exp_sel = {tuple(s["key"]): s for s in exp["selected"]}
    assert set(sel.keys()) == set(exp_sel.keys())
    pooling = bool(method["pooling"])
    for k, d in sel.items():
        assert len(d["y"]) == exp_sel[k]["n_reads"]
        assert list(d["condition_names"] if pooling else d["run_names"]) == exp_sel[k]["group_names"]
        assert float(np.sum(d["y"])) == fnum(exp_sel[k]["y_sum"])

    # fitted set (prefilter gate, diffmod.py:52-58): exact
    exp_fit = {tuple(p["key"]): p for p in exp["positions"]}
    assert set(fits.keys()) == set(exp_fit.keys())

    # parameters: the restatement follows the reference's operation order => expect ~bit equality
    for k, f in fits.items():
        e = exp_fit[k]
        assert f["n_iterations"] == e["n_iterations"], k
        assert f["converged"] == e["converged"]
        for name in ("location", "lambda", "alpha", "beta"):
            np.testing.assert_allclose(f[name], [fnum(v) for v in e[name]], rtol=1e-12, atol=0, err_msg=str(k))
        np.testing.assert_allclose(f["conc"], np.array([[fnum(v) for v in r] for r in e["concentration"]]),
                                   rtol=1e-12, atol=1e-300)
        np.testing.assert_allclose(f["N"], np.array([[fnum(v) for v in r] for r in e["N"]]), rtol=1e-12, atol=1e-300)
        if e["prefilter"] is not None:
            pv = fnum(list(e["prefilter"].values())[0])
            assert (np.isnan(pv) and np.isnan(f["prefilter"])) or pv == f["prefilter"]
        if e["elbo_prev"] is not None and e["n_elbos"] > 1:
            # log_elbos[-1] is ELBO of the previous sweep (gmm.py:116); entry 0 is the random-z one
            assert f["n_elbos"] == e["n_elbos"]
            np.testing.assert_allclose(f["elbo_prev"], fnum(e["elbo_prev"]), rtol=1e-12)

    # emitted rows (post row-filter, io.py:363-367): same keys, same values
    exp_rows = {tuple(r[:3]): r for r in exp["rows"]}
    got_rows = {tuple(r[:3]): r for r in rows}
    assert set(got_rows) == set(exp_rows)
    pre = method["prefiltering"]
    assert orc.table_header(data_info, pooling, pre["method"] if pre else None) == exp["header"]
    for k, r in got_rows.items():
        e = exp_rows[k]
        assert len(r) == len(e) == len(exp["header"])
        for a, b in zip(r[3:], e[3:]):
            if b is None or isinstance(a, str):
                assert a == b
            else:
                b = fnum(b)
                assert (np.isnan(a) and np.isnan(b)) or a == pytest.approx(b, rel=1e-12, abs=1e-300)


@pytest.mark.parametrize("name", ["g1_runs", "g3_pooled"])
def test_oracle_model_nodes(name):
    """The node values `--save_models` dumps (io.py:99-139): responsibilities of the last E-step and the y-node
    statistics, oracle against the live reference (tests/golden/make_golden_models.py)."""
    import json
    import os

    from conftest import GOLDEN
    models = json.load(open(os.path.join(GOLDEN, name, "models.json")))
    assert len(models) >= 4
    for m in models:
        y = np.asarray(m["y_data"])
        g = np.asarray(m["x_group_of_read"])
        X = np.zeros((len(y), len(m["x_group_names"])), dtype=bool)
        X[np.arange(len(y)), g] = True
        from xpore_b200.kmer_model import KmerModel
        km = KmerModel.load()
        row = km.kmers.index(m["kmer"]) if isinstance(km.kmers, list) else list(km.kmers).index(m["kmer"])
        f = orc.fit_position(y, X, float(km.mean[row]), float(km.stdv[row]), rng=np.random.RandomState(0), trace=True)
        np.testing.assert_allclose(f["z_prob"], np.asarray(m["z_prob"]), rtol=1e-9, atol=1e-300)
        np.testing.assert_allclose(f["z_ln_prob"], np.asarray(m["z_ln_prob"]), rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(f["y_mean"], m["y_mean"], rtol=1e-12)
        np.testing.assert_allclose(f["y_variance"], m["y_variance"], rtol=1e-10)
        np.testing.assert_allclose(f["N"], np.asarray(m["y_N"]), rtol=1e-10, atol=1e-300)
        np.testing.assert_allclose(f["conc"], np.asarray(m["w_concentration"]), rtol=1e-10)
        for k in ("location", "lambda", "alpha", "beta"):
            np.testing.assert_allclose(f[k], m["mu_tau"][k], rtol=1e-10)
