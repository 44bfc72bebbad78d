"""1000+ positions of every BASELINE shape against outputs of the UNMODIFIED reference (tests/golden/big/*.npz,
made by tests/golden/make_golden_big.py; the inputs are regenerated from the seed and checked by digest).

CPU: the oracle against the reference on a sample (keeps the oracle pinned at these shapes, 6 conditions x 3
replicates included).  GPU: the CUDA path through the C ABI against the reference on EVERY position - selection,
prefilter p-values, sweep counts, posterior parameters, emitted rows - and EVERY statistic of every fitted position
(kept by the row filter or not) against the oracle's statistics evaluated on the reference's own parameters.
"""
import hashlib
import json
import os

import numpy as np
import pytest

from xpore_b200 import fit as xf
from xpore_b200.kmer_model import KmerModel
from xpore_b200.synth import Spec, make_batch

from helpers import compare_row_cells, compare_stats_with_oracle, half_integer_tests

BIG = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "big")
SHAPES = ["c1", "c2", "c3", "c4", "c5"]


def load_shape(name):
    with open(os.path.join(BIG, "specs.json")) as f:
        sj = json.load(f)[name]
    spec = Spec(sj["conditions"], n_genes=sj["n_genes"], pos_per_gene=sj["pos_per_gene"], reads_lo=sj["reads_lo"],
                reads_hi=sj["reads_hi"], f_mod=sj["f_mod"], seed=sj["seed"],
                loguniform_total=tuple(sj["loguniform_total"]) if sj["loguniform_total"] else None,
                method=sj["method"] or {}, criteria=sj["criteria"])
    gold = np.load(os.path.join(BIG, name + ".npz"))
    batch = make_batch(spec, KmerModel.load(), keys=True)
    h = hashlib.sha256()
    for a in (batch.y[:batch.n_reads], batch.read_off, batch.grp_len, batch.kmer_mean, batch.kmer_std):
        h.update(np.ascontiguousarray(a).tobytes())
    assert h.hexdigest() == str(gold["digest"]), "the regenerated inputs differ from the ones the golden was made on"
    return spec, batch, gold


def golden_fit(gold, batch, p):
    """The reference's posterior of position p in the oracle's dict form (groups = the present slots, slot order)."""
    present = np.nonzero(batch.grp_len[p] > 0)[0]
    par = gold["params"][p]
    N = gold["N"][p][present]
    return {"location": par[0:2], "lambda": par[2:4], "alpha": par[4:6], "beta": par[6:8], "N": N, "conc": 0.001 + N}, present


def test_oracle_on_big_golden_sample():
    """The oracle reproduces the reference at the BASELINE shapes (a sample of positions per shape; the CPU suite
    must stay short), C = 6 / G = 18 included: prefilter p-value, sweep count, parameters."""
    from oracle import xpore_oracle as orc
    for name in SHAPES:
        spec, batch, gold = load_shape(name)
        method = xf.Method.from_dict(spec.method)
        rng = np.random.default_rng(1)
        n_take = 6 if name in ("c4", "c5") else 10
        gc = batch.layout.grp_cond
        for p in rng.choice(batch.n_positions, size=n_take, replace=False):
            y = batch.y[batch.read_off[p]:batch.read_off[p + 1]]
            if method.prefilter_method:
                x, _ = orc.one_hot(np.repeat(gc, batch.grp_len[p]))
                pv = orc.prefilter_pvalue(y, x)
                assert (np.isnan(pv) and np.isnan(gold["pval"][p])) or pv == gold["pval"][p]
                assert bool(gold["fitted"][p]) == bool(np.isnan(pv) or pv < method.prefilter_threshold)
            if not gold["fitted"][p]:
                continue
            gf, present = golden_fit(gold, batch, p)
            X = np.zeros((len(y), len(present)), dtype=bool)
            o = 0
            for j, g in enumerate(present):
                X[o:o + batch.grp_len[p, g], j] = True
                o += batch.grp_len[p, g]
            f = orc.fit_position(y, X, batch.kmer_mean[p], batch.kmer_std[p], method.max_iters, method.stopping_criteria,
                                 rng=np.random.RandomState(0))
            assert f["n_iterations"] == gold["n_iter"][p] and f["converged"] == bool(gold["converged"][p])
            for k in ("location", "lambda", "alpha", "beta"):
                np.testing.assert_allclose(f[k], gf[k], rtol=1e-12)
            np.testing.assert_allclose(f["N"], gf["N"], rtol=1e-12, atol=1e-300)


@pytest.mark.parametrize("name", ["c1", "c5"])
def test_kernel_arithmetic_on_big_golden(name):
    """The host build of the kernels' arithmetic (tests/emul) against the reference on every position of two shapes:
    the same selection, sweep counts and parameters.  (c5 holds a position whose median read equals the k-mer mean:
    the reference's choice of the second start location, gmm.py:39-44, then hangs on the rounding of numpy's lerp.)"""
    from helpers import emul_fit_batch, load_emulator
    spec, batch, gold = load_shape(name)
    method = xf.Method.from_dict(spec.method)
    res = emul_fit_batch(load_emulator(), batch, method)
    f = gold["fitted"]
    np.testing.assert_array_equal(res.selected, f)
    np.testing.assert_array_equal(res.n_iter[f], gold["n_iter"][f])
    par = np.concatenate([res.mu, res.lam, res.alpha, res.beta], axis=1)
    np.testing.assert_allclose(par[f], gold["params"][f], rtol=1e-9)


@pytest.mark.gpu
@pytest.mark.parametrize("name", SHAPES)
def test_cuda_path_on_big_golden(name):
    import torch
    assert torch.cuda.is_available(), "gpu tests need a CUDA device"
    from xpore_b200.table import get_result_table_header, result_rows
    spec, batch, gold = load_shape(name)
    method = xf.Method.from_dict(spec.method)
    pm = method.prefilter_method
    batch.encode_reads()
    res = xf.fit_batch(batch, method)
    P = batch.n_positions
    # selection (diffmod.py:52-58): exact; prefilter p-value (statstest.py:11-18)
    np.testing.assert_array_equal(res.selected, gold["fitted"])
    assert not (res.status & (xf.NOT_FITTED | xf.NEAR_THRESHOLD)).any()
    if pm:
        ok = ~np.isnan(gold["pval"])
        assert (np.isnan(res.pval) == ~ok).all()
        np.testing.assert_allclose(res.pval[ok], gold["pval"][ok], rtol=1e-4)
        if ok.any():
            print("%s: max rel. deviation of the t-test p-value %.2e" % (name, np.max(np.abs(res.pval[ok] / gold["pval"][ok] - 1))))
    fitted = np.nonzero(gold["fitted"])[0]
    # sweep counts, convergence flags, posterior parameters: every fitted position
    same = int((res.n_iter[fitted] == gold["n_iter"][fitted]).sum())
    print("%s: %d of %d fitted positions with the reference's sweep count" % (name, same, len(fitted)))
    assert same == len(fitted)
    np.testing.assert_array_equal((res.status[fitted] & xf.CONVERGED) != 0, gold["converged"][fitted])
    par = np.concatenate([res.mu, res.lam, res.alpha, res.beta], axis=1)
    np.testing.assert_allclose(par[fitted], gold["params"][fitted], rtol=1e-5)
    gN = gold["N"][fitted]
    have = ~np.isnan(gN)
    np.testing.assert_allclose(res.N[fitted][have], gN[have], rtol=1e-5, atol=1e-9)
    print("%s: max rel. deviation of mu/lambda/alpha/beta %.2e" % (name, np.max(np.abs(par[fitted] / gold["params"][fitted] - 1))))
    # emitted rows (io.py:346-367): the same set, the same cells
    header = [str(h) for h in gold["header"]]
    lay = batch.layout
    assert get_result_table_header(lay, pm) == header
    all_rows = result_rows(batch, res, pm, only_kept=False)
    row_of = {int(p): r for p, r in zip(np.nonzero(res.selected)[0], all_rows)}
    kept = np.nonzero(res.keep_row)[0]
    np.testing.assert_array_equal(kept, np.sort(gold["row_pos"]))
    num_cols = [i for i, h in enumerate(header) if i >= 3 and h != "mod_assignment"]
    a_col = header.index("mod_assignment")
    knife = bound = 0
    for j, p in enumerate(gold["row_pos"]):
        exp = list(row_of[int(p)][:3]) + [None] * (len(header) - 3)
        for c, v, none in zip(num_cols, gold["row_vals"][j], gold["row_none"][j]):
            exp[c] = None if none else float(v)
        exp[a_col] = "higher" if gold["row_higher"][j] else "lower"
        knife += compare_row_cells(exp[:3], header, row_of[int(p)], exp, lay, pm)
        bound += half_integer_tests(batch, int(p))
    # every statistic of every fitted position (kept or not) against the oracle evaluated on the reference's parameters
    data_info = spec.data_info()
    knife_all = bound_all = 0
    for p in fitted:
        gf, present = golden_fit(gold, batch, int(p))
        knife_all += compare_stats_with_oracle(batch, res, int(p), gf, present, data_info, pm, row_of[int(p)], header)
        bound_all += half_integer_tests(batch, int(p))
    print("%s: knife-edge z-test cells accepted: %d of %d eligible in the %d reference rows; %d of %d eligible over all %d "
          "fitted positions" % (name, knife, bound, len(kept), knife_all, bound_all, len(fitted)))
    assert knife <= bound and knife_all <= bound_all
