"""Golden outputs of the UNMODIFIED reference on >= 1000 positions of every BASELINE shape.

The inputs are not stored: they are regenerated from the seed by ``xpore_b200.synth.make_batch`` (numpy
``default_rng``; a digest of the arrays is stored and checked by the tests).  What is stored, per shape, in
``tests/golden/big/<shape>.npz`` is what the reference computed, straight from its objects after
``xpore.scripts.diffmod.execute`` (import shims as in make_golden.py):

* for EVERY position: prefilter p-value, fitted flag;
* for every FITTED position: n_iterations, converged, location / lambda / alpha / beta, N[g][k] in this
  repository's slot order (NaN where a group is absent);
* the rows ``io.generate_result_table`` emitted (the reference can only produce the rows that pass its filter,
  io.py:363-367): position index + the numeric cells (row_none marks the empty ones) + mod_assignment.

Runs only in the build container (needs ``/root/reference``):

    python tests/golden/make_golden_big.py [shape ...]
"""
import contextlib
import hashlib
import json
import multiprocessing as mp
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)
sys.path.insert(0, "/root/reference")
sys.modules["ujson"] = json
sys.modules["h5py"] = types.ModuleType("h5py")

from xpore_b200.synth import Spec, make_batch  # noqa: E402

SIX = ["A549", "HEK293T", "HEPG2", "K562", "KO", "MCF7"]
# BASELINE.json shapes, 1000+ positions each, 100 positions per gene.  Read counts are chosen inside
# the read-count criteria, so no position or run is filtered and positions keep their index in either grouping.
SHAPES = {
    "c2": dict(spec=Spec({"KO": 3, "WT": 3}, n_genes=12, pos_per_gene=100, reads_lo=20, reads_hi=100, f_mod=0.3, seed=101,
                         method={"prefiltering": {"method": "t-test", "threshold": 0.1}})),
    "c3": dict(spec=Spec({"KO": 3, "WT": 3}, n_genes=10, pos_per_gene=100, reads_lo=20, reads_hi=100, f_mod=0.3, seed=102,
                         method={"pooling": True, "max_iters": 500})),
    "c4": dict(spec=Spec({"KO": 2, "WT": 2}, n_genes=10, pos_per_gene=100, f_mod=0.3, seed=103, loguniform_total=(20, 5000),
                         criteria={"readcount_min": 5, "readcount_max": 5000})),
    "c5": dict(spec=Spec({c: 3 for c in SIX}, n_genes=10, pos_per_gene=100, reads_lo=20, reads_hi=60, f_mod=0.3, seed=104,
                         method={"prefiltering": {"method": "t-test", "threshold": 0.1}})),
    # c1 shape at scale: two replicates per condition, where the z-test's n = round(mean coverage) sits on x.5 for
    # about half of all positions
    "c1": dict(spec=Spec({"KO": 2, "WT": 2}, n_genes=10, pos_per_gene=100, reads_lo=50, reads_hi=200, f_mod=0.4, seed=105,
                         method={"prefiltering": {"method": "t-test", "threshold": 0.1}})),
}


def batch_digest(batch):
    h = hashlib.sha256()
    for a in (batch.y[:batch.n_reads], batch.read_off, batch.grp_len, batch.kmer_mean, batch.kmer_std):
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def runs_spec(spec):
    """The same spec grouped by run (the reference's inputs are always per run)."""
    m = {k: v for k, v in spec.method.items() if k != "pooling"}
    return Spec(spec.conditions, spec.ko_like, spec.n_genes, spec.pos_per_gene, spec.reads_lo, spec.reads_hi, spec.f_mod,
                spec.seed, spec.loguniform_total, m, spec.criteria)


def _gene_task(args):
    """One gene through the unmodified reference's execute(); returns plain arrays."""
    (gi, idx, keys, ys, glens, data_info, method, criteria, prior_params, run_names) = args
    import pandas
    from xpore.diffmod import io as ref_io
    from xpore.scripts import diffmod as ref_diffmod

    np.random.seed(12345 + gi)  # only seeds the discarded random q(z) init (gmm.py:207)
    model_kmer = pandas.read_csv("/root/reference/xpore/diffmod/model_kmer.csv").set_index("model_kmer")
    data_dict = {}
    r = 0
    for cond, runs in data_info.items():
        for run in runs:
            d = {}
            for (pos, kmer), y, gl in zip(keys, ys, glens):
                o = int(gl[:r].sum())
                d[pos] = {kmer: [float(v) for v in y[o:o + gl[r]]]}
            data_dict[(cond, run)] = {idx: d}
            r += 1
    captured = {}
    real_table = ref_io.generate_result_table

    def spy(models, info):
        rows = real_table(models, info)
        captured["models"] = models
        captured["rows"] = rows
        return rows

    ref_diffmod.io.generate_result_table = spy
    out_paths = {k: os.devnull for k in ["model", "table", "log"]}
    locks = {k: contextlib.nullcontext() for k in out_paths}
    try:
        ref_diffmod.execute(idx, data_dict, data_info, method, criteria, model_kmer, prior_params, out_paths, False, locks)
    finally:
        ref_diffmod.io.generate_result_table = real_table
    # the prefilter p-value of EVERY position (execute() keeps it only for the positions it fits): the reference's own
    # load_data + StatsTest on the same inputs (diffmod.py:20,52-53)
    pvals = {}
    if method["prefiltering"]:
        from xpore.diffmod.statstest import StatsTest
        sel = ref_io.load_data(idx, data_dict, criteria["readcount_min"], criteria["readcount_max"], method["pooling"])
        for key, d in sel.items():
            pvals[(key[1], key[2])] = float(StatsTest(d).fit(method=method["prefiltering"]["method"]))
    fits = {}
    for key, (m, pre) in captured.get("models", {}).items():
        mt = m.nodes["mu_tau"].params
        fits[(key[1], key[2])] = dict(
            pre=None if pre is None else float(list(pre.values())[0]),
            n_iter=int(m.info["n_iterations"]), converged=bool(m.info["converged"]),
            location=np.asarray(mt["location"], float), lam=np.asarray(mt["lambda"], float),
            alpha=np.asarray(mt["alpha"], float), beta=np.asarray(mt["beta"], float),
            N=np.asarray(m.nodes["y"].params["N"], float), group_names=list(m.nodes["x"].params["group_names"]))
    rows = [list(r) for r in captured.get("rows", [])]
    return gi, fits, rows, pvals


def run_shape(name, spec, procs):
    import pandas  # noqa: F401
    from xpore.diffmod import io as ref_io
    from xpore.diffmod.configurator import Configurator

    from xpore_b200.batch import Layout
    from xpore_b200.kmer_model import KmerModel
    km = KmerModel.load()
    rb = make_batch(runs_spec(spec), km, keys=True)              # per run: what the reference reads
    batch = make_batch(spec, km, keys=True)                      # the grouping under test
    P = batch.n_positions
    assert P == rb.n_positions == spec.n_genes * spec.pos_per_gene, "a position was filtered: counts must fit the criteria"
    # the reference's own defaults for method / criteria / priors
    cfg_yaml = {"data": {c: {run.split("-", 1)[1]: "." for run in rd} for c, rd in spec.data_info().items()}, "out": os.devnull}
    if spec.method:
        cfg_yaml["method"] = dict(spec.method)
    if spec.criteria:
        cfg_yaml["criteria"] = dict(spec.criteria)
    import tempfile

    import yaml
    with tempfile.TemporaryDirectory() as td:
        cfg_yaml["out"] = os.path.join(td, "out")
        path = os.path.join(td, "c.yml")
        with open(path, "w") as f:
            yaml.safe_dump(cfg_yaml, f, sort_keys=False)
        config = Configurator(path)
        data_info, method, criteria, prior_params = (config.get_data_info(), config.get_method(), config.get_criteria(),
                                                     config.get_priors())
    lay = Layout.from_data_info(data_info, method["pooling"])
    assert lay.group_names == batch.layout.group_names
    header = ref_io.get_result_table_header(data_info, method)
    tasks = []
    for gi in range(spec.n_genes):
        sl = range(gi * spec.pos_per_gene, (gi + 1) * spec.pos_per_gene)
        tasks.append((gi, rb.ids[sl[0]], [(rb.positions[p], rb.kmers[p]) for p in sl],
                      [rb.y[rb.read_off[p]:rb.read_off[p + 1]] for p in sl], [rb.grp_len[p] for p in sl], data_info, method,
                      criteria, prior_params, rb.layout.run_names))
    with mp.get_context("fork").Pool(procs) as pool:
        results = pool.map(_gene_task, tasks, chunksize=1)
    G = lay.n_groups
    fitted = np.zeros(P, bool)
    pval = np.full(P, np.nan)
    n_iter = np.zeros(P, np.int32)
    converged = np.zeros(P, bool)
    params = np.full((P, 8), np.nan)
    N = np.full((P, G, 2), np.nan)
    row_pos, row_vals, row_assign, row_none = [], [], [], []
    num_cols = [i for i, h in enumerate(header) if i >= 3 and h != "mod_assignment"]
    a_col = header.index("mod_assignment")
    for gi, fits, rows, pvals in sorted(results, key=lambda t: t[0]):
        base = gi * spec.pos_per_gene
        index_of = {(batch.positions[p], batch.kmers[p]): p for p in range(base, base + spec.pos_per_gene)}
        for k, v in pvals.items():
            pval[index_of[k]] = v
        for k, f in fits.items():
            p = index_of[k]
            fitted[p] = True
            if f["pre"] is not None:
                assert (np.isnan(f["pre"]) and np.isnan(pval[p])) or pval[p] == f["pre"]
            n_iter[p], converged[p] = f["n_iter"], f["converged"]
            params[p] = np.concatenate([f["location"], f["lam"], f["alpha"], f["beta"]])
            for j, nm in enumerate(f["group_names"]):
                N[p, lay.group_names.index(nm)] = f["N"][j]
        for r in rows:
            row_pos.append(index_of[(r[1], r[2])])
            row_vals.append([np.nan if r[i] is None else float(r[i]) for i in num_cols])
            row_none.append([r[i] is None for i in num_cols])
            row_assign.append(r[a_col])
    out = os.path.join(HERE, "big")
    os.makedirs(out, exist_ok=True)
    np.savez_compressed(os.path.join(out, name + ".npz"), digest=batch_digest(batch), header=np.array(header),
                        fitted=fitted, pval=pval, n_iter=n_iter, converged=converged, params=params, N=N,
                        row_pos=np.asarray(row_pos, np.int64), row_vals=np.asarray(row_vals, float).reshape(len(row_pos), len(num_cols)),
                        row_higher=np.asarray([a == "higher" for a in row_assign], bool),
                        row_none=np.asarray(row_none, bool).reshape(len(row_pos), len(num_cols)),
                        versions=np.array([np.__version__, __import__("scipy").__version__]))
    print(name, "positions", P, "fitted", int(fitted.sum()), "rows", len(row_pos), "max sweeps", int(n_iter.max()))


def spec_json(spec):
    return dict(conditions=dict(spec.conditions), n_genes=spec.n_genes, pos_per_gene=spec.pos_per_gene,
                reads_lo=spec.reads_lo, reads_hi=spec.reads_hi, f_mod=spec.f_mod, seed=spec.seed,
                loguniform_total=spec.loguniform_total, method=spec.method, criteria=spec.criteria)


def main():
    only = set(sys.argv[1:])
    procs = len(os.sched_getaffinity(0))
    for name, case in SHAPES.items():
        if only and name not in only:
            continue
        run_shape(name, case["spec"], procs)
    out = os.path.join(HERE, "big")
    with open(os.path.join(out, "specs.json"), "w") as f:
        json.dump({k: spec_json(v["spec"]) for k, v in SHAPES.items()}, f, indent=1)


if __name__ == "__main__":
    main()
