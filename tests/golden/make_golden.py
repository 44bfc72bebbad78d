"""Generate the golden fixtures by running the UNMODIFIED reference live.

Runs only in the build container (needs ``/root/reference``); the fixtures it writes are
committed and are what travels to the GPU box.  The reference has no tests or golden vectors of
its own, so these outputs of the reference itself are the pin for the oracle and
for the CUDA path.

For each case it writes ``tests/golden/<case>/``:

* ``<cond>_rep<i>/data.json|data.index`` — the synthetic dataprep output (inputs);
* ``spec.json``   — conditions/runs, method, criteria (config.yml is rebuilt from it at test time);
* ``expected.json`` — per fitted position, straight from the reference objects after
  ``xpore.scripts.diffmod.execute``: prefilter p-value, n_iterations, converged, final ELBO,
  Normal-Gamma params, Dirichlet concentrations, N_gk, group names; the list of selected keys;
  and the rows ``io.generate_result_table`` emitted (post row-filter), plus the header.

Import shims: ``ujson`` -> ``json`` and an empty ``h5py`` module; neither is
installed here and neither touches the arithmetic.

    python tests/golden/make_golden.py
"""
import contextlib
import json
import os
import shutil
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)
sys.path.insert(0, "/root/reference")
sys.modules["ujson"] = json
sys.modules["h5py"] = types.ModuleType("h5py")

import pandas  # noqa: E402
from xpore.diffmod import io as ref_io  # noqa: E402
from xpore.diffmod.configurator import Configurator  # noqa: E402
from xpore.scripts import diffmod as ref_diffmod  # noqa: E402
from xpore.scripts import helper as ref_helper  # noqa: E402

from xpore_b200.synth import Spec, write_dataset  # noqa: E402


def f(x):
    x = float(x)
    return x if np.isfinite(x) else repr(x)


def run_reference(config_path):
    np.random.seed(12345)  # only seeds the discarded random q(z) init (gmm.py:207)
    config = Configurator(config_path)
    paths = config.get_paths()
    data_info = config.get_data_info()
    method = config.get_method()
    criteria = config.get_criteria()
    prior_params = config.get_priors()
    model_kmer = pandas.read_csv(paths["model_kmer"]).set_index("model_kmer")
    out_paths = {k: os.path.join(paths["out_dir"], "diffmod.%s" % k) for k in ["model", "table", "log"]}
    locks = {k: contextlib.nullcontext() for k in out_paths}
    f_index, f_data = {}, {}
    for cond, runs in data_info.items():
        for run, d in runs.items():
            df = pandas.read_csv(os.path.join(d, "data.index"))
            f_index[run] = dict(zip(df["idx"], zip(df["start"], df["end"])))
            f_data[run] = open(os.path.join(d, "data.json"))
    ids = ref_helper.get_ids(f_index, data_info)

    captured = {}
    real_table = ref_io.generate_result_table

    def spy(models, info):
        rows = real_table(models, info)
        captured["models"] = models
        captured["rows"] = rows
        return rows

    ref_diffmod.io.generate_result_table = spy
    positions, rows_all, selected = [], [], []
    try:
        for idx in ids:
            data_dict = {}
            for cond, runs in data_info.items():
                for run in runs:
                    if idx in f_index[run]:
                        s, e = f_index[run][idx]
                        f_data[run].seek(s)
                        data_dict[(cond, run)] = json.loads(f_data[run].read(e - s))
                    else:
                        data_dict[(cond, run)] = None
            sel = ref_io.load_data(idx, data_dict, criteria["readcount_min"], criteria["readcount_max"],
                                   method["pooling"])
            for key, d in sel.items():
                selected.append({"key": list(key), "n_reads": int(len(d["y"])),
                                 "group_names": list(d["condition_names"] if method["pooling"] else d["run_names"]),
                                 "y_sum": f(np.sum(d["y"]))})
            captured.clear()
            ref_diffmod.execute(idx, data_dict, data_info, method, criteria, model_kmer, prior_params,
                                out_paths, False, locks)
            for key, (m, pre) in captured.get("models", {}).items():
                mt = m.nodes["mu_tau"].params
                positions.append({
                    "key": list(key),
                    "prefilter": None if pre is None else {k: f(v) for k, v in pre.items()},
                    "n_iterations": int(m.info["n_iterations"]),
                    "converged": bool(m.info["converged"]),
                    "elbo_prev": f(m.info["log_elbos"][-1]) if m.info["log_elbos"] else None,
                    "n_elbos": len(m.info["log_elbos"]),
                    "convergence_ratio": f(m.info["convergence_ratio"]),
                    "location": [f(v) for v in mt["location"]],
                    "lambda": [f(v) for v in mt["lambda"]],
                    "alpha": [f(v) for v in mt["alpha"]],
                    "beta": [f(v) for v in mt["beta"]],
                    "concentration": [[f(v) for v in r] for r in m.nodes["w"].params["concentration"]],
                    "N": [[f(v) for v in r] for r in m.nodes["y"].params["N"]],
                    "group_names": list(m.nodes["x"].params["group_names"]),
                    "kmer_mean": f(m.kmer_signal["mean"]), "kmer_std": f(m.kmer_signal["std"]),
                })
            for row in captured.get("rows", []):
                rows_all.append([None if v is None else (v if isinstance(v, str) else f(v)) for v in row])
    finally:
        ref_diffmod.io.generate_result_table = real_table
        for fh in f_data.values():
            fh.close()
    header = ref_io.get_result_table_header(data_info, method)
    return {"ids": ids, "header": header, "selected": selected, "positions": positions, "rows": rows_all,
            "versions": {"numpy": np.__version__, "scipy": __import__("scipy").__version__,
                         "pandas": pandas.__version__}}


CASES = {
    # c1-shaped: 2 conditions x 2 replicates, run-grouped, every position fitted
    "g1_runs": dict(spec=Spec({"KO": 2, "WT": 2}, n_genes=2, pos_per_gene=20, reads_lo=50, reads_hi=200,
                              f_mod=0.4, seed=11)),
    # same inputs, t-test prefilter at 0.1 (selection must match exactly)
    "g2_prefilter": dict(spec=Spec({"KO": 2, "WT": 2}, n_genes=2, pos_per_gene=20, reads_lo=50, reads_hi=200,
                                   f_mod=0.4, seed=11,
                                   method={"prefiltering": {"method": "t-test", "threshold": 0.1}})),
    # same inputs, pooled by condition
    "g3_pooled": dict(spec=Spec({"KO": 2, "WT": 2}, n_genes=2, pos_per_gene=20, reads_lo=50, reads_hi=200,
                                f_mod=0.4, seed=11, method={"pooling": True})),
    # 3 conditions x 2 replicates: one-vs-all columns, prefilter yields NaN => all fitted
    "g4_multicond": dict(spec=Spec({"WT": 2, "KO": 2, "HEPG2": 2}, n_genes=1, pos_per_gene=24, reads_lo=20,
                                   reads_hi=80, f_mod=0.5, seed=12,
                                   method={"prefiltering": {"method": "t-test", "threshold": 0.1}})),
    # ragged: runs falling below readcount_min are dropped; small max_iters hits the iteration cap
    "g5_ragged": dict(spec=Spec({"KO": 2, "WT": 3}, n_genes=2, pos_per_gene=20, reads_lo=8, reads_hi=40,
                                f_mod=0.5, seed=13, method={"max_iters": 12})),
    # ragged + pooled + custom criteria
    "g6_ragged_pooled": dict(spec=Spec({"KO": 2, "WT": 3}, n_genes=2, pos_per_gene=20, reads_lo=8, reads_hi=40,
                                       f_mod=0.5, seed=13, method={"pooling": True},
                                       criteria={"readcount_min": 40, "readcount_max": 90})),
    # BASELINE c5 shape: 6 conditions x 3 replicates (15 pairwise + 6 one-vs-all z-tests, 63 test columns); the
    # t-test prefilter is configured and yields NaN (more than two conditions) => every position is fitted
    "g7_six_conditions": dict(spec=Spec({c: 3 for c in ["A549", "HEK293T", "HEPG2", "K562", "KO", "MCF7"]},
                                        n_genes=1, pos_per_gene=36, reads_lo=20, reads_hi=60, f_mod=0.6, seed=14,
                                        method={"prefiltering": {"method": "t-test", "threshold": 0.1}})),
}


def mutate_g5(root, spec):
    """Make the ragged cases rougher: drop gene 2 from one run, drop one position from another run."""
    info = spec.data_info(root)
    run_dirs = [d for rd in info.values() for d in rd.values()]
    # remove the second gene from the first run entirely
    d0 = run_dirs[0]
    lines = open(os.path.join(d0, "data.json")).read().splitlines(keepends=True)
    with open(os.path.join(d0, "data.json"), "w") as fj, open(os.path.join(d0, "data.index"), "w") as fi:
        fi.write("idx,start,end\n")
        for ln in lines[:1]:
            s = fj.tell()
            fj.write(ln)
            fi.write("%s,%d,%d\n" % (list(json.loads(ln).keys())[0], s, fj.tell()))
    # remove position 105 from every gene of the last run
    d1 = run_dirs[-1]
    lines = open(os.path.join(d1, "data.json")).read().splitlines()
    with open(os.path.join(d1, "data.json"), "w") as fj, open(os.path.join(d1, "data.index"), "w") as fi:
        fi.write("idx,start,end\n")
        for ln in lines:
            g = json.loads(ln)
            idx = list(g.keys())[0]
            g[idx].pop("105", None)
            s = fj.tell()
            fj.write("{")
            fj.write('"%s":' % idx)
            json.dump(g[idx], fj, separators=(",", ":"))
            fj.write("}\n")
            fi.write("%s,%d,%d\n" % (idx, s, fj.tell()))


def main():
    only = set(sys.argv[1:])
    for name, case in CASES.items():
        if only and name not in only:
            continue
        spec = case["spec"]
        root = os.path.join(HERE, name)
        if os.path.isdir(root):
            shutil.rmtree(root)
        os.makedirs(root)
        cfg = write_dataset(root, spec)
        if name.startswith(("g5", "g6")):
            mutate_g5(root, spec)
        exp = run_reference(cfg)
        shutil.rmtree(os.path.join(root, "out"), ignore_errors=True)
        os.remove(cfg)
        info = spec.data_info(".")
        with open(os.path.join(root, "spec.json"), "w") as fh:
            json.dump({"data": {c: {run.split("-", 1)[1]: os.path.basename(d) for run, d in rd.items()}
                                for c, rd in info.items()},
                       "method": spec.method, "criteria": spec.criteria}, fh, indent=1)
        with open(os.path.join(root, "expected.json"), "w") as fh:
            json.dump(exp, fh)
        print(name, "ids", len(exp["ids"]), "selected", len(exp["selected"]), "fitted", len(exp["positions"]),
              "rows", len(exp["rows"]),
              "iters", sorted(p["n_iterations"] for p in exp["positions"])[-3:])


if __name__ == "__main__":
    main()
