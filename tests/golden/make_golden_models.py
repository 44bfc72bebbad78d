"""Golden node values for ``--save_models``: run the UNMODIFIED reference live on two of the committed
datasets and keep, for a few positions, everything ``io.save_models_to_hdf5`` (io.py:99-139) would write:
the nodes x (group matrix as a group index per read, group_names), y (reads, N, mean, variance), z (prob,
ln_prob), w (concentration), mu_tau (location, lambda, alpha, beta), y_run_names / y_condition_names, and
the kmer attribute.  Runs only in the build container (needs /root/reference).

    python tests/golden/make_golden_models.py
"""
import contextlib
import json
import os
import sys
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))
sys.path.insert(0, "/root/reference")
sys.modules["ujson"] = json
sys.modules["h5py"] = types.ModuleType("h5py")

import pandas  # noqa: E402
from xpore.diffmod import io as ref_io  # noqa: E402
from xpore.diffmod.configurator import Configurator  # noqa: E402
from xpore.scripts import diffmod as ref_diffmod  # noqa: E402
from xpore.scripts import helper as ref_helper  # noqa: E402

from conftest import load_case  # noqa: E402

KEEP = 4  # positions per case


def run(case_name):
    import pathlib
    with tempfile.TemporaryDirectory() as td:
        case = load_case(case_name, pathlib.Path(td))
        np.random.seed(12345)
        config = Configurator(case["config_path"])
        paths, data_info, method = config.get_paths(), config.get_data_info(), config.get_method()
        criteria, prior_params = config.get_criteria(), config.get_priors()
        model_kmer = pandas.read_csv(paths["model_kmer"]).set_index("model_kmer")
        out_paths = {k: os.path.join(paths["out_dir"], "diffmod.%s" % k) for k in ["model", "table", "log"]}
        locks = {k: contextlib.nullcontext() for k in out_paths}
        f_index, f_data = {}, {}
        for cond, runs in data_info.items():
            for r, d in runs.items():
                df = pandas.read_csv(os.path.join(d, "data.index"))
                f_index[r] = dict(zip(df["idx"], zip(df["start"], df["end"])))
                f_data[r] = open(os.path.join(d, "data.json"))
        ids = ref_helper.get_ids(f_index, data_info)
        captured = {}
        real = ref_io.generate_result_table

        def spy(models, info):
            captured["models"] = models
            return real(models, info)

        ref_diffmod.io.generate_result_table = spy
        out = []
        try:
            idx = ids[0]
            data_dict = {}
            for cond, runs in data_info.items():
                for r in runs:
                    if idx in f_index[r]:
                        s, e = f_index[r][idx]
                        f_data[r].seek(s)
                        data_dict[(cond, r)] = json.loads(f_data[r].read(e - s))
                    else:
                        data_dict[(cond, r)] = None
            ref_diffmod.execute(idx, data_dict, data_info, method, criteria, model_kmer, prior_params, out_paths,
                                False, locks)
            keys = sorted(captured["models"], key=lambda k: (int(k[1]), k[2]))[:KEEP]
            for key in keys:
                m, _ = captured["models"][key]
                n = m.nodes
                x = np.asarray(n["x"].data)
                out.append({
                    "key": list(key),
                    "kmer": key[2],
                    "x_group_names": list(n["x"].params["group_names"]),
                    "x_group_of_read": [int(v) for v in np.argmax(x, axis=1)],
                    "y_data": [float(v) for v in n["y"].data],
                    "y_N": np.asarray(n["y"].params["N"]).tolist(),
                    "y_mean": np.asarray(n["y"].params["mean"]).tolist(),
                    "y_variance": np.asarray(n["y"].params["variance"]).tolist(),
                    "z_prob": np.asarray(n["z"].params["prob"]).tolist(),
                    "z_ln_prob": np.asarray(n["z"].params["ln_prob"]).tolist(),
                    "w_concentration": np.asarray(n["w"].params["concentration"]).tolist(),
                    "mu_tau": {k: np.asarray(v).tolist() for k, v in n["mu_tau"].params.items()},
                    "y_run_names": list(n["y_run_names"].data),
                    "y_condition_names": list(n["y_condition_names"].data),
                })
        finally:
            ref_diffmod.io.generate_result_table = real
            for fh in f_data.values():
                fh.close()
    return out


def main():
    for name in ("g1_runs", "g3_pooled"):
        models = run(name)
        with open(os.path.join(HERE, name, "models.json"), "w") as fh:
            json.dump(models, fh)
        print(name, [m["key"] for m in models], [len(m["y_data"]) for m in models])


if __name__ == "__main__":
    main()
