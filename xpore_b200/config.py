"""YAML configuration of ``xpore diffmod`` — same keys, defaults and quirks as the reference.

Mirrors ``xpore/diffmod/configurator.py:10-77`` (class ``Configurator``): method names and
return shapes are kept so driver code reads like the reference's ``diffmod()``
(``xpore/scripts/diffmod.py:87-92``).  Quirks kept on purpose:

* run names become ``"<condition>-<replicate>"`` (``configurator.py:7-8,33``);
* ``criteria`` defaults to ``readcount_min=15, readcount_max=1000`` only when the whole key is
  absent (``:36-44``);
* ``method`` defaults (``:46-60``); only ``max_iters``, ``stopping_criteria``, ``pooling`` and
  ``prefiltering{method,threshold}`` are read by the path;
* ``get_priors`` ignores a YAML ``priors`` section (``:62-77``): lambda0 = 1, beta scale = 0.5,
  Dirichlet concentration = 0.001; the configured ``alpha=[0.5,0.5]`` is never used, alpha0 is
  the k-mer precision 1/stdv^2 (``xpore/scripts/diffmod.py:40-41``).
"""
from __future__ import annotations

import os
from collections import OrderedDict

import yaml


def get_condition_run_name(condition_name, run_name):
    return "-".join([str(condition_name), str(run_name)])


class Configurator:
    def __init__(self, config_filepath):
        self.filepath = os.path.abspath(config_filepath)
        self.filename = os.path.basename(self.filepath)
        with open(self.filepath, "r") as f:
            self.yaml = yaml.safe_load(f)

    def get_paths(self):
        paths = {}
        if "prior" in self.yaml:
            paths["model_kmer"] = os.path.abspath(self.yaml["prior"])
        else:
            paths["model_kmer"] = None  # shipped table (xpore_b200/data/model_kmer.npz)
        paths["out_dir"] = os.path.abspath(self.yaml["out"])
        os.makedirs(paths["out_dir"], exist_ok=True)
        paths["models"] = os.path.join(paths["out_dir"], "models")
        os.makedirs(paths["models"], exist_ok=True)
        paths["model_filepath"] = os.path.join(paths["out_dir"], "models", "%s.model")
        return paths

    def get_data_info(self):
        data = OrderedDict()
        for condition_name, run_names in self.yaml["data"].items():
            data.setdefault(condition_name, OrderedDict())
            for run_name, dirpath in run_names.items():
                data[condition_name][get_condition_run_name(condition_name, run_name)] = dirpath
        return data

    def get_criteria(self):
        if "criteria" in self.yaml:
            return self.yaml["criteria"]
        return {"readcount_min": 15, "readcount_max": 1000}

    def get_method(self):
        method = self.yaml["method"] if "method" in self.yaml else {}
        method.setdefault("name", "gmm")
        method.setdefault("max_iters", 500)
        method.setdefault("stopping_criteria", 0.00001)
        method.setdefault("compute_elbo", True)
        method.setdefault("verbose", False)
        method.setdefault("update", ["z", "y", "w", "mu_tau"])
        method.setdefault("pooling", False)
        method.setdefault("prefiltering", False)
        return method

    def get_priors(self):
        return {
            "mu_tau": {
                "location": ["model_kmer_mean", "model_kmer_mean"],
                "lambda": [1, 1],
                "alpha": [0.5, 0.5],
                "beta": ["model_kmer_tau", "model_kmer_tau"],
                "beta_scale": [0.5, 0.5],
            },
            "w": {"concentration": [0.001, 0.001]},
        }
