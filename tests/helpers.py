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
