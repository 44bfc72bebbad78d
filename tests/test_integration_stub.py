"""The reference-side binding INTEGRATION.md shows a maintainer (section 1b) is executed as written: the code block is
extracted from the document, loaded as a module and run on a golden gene prepared the way `io.load_data` prepares it
(xpore/diffmod/io.py:21-90); its records must equal the product packer + `fit_batch` on the same gene."""
import os
import re
import types

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def stub_source():
    text = open(os.path.join(REPO, "INTEGRATION.md")).read()
    blocks = re.findall(r"```python\n(.*?)```", text, flags=re.S)
    src = [b for b in blocks if "def fit_gene(" in b]
    assert len(src) == 1
    return src[0]


def test_stub_mirrors_the_header():
    """CPU: the struct mirrors of the stub have the very fields of the product binding (xpore_b200/_lib.py), which
    tests/test_abi.py pins to include/xpore_b200.h."""
    from xpore_b200 import _lib
    src = stub_source()
    for cls in ("XpBatch", "XpMethod", "XpResult"):
        names = re.findall(r'\("(\w+)", C\.c_\w+\)', src[src.index("class %s" % cls):src.index("]", src.index("class %s" % cls)) + 400]
                           .split("class ")[1])
        assert names == [f[0] for f in getattr(_lib, cls)._fields_], cls


@pytest.mark.gpu
@pytest.mark.parametrize("case_name", ["g2_prefilter", "g3_pooled", "g5_ragged"])
def test_stub_runs_against_a_golden_gene(case_name, tmp_path):
    import pandas
    import torch
    assert torch.cuda.is_available()
    from oracle import xpore_oracle as orc
    from xpore_b200 import _lib
    from xpore_b200 import fit as xf
    from xpore_b200.kmer_model import KmerModel

    from conftest import load_case
    from helpers import gene_dict, open_case, pack_case
    _lib.build()
    os.environ["XPORE_B200_LIB"] = _lib.LIB_PATH
    mod = types.ModuleType("b200_stub")
    exec(compile(stub_source(), "INTEGRATION.md:stub", "exec"), mod.__dict__)
    case = load_case(case_name, tmp_path)
    cfg, data_info, method, criteria, index, handles = open_case(case["config_path"])
    km = KmerModel.load()
    model_kmer = pandas.DataFrame({"model_mean": km.mean, "model_stdv": km.stdv}, index=km.kmers)
    idx = case["expected"]["ids"][0]
    data = orc.select_positions(idx, gene_dict(idx, data_info, index, handles), criteria["readcount_min"],
                                criteria["readcount_max"], bool(method["pooling"]))
    for d in data.values():  # the per-read labels io.load_data keeps beside the one-hot matrices (io.py:86-88)
        d["y_condition_names"] = np.asarray(d["condition_names"])[d["x"].argmax(axis=1)]
        d["y_run_names"] = np.asarray(d["run_names"])[d["r"].argmax(axis=1)]
    out = mod.fit_gene(data, data_info, method, model_kmer)
    batch, _, _ = pack_case(case["config_path"], [idx])
    res = xf.fit_batch(batch, xf.Method.from_dict(method))
    assert list(data.keys()) == list(zip(batch.ids, batch.positions, batch.kmers))
    np.testing.assert_array_equal(out["status"], res.status)
    np.testing.assert_array_equal(out["n_iter"], res.n_iter)
    np.testing.assert_array_equal(out["pval"], res.pval)
    sel = res.selected
    assert sel.any()
    np.testing.assert_array_equal(out["fit"][sel], res.fit[sel])
    np.testing.assert_array_equal(out["stats"][sel], res.stats[sel])
