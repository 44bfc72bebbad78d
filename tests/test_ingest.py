"""Native ingest (csrc/xp_ingest.cpp through the C ABI) against the Python statement of the
reference's selection rules (xpore_b200.batch.BatchBuilder, itself checked against the unmodified
reference's selected keys in tests/test_oracle_golden.py / test_emul_parity.py): bit-identical
batches on every golden case, on adversarial hand-written genes, and the error behaviour."""
import json
import os

import numpy as np
import pytest

from xpore_b200 import _lib
from xpore_b200.batch import BatchBuilder, Layout
from xpore_b200.ingest import NativeIngest
from xpore_b200.kmer_model import KmerModel

from helpers import gene_dict, open_case


@pytest.fixture(scope="module", autouse=True)
def _built():
    _lib.build()


def _same(a, b):
    assert a.ids == b.ids and a.positions == b.positions and a.kmers == b.kmers
    np.testing.assert_array_equal(a.read_off, b.read_off)
    np.testing.assert_array_equal(a.grp_len, b.grp_len)
    assert a.y.tobytes() == b.y[:a.n_reads].tobytes()  # bit-identical doubles
    np.testing.assert_array_equal(a.kmer_mean, b.kmer_mean)
    np.testing.assert_array_equal(a.kmer_std, b.kmer_std)


@pytest.mark.parametrize("threads", [1, 4])
def test_native_ingest_matches_python_packer(golden_case, threads):
    exp = golden_case["expected"]
    cfg, data_info, method, criteria, index, handles = open_case(golden_case["config_path"])
    lay = Layout.from_data_info(data_info, method["pooling"])
    km = KmerModel.load()
    lo, hi = criteria["readcount_min"], criteria["readcount_max"]
    bb = BatchBuilder(lay, km, lo, hi)
    for idx in exp["ids"]:
        bb.add_gene(idx, gene_dict(idx, data_info, index, handles))
    for h in handles.values():
        h.close()
    want = bb.build()
    ni = NativeIngest(data_info, lay, km, lo, hi)
    got = ni.read(exp["ids"], index, n_threads=threads)
    ni.close()
    assert got.n_positions == want.n_positions > 0
    _same(got, want)


def _write_runs(tmp_path, runs):
    """runs: {(cond, run): {idx: {pos: {kmer: [floats]}}} or raw text per idx}; returns data_info, index."""
    data_info, index = {}, {}
    for (cond, run), genes in runs.items():
        d = tmp_path / ("%s_%s" % (cond, run))
        d.mkdir()
        index[run] = {}
        with open(d / "data.json", "w") as fj, open(d / "data.index", "w") as fi:
            fi.write("idx,start,end\n")
            pos = 0
            for idx, body in genes.items():
                text = body if isinstance(body, str) else json.dumps({idx: body})
                fj.write(text + "\n")
                index[run][idx] = (pos, pos + len(text))
                fi.write("%s,%d,%d\n" % (idx, pos, pos + len(text)))
                pos += len(text) + 1
        data_info.setdefault(cond, {})[run] = str(d)
    return data_info, index


def _python_batch(data_info, index, ids, pooling, lo, hi):
    lay = Layout.from_data_info(data_info, pooling)
    km = KmerModel.load()
    handles = {run: open(os.path.join(d, "data.json")) for c, rd in data_info.items() for run, d in rd.items()}
    bb = BatchBuilder(lay, km, lo, hi)
    for idx in ids:
        bb.add_gene(idx, gene_dict(idx, data_info, index, handles))
    for h in handles.values():
        h.close()
    return lay, km, bb.build()


@pytest.mark.parametrize("pooling", [False, True])
def test_selection_edge_cases(tmp_path, pooling):
    rng = np.random.default_rng(5)
    r = lambda n: [round(float(v), 2) for v in rng.normal(100, 3, n)]  # noqa: E731
    runs = {
        # g1: position 12 missing in one run (dropped by the intersection); 7 has a run below the minimum;
        #     9 is only rich enough in one condition; 100 sorts after 20 (numeric, not text, order);
        #     two k-mers at position 30
        ("KO", "k1"): {"g1": {"7": {"AAAAA": r(3)}, "9": {"ACGTA": r(40)}, "12": {"CCCCC": r(30)},
                              "100": {"GGGGG": r(25)}, "20": {"TTTTT": r(25)},
                              "30": {"AACCA": r(20), "AACCC": r(22)}},
                       "g2": {"5": {"ACACA": r(16)}}},
        ("KO", "k2"): {"g1": {"7": {"AAAAA": r(30)}, "9": {"ACGTA": r(2)}, "12": {"CCCCC": r(30)},
                              "100": {"GGGGG": r(1200)}, "20": {"TTTTT": r(25)},
                              "30": {"AACCA": r(20), "AACCC": r(22)}}},
        ("WT", "w1"): {"g1": {"7": {"AAAAA": r(30)}, "9": {"ACGTA": r(1)}, "100": {"GGGGG": r(25)},
                              "20": {"TTTTT": r(25)}, "30": {"AACCA": r(20), "AACCC": r(22)}},
                       # g2 written with spaces, exponents and integers: still the same doubles
                       "g2": '{"g2": {"5": {"ACACA": [ 101, 9.95e1, 100.25 ,99.5,1.0e2,100,100,100,100,100,100,100,100,100,100,100 ]}}}'},
    }
    data_info, index = _write_runs(tmp_path, runs)
    ids = ["g1", "g2", "missing"]
    for run in index:
        index[run].pop("missing", None)
    lay, km, want = _python_batch(data_info, index, ids, pooling, 15, 1000)
    ni = NativeIngest(data_info, lay, km, 15, 1000)
    got = ni.read(ids, index, n_threads=2)
    ni.close()
    assert want.n_positions >= 3
    _same(got, want)
    g1 = [p for i, p in zip(got.ids, got.positions) if i == "g1"]
    assert g1 == sorted(g1, key=int)


def test_unknown_kmer_and_bad_json(tmp_path):
    runs = {("A", "a1"): {"g": {"1": {"NNNNZ": [1.0] * 20}}}, ("B", "b1"): {"g": {"1": {"NNNNZ": [1.0] * 20}}}}
    data_info, index = _write_runs(tmp_path, runs)
    lay = Layout.from_data_info(data_info, False)
    km = KmerModel.load()
    ni = NativeIngest(data_info, lay, km, 15, 1000)
    with pytest.raises(KeyError):  # model_kmer.loc[kmer] raises KeyError in the reference (diffmod.py:23)
        ni.read(["g"], index)
    ni.close()
    bad = tmp_path / "bad"
    bad.mkdir()
    (bad / "x").mkdir()
    runs = {("A", "a1"): {"g": '{"g": {"1": {"AAAAA": [1.0, }}}'}, ("B", "b1"): {"g": {"1": {"AAAAA": [1.0] * 20}}}}
    data_info, index = _write_runs(bad / "x", runs)
    ni = NativeIngest(data_info, Layout.from_data_info(data_info, False), km, 15, 1000)
    with pytest.raises(RuntimeError):
        ni.read(["g"], index)
    ni.close()
