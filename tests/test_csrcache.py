"""Binary CSR hand-off (xpore_b200.csrcache): same batches and the same diffmod.table as the data.json path;
stale caches are refused."""
import argparse
import os

import numpy as np

from xpore_b200 import csrcache
from xpore_b200 import diffmod as xd
from xpore_b200.batch import Layout
from xpore_b200.config import Configurator
from xpore_b200.ingest import NativeIngest
from xpore_b200.kmer_model import KmerModel

from conftest import load_case
from helpers import emul_fit_batch, load_emulator, read_table


def _setup(case):
    cfg = Configurator(case["config_path"])
    data_info, method, criteria = cfg.get_data_info(), cfg.get_method(), cfg.get_criteria()
    km = KmerModel.load(cfg.get_paths()["model_kmer"])
    lay = Layout.from_data_info(data_info, method["pooling"])
    f_index = {run: xd.read_index(d) for runs in data_info.values() for run, d in runs.items()}
    return cfg, data_info, method, criteria, km, lay, f_index


def test_cache_reproduces_ingest(golden_case, tmp_path):
    cfg, data_info, method, criteria, km, lay, f_index = _setup(golden_case)
    path = str(tmp_path / "inputs.xpb")
    info = csrcache.build_cache(golden_case["config_path"], path, n_threads=2)
    cache = csrcache.load_cache(path, data_info, lay, criteria, km)
    assert cache is not None and info["positions"] == len(cache.positions)
    ids = xd.get_ids(f_index, data_info)
    assert cache.ids == ids
    reader = NativeIngest(data_info, lay, km, criteria["readcount_min"], criteria["readcount_max"])
    for pick in (ids, ids[::-1], ids[:1], []):
        want = reader.read(pick, f_index)
        got = cache.batch(pick, lay)
        assert got.n_positions == want.n_positions
        np.testing.assert_array_equal(got.y[:got.n_reads], want.y[:want.n_reads])  # bit-identical doubles
        np.testing.assert_array_equal(got.read_off, want.read_off)
        np.testing.assert_array_equal(got.grp_len, want.grp_len)
        np.testing.assert_array_equal(got.kmer_mean, want.kmer_mean)
        np.testing.assert_array_equal(got.kmer_std, want.kmer_std)
        assert (got.ids, got.positions, got.kmers) == (want.ids, want.positions, want.kmers)
        if want.n_positions:
            fmt = want.encode_reads()
            assert got.y_format == fmt
            if fmt:
                np.testing.assert_array_equal(got.y_enc, want.y_enc)
    reader.close()


def test_stale_or_foreign_cache_is_refused(tmp_path):
    case = load_case("g1_runs", tmp_path)
    cfg, data_info, method, criteria, km, lay, f_index = _setup(case)
    path = str(tmp_path / "inputs.xpb")
    csrcache.build_cache(case["config_path"], path)
    assert csrcache.load_cache(path, data_info, lay, criteria, km) is not None
    # other criteria / pooling mode: different selection, the cache must not be used
    assert csrcache.load_cache(path, data_info, lay, dict(criteria, readcount_min=criteria["readcount_min"] + 1), km) is None
    assert csrcache.load_cache(path, data_info, Layout.from_data_info(data_info, not method["pooling"]), criteria, km) is None
    # an input file changed after the cache was written
    d = next(iter(next(iter(data_info.values())).values()))
    st = os.stat(os.path.join(d, "data.json"))
    os.utime(os.path.join(d, "data.json"), ns=(st.st_atime_ns, st.st_mtime_ns + 1_000_000_000))
    try:
        assert csrcache.load_cache(path, data_info, lay, criteria, km) is None
    finally:
        os.utime(os.path.join(d, "data.json"), ns=(st.st_atime_ns, st.st_mtime_ns))
    # not a cache at all
    junk = tmp_path / "junk.xpb"
    junk.write_bytes(b"not a cache")
    assert csrcache.load_cache(str(junk), data_info, lay, criteria, km) is None
    assert csrcache.load_cache(str(tmp_path / "missing.xpb"), data_info, lay, criteria, km) is None


def test_diffmod_table_identical_with_cache(tmp_path):
    case = load_case("g2_prefilter", tmp_path)
    emu = load_emulator()

    def run(**kw):
        args = argparse.Namespace(config=case["config_path"], n_processes=1, save_models=False, resume=False, ids=[],
                                  batch_genes=8192, batch_bytes=256 << 20, csr_cache=kw.get("csr_cache"))
        xd.diffmod(args, fit_fn=lambda b, m: emul_fit_batch(emu, b, m))
        out = Configurator(case["config_path"]).get_paths()["out_dir"]
        return open(os.path.join(out, "diffmod.table")).read(), open(os.path.join(out, "diffmod.log")).read()

    plain = run()
    path = str(tmp_path / "inputs.xpb")
    first = run(csr_cache=path)      # builds the cache, then reads it
    assert os.path.exists(path)
    again = run(csr_cache=path)      # reuses it
    assert plain == first == again
    header, rows = read_table(os.path.join(Configurator(case["config_path"]).get_paths()["out_dir"], "diffmod.table"))
    assert rows
