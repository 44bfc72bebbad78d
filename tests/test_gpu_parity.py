"""Parity of the CUDA path (through the C ABI) with the reference / the oracle.  Needs a B200."""
import ctypes as C

import numpy as np
import pytest
import scipy.special
import scipy.stats

from xpore_b200 import _lib
from xpore_b200 import fit as xf
from xpore_b200.kmer_model import KmerModel
from xpore_b200.synth import Spec, make_batch

from helpers import compare_stats_with_oracle, compare_with_golden, half_integer_tests, pack_case

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    assert torch.cuda.is_available(), "gpu tests need a CUDA device"
    _lib.build()
    return torch


def dev_special(torch, which, x, aux=None):
    L = _lib.lib()
    xd = torch.tensor(np.ascontiguousarray(x, dtype=np.float64), device="cuda")
    ad = torch.tensor(np.ascontiguousarray(aux if aux is not None else np.zeros_like(x), dtype=np.float64), device="cuda")
    out = torch.empty_like(xd)
    _lib.check(L.xpb200_eval_special(which, xd.data_ptr(), ad.data_ptr(), out.data_ptr(), xd.numel(), None))
    torch.cuda.synchronize()
    return out.cpu().numpy()


def test_device_special_functions(torch_cuda):
    torch = torch_cuda
    x = np.concatenate([np.geomspace(1e-3, 1e5, 4000), np.linspace(0.001, 30, 3000)])
    np.testing.assert_allclose(dev_special(torch, 0, x), scipy.special.digamma(x), rtol=2e-13, atol=2e-14)
    np.testing.assert_allclose(dev_special(torch, 1, x), scipy.special.gammaln(x), rtol=2e-13, atol=2e-14)
    xl = np.geomspace(1e-300, 1e300, 5000)
    np.testing.assert_allclose(dev_special(torch, 2, xl), np.log(xl), rtol=4e-16)
    xe = np.linspace(-700, 700, 20001)
    np.testing.assert_allclose(dev_special(torch, 3, xe), np.exp(xe), rtol=1e-15)
    d = np.concatenate([np.linspace(-60, 60, 4001), [-1e4, -600, -500.5, 500.5, 600, 1e4, 0.0]])
    np.testing.assert_allclose(dev_special(torch, 5, d), scipy.special.expit(np.clip(d, -500, 500)), rtol=2e-15,
                               atol=1e-300)
    rm = scipy.special.expit(-np.abs(d))
    ent = scipy.special.xlogy(rm, rm) + (1 - rm) * np.log1p(-rm)
    np.testing.assert_allclose(dev_special(torch, 6, d), ent, rtol=1e-12, atol=1e-15)
    rng = np.random.default_rng(0)
    t = np.concatenate([rng.normal(0, 3, 4000), rng.normal(0, 30, 500)])
    df = rng.integers(2, 2000, 4500).astype(np.float64)
    ref = scipy.stats.t.sf(np.abs(t), df) * 2
    got = dev_special(torch, 4, t, df)
    ok = ref > 1e-290
    np.testing.assert_allclose(got[ok], ref[ok], rtol=5e-12)


def test_golden_through_c_abi(torch_cuda, golden_case):
    """Host buffers -> xpb200_fit_host -> records, against the unmodified reference's outputs:
    selection and fitted set exact, parameters rel 1e-5, p-values rel 1e-4 (north_star)."""
    exp = golden_case["expected"]
    batch, method, _ = pack_case(golden_case["config_path"], exp["ids"])
    res = xf.fit_batch(batch, xf.Method.from_dict(method))
    iters_diff, knife = compare_with_golden(batch, res, method, exp, rtol_param=1e-5, rtol_p=1e-4, iters_exact=False)
    print("%s: positions with a different sweep count: %d of %d; knife-edge z-tests: %d"
          % (golden_case["name"], iters_diff, res.n_fitted, knife))
    assert iters_diff == 0


def _oracle_fit(batch, p, method):
    from oracle import xpore_oracle as orc
    y = batch.y[batch.read_off[p]:batch.read_off[p + 1]]
    present = np.nonzero(batch.grp_len[p] > 0)[0]
    X = np.zeros((len(y), len(present)), dtype=bool)
    o = 0
    for j, g in enumerate(present):
        X[o:o + batch.grp_len[p, g], j] = True
        o += batch.grp_len[p, g]
    f = orc.fit_position(y, X, batch.kmer_mean[p], batch.kmer_std[p], method.max_iters, method.stopping_criteria,
                         rng=np.random.RandomState(1))
    return f, present


@pytest.mark.parametrize("shape", ["c2", "c3_pooled", "c5_multicond", "c4_ragged"])
def test_device_path_vs_oracle(torch_cuda, shape):
    """Seeded synthetic batches of the BASELINE shapes, device-resident entry point, against the
    CPU oracle on a sample of positions (the oracle needs ~40 ms per position)."""
    specs = {
        "c2": (Spec({"KO": 3, "WT": 3}, reads_lo=20, reads_hi=100, f_mod=0.3, seed=21,
                    method={"prefiltering": {"method": "t-test", "threshold": 0.1}}), 600, 40),
        "c3_pooled": (Spec({"KO": 3, "WT": 3}, reads_lo=20, reads_hi=100, f_mod=0.3, seed=22,
                           method={"pooling": True}), 300, 30),
        "c5_multicond": (Spec({c: 3 for c in ["A549", "HEK293T", "HEPG2", "K562", "KO", "MCF7"]}, reads_lo=20,
                              reads_hi=60, f_mod=0.3, seed=23,
                              method={"prefiltering": {"method": "t-test", "threshold": 0.1}}), 200, 20),
        "c4_ragged": (Spec({"KO": 2, "WT": 2}, f_mod=0.3, seed=24, loguniform_total=(20, 5000),
                           criteria={"readcount_min": 5, "readcount_max": 5000}), 300, 24),
    }
    spec, P, n_check = specs[shape]
    torch = torch_cuda
    batch = make_batch(spec, KmerModel.load(), n_positions=P)
    method = xf.Method.from_dict(spec.method)
    db = xf.DeviceBatch(batch)
    out = xf.fit_device(db, method)
    torch.cuda.synchronize()
    res = out.to_host()
    # the host-buffer entry point must give the same bits
    res2 = xf.fit_batch(batch, method)
    assert (res.status == res2.status).all() and (res.n_iter == res2.n_iter).all()
    sel = res.selected
    np.testing.assert_array_equal(res.fit[sel], res2.fit[sel])
    np.testing.assert_array_equal(res.stats[sel], res2.stats[sel])
    # prefilter: exact selection against scipy on every position
    from oracle import xpore_oracle as orc
    if method.prefilter_method:
        gc = batch.layout.grp_cond
        for p in range(batch.n_positions):
            y = batch.y[batch.read_off[p]:batch.read_off[p + 1]]
            cond = np.repeat(gc, batch.grp_len[p])
            x, _ = orc.one_hot(cond)
            pv = orc.prefilter_pvalue(y, x)
            if np.isnan(pv):
                assert np.isnan(res.pval[p]) and sel[p]
            else:
                np.testing.assert_allclose(res.pval[p], pv, rtol=1e-9)
                assert sel[p] == bool(pv < method.prefilter_threshold)
    else:
        assert sel.all() and np.isnan(res.pval).all()
    assert res.n_fitted == int(sel.sum())
    # fitted parameters and statistics against the oracle on a sample
    rng = np.random.default_rng(5)
    idx = np.nonzero(sel)[0]
    idx = rng.choice(idx, size=min(n_check, len(idx)), replace=False)
    same_iters = 0
    from xpore_b200.table import get_result_table_header, result_rows
    batch.ids = ["id"] * batch.n_positions
    batch.positions = [str(i) for i in range(batch.n_positions)]
    batch.kmers = ["NNNNN"] * batch.n_positions
    pm = method.prefilter_method
    header = get_result_table_header(batch.layout, pm)
    row_of = dict(zip(np.nonzero(sel)[0].tolist(), result_rows(batch, res, pm, only_kept=False)))
    knife = bound = 0
    for p in idx:
        f, present = _oracle_fit(batch, p, method)
        same_iters += int(f["n_iterations"] == res.n_iter[p])
        np.testing.assert_allclose(res.mu[p], f["location"], rtol=1e-5)
        np.testing.assert_allclose(res.alpha[p], f["alpha"], rtol=1e-5)
        np.testing.assert_allclose(res.beta[p], f["beta"], rtol=1e-5)
        np.testing.assert_allclose(res.N[p][present], f["N"], rtol=1e-5, atol=1e-9)
        assert bool(res.status[p] & xf.CONVERGED) == f["converged"]
        # every statistic of the position, kept by the row filter or not (io.py:255-367)
        knife += compare_stats_with_oracle(batch, res, int(p), f, present, spec.data_info(), pm, row_of[int(p)], header)
        bound += half_integer_tests(batch, int(p))
    print("%s: %d/%d sampled positions with the oracle's sweep count; knife-edge z-test cells %d of %d eligible"
          % (shape, same_iters, len(idx), knife, bound))
    assert same_iters == len(idx)
    assert knife <= bound


@pytest.mark.parametrize("shift,expect", [(0.0, 2), (-200.0, 1), (1e-9, 0)])
def test_encoded_reads_bit_identical(torch_cuda, shift, expect):
    """u16 / i32 scaled-integer read buffers decode to the very same doubles: every fit / statistics
    record must be bit-identical to the f64 path, through both entry points; non-representable data
    stays f64.  The prefilter of the u16 format works on exact integer sums (its t statistic carries
    only the final roundings), so its p-values agree with the f64 path to rounding, not bit for bit;
    the selection must be identical."""
    torch = torch_cuda
    spec = Spec({"KO": 2, "WT": 3}, reads_lo=10, reads_hi=70, f_mod=0.4, seed=41,
                method={"prefiltering": {"method": "t-test", "threshold": 0.2}})
    batch = make_batch(spec, KmerModel.load(), n_positions=700)
    batch.y[:] = np.round(batch.y + shift, 2) if shift <= 0 else batch.y + shift
    method = xf.Method.from_dict(spec.method)
    ref = xf.fit_batch(batch, method)
    assert batch.encode_reads() == expect
    enc = xf.fit_batch(batch, method)
    out = xf.fit_device(xf.DeviceBatch(batch), method)
    torch.cuda.synchronize()
    dev = out.to_host()
    for r in (enc, dev):
        np.testing.assert_array_equal(r.status, ref.status)
        np.testing.assert_array_equal(r.n_iter, ref.n_iter)
        if expect == 2:
            np.testing.assert_allclose(r.pval, ref.pval, rtol=1e-11, atol=0, equal_nan=True)
        else:
            np.testing.assert_array_equal(r.pval, ref.pval)
        sel = ref.selected
        np.testing.assert_array_equal(r.fit[sel], ref.fit[sel])
        np.testing.assert_array_equal(r.stats[sel], ref.stats[sel])
    assert 50 < ref.n_fitted < 700


def _permute_slots(batch, perm):
    """The same batch with its group slots in the order `perm` (reads of a position reordered accordingly)."""
    from dataclasses import replace
    from xpore_b200.batch import Batch, Layout
    lay = batch.layout
    assert not lay.pooling
    new_lay = Layout(lay.condition_names, [lay.run_names[j] for j in perm], [lay.run_cond[j] for j in perm], False)
    P, G = batch.grp_len.shape
    y = np.empty_like(batch.y)
    o = 0
    for p in range(P):
        starts = batch.read_off[p] + np.concatenate([[0], np.cumsum(batch.grp_len[p])[:-1]])
        for j in perm:
            n = batch.grp_len[p, j]
            y[o:o + n] = batch.y[starts[j]:starts[j] + n]
            o += n
    return Batch(new_lay, y, batch.read_off.copy(), np.ascontiguousarray(batch.grp_len[:, perm]), batch.kmer_mean.copy(),
                 batch.kmer_std.copy())


def _drop_runs(batch, drops):
    """Remove the reads of run j at position p for every (p, j) in drops (in place: y, read_off, grp_len)."""
    P, G = batch.grp_len.shape
    pieces, glen = [], batch.grp_len.copy()
    dropped = set(drops)
    for p in range(P):
        o = batch.read_off[p]
        for j in range(G):
            n = batch.grp_len[p, j]
            if (p, j) in dropped:
                glen[p, j] = 0
            else:
                pieces.append(batch.y[o:o + n])
            o += n
    batch.y = np.ascontiguousarray(np.concatenate(pieces))
    batch.grp_len = glen
    batch.read_off = np.concatenate([[0], np.cumsum(glen.sum(axis=1))]).astype(np.int64)
    return batch.validate()


@pytest.mark.parametrize("perm", [[2, 3, 4, 0, 1], [0, 2, 1, 3, 4], [4, 0, 3, 1, 2]])
def test_prefilter_slot_orders(torch_cuda, perm):
    """The u16 prefilter keeps exact integer sums, so its p-values must not depend on the order of the group slots:
    condition blocks swapped (prefix = the higher condition: masked streaming path) and conditions interleaved
    (group-walk path) against the builder's order, bit for bit; the synthetic batch has absent runs and
    one-condition positions."""
    spec = Spec({"KO": 2, "WT": 3}, reads_lo=1, reads_hi=60, f_mod=0.4, seed=43,
                method={"prefiltering": {"method": "t-test", "threshold": 0.2}})
    base = make_batch(spec, KmerModel.load(), n_positions=600)
    _drop_runs(base, [(p, j) for p in range(0, 600, 37) for j in ((0, 1) if p % 2 else (2, 3, 4))])  # one condition left
    method = xf.Method.from_dict(spec.method)
    assert base.encode_reads() == 2
    ref = xf.fit_batch(base, method)
    other = _permute_slots(base, perm)
    assert other.encode_reads() == 2
    res = xf.fit_batch(other, method)
    np.testing.assert_array_equal(res.pval, ref.pval)
    np.testing.assert_array_equal(res.status & 1, ref.status & 1)
    assert np.isnan(ref.pval).sum() > 0 and (~np.isnan(ref.pval)).sum() > 400  # one-condition positions exist


def test_empty_and_tiny_batches(torch_cuda):
    spec = Spec({"KO": 2, "WT": 2}, reads_lo=15, reads_hi=16, seed=3)
    b0 = make_batch(spec, KmerModel.load(), n_positions=0) if False else None
    b1 = make_batch(spec, KmerModel.load(), n_positions=1)
    r1 = xf.fit_batch(b1, xf.Method())
    assert r1.n_fitted == 1 and r1.n_iter[0] >= 2 and np.isfinite(r1.fit[0]).all()
    # max_iters edge: at most max_iters-1 sweeps (gmm.py:103)
    r2 = xf.fit_batch(b1, xf.Method(max_iters=3))
    assert r2.n_iter[0] == 2
    # empty batch
    from xpore_b200.batch import Batch
    e = Batch(b1.layout, np.zeros(0), np.zeros(1, np.int64), np.zeros((0, 4), np.int32), np.zeros(0), np.zeros(0))
    r0 = xf.fit_batch(e, xf.Method())
    assert r0.n_fitted == 0 and r0.fit.shape[0] == 0


def test_full_size_properties(torch_cuda):
    """Size-independent properties on a large batch (100k positions, c2 shape): permutation
    invariance of the result records and determinism across runs."""
    torch = torch_cuda
    spec = Spec({"KO": 3, "WT": 3}, reads_lo=20, reads_hi=100, f_mod=0.1, seed=31,
                method={"prefiltering": {"method": "t-test", "threshold": 0.1}})
    batch = make_batch(spec, KmerModel.load(), n_positions=100000)
    method = xf.Method.from_dict(spec.method)
    res = xf.fit_device(xf.DeviceBatch(batch), method)
    torch.cuda.synchronize()
    a = res.to_host()
    perm = np.random.default_rng(0).permutation(batch.n_positions)
    shuffled = batch.subset(perm)
    b = xf.fit_batch(shuffled, method)
    assert a.n_fitted == b.n_fitted and 0.1 * batch.n_positions < a.n_fitted < 0.5 * batch.n_positions
    np.testing.assert_array_equal(a.status[perm], b.status)
    np.testing.assert_array_equal(a.n_iter[perm], b.n_iter)
    np.testing.assert_array_equal(a.pval[perm], b.pval)
    sel = b.selected
    np.testing.assert_array_equal(a.fit[perm][sel], b.fit[sel])
    np.testing.assert_array_equal(a.stats[perm][sel], b.stats[sel])
    # sanity of the fitted values
    assert np.isfinite(b.fit[sel]).all()
    N = b.N[sel]
    n_reads = np.diff(shuffled.read_off)[sel]
    np.testing.assert_allclose(N.sum(axis=(1, 2)), n_reads, rtol=1e-12)
    assert (b.n_iter[sel] >= 2).all() and (b.n_iter[sel] <= 499).all()


def test_pack_records_device(torch_cuda):
    """Compact records (payload of the final gather) = the selected rows of the dense arrays."""
    torch = torch_cuda
    from xpore_b200.dist import unpack_records
    spec = Spec({"KO": 2, "WT": 2}, reads_lo=20, reads_hi=60, f_mod=0.4, seed=51,
                method={"prefiltering": {"method": "t-test", "threshold": 0.1}})
    batch = make_batch(spec, KmerModel.load(), n_positions=3000)
    out = xf.fit_device(xf.DeviceBatch(batch), xf.Method.from_dict(spec.method))
    rec, _ = xf.pack_records_device(out, index_offset=7000)
    res = out.to_host()
    idx, status, n_iter, pval, fit, stats = unpack_records(rec.cpu().numpy(), xf.fit_width(4), xf.stats_width(4, 2))
    assert len(idx) == res.n_fitted == int(res.selected.sum())
    order = np.argsort(idx)
    sel = np.nonzero(res.selected)[0]
    np.testing.assert_array_equal(idx[order] - 7000, sel)
    np.testing.assert_array_equal(status[order], res.status[sel])
    np.testing.assert_array_equal(n_iter[order], res.n_iter[sel])
    np.testing.assert_array_equal(pval[order], res.pval[sel])
    np.testing.assert_array_equal(fit[order], res.fit[sel])
    np.testing.assert_array_equal(stats[order], res.stats[sel])
    # longest positions first (worklist order)
    n = np.diff(batch.read_off)[idx - 7000]
    assert n[0] >= n[-1]


def test_scaled_integer_decode_exhaustive(torch_cuda):
    """The division-free decode of u16/i32 reads equals IEEE q/100 for every u16 value and for random
    i32 values (prefilter p-values of single-value positions would differ otherwise); checked through
    the fit of a batch whose reads enumerate all u16 codes.  512 reads per position: one warp per position in either
    encoding (the tensor-memory kernel for u16, the shared-memory one for f64 / i32), whose summation orders are the
    same - the records can only be bit-identical if every decoded read is."""
    torch = torch_cuda
    from xpore_b200.batch import Batch, Layout
    lay = Layout(["A", "B"], ["A-r1", "B-r1"], [0, 1], False)
    rng = np.random.default_rng(3)
    for fmt, q in ((2, np.arange(65536, dtype=np.int64)), (1, rng.integers(-2 ** 31 + 2, 2 ** 31 - 2, 65536))):
        y = q / 100.0
        P = 128
        per = len(y) // P
        grp_len = np.full((P, 2), per // 2, dtype=np.int32)
        read_off = np.arange(P + 1, dtype=np.int64) * per
        b = Batch(lay, y.copy(), read_off, grp_len, np.full(P, 100.0), np.full(P, 3.0))
        m = xf.Method(max_iters=3, prefilter_method="t-test", prefilter_threshold=2.0)
        ref = xf.fit_batch(b, m)
        assert b.encode_reads() == fmt
        enc = xf.fit_batch(b, m)
        if fmt == 2:  # the u16 prefilter sums integers exactly: equal to the rounding of the t statistic, which a p-value of
            # 1e-155 (512 reads on either side, far apart) magnifies by its logarithmic slope; the fit decodes every read
            np.testing.assert_allclose(enc.pval, ref.pval, rtol=1e-6, atol=0, equal_nan=True)
        else:
            np.testing.assert_array_equal(ref.pval, enc.pval)
        np.testing.assert_array_equal(ref.fit, enc.fit)


def test_repeated_launches_bit_identical(torch_cuda):
    """The same resident batch fitted five times gives the very same bits every time, also when the
    shared memory of every SM is overwritten with NaN patterns / other data in between: the fit
    kernel must not depend on what a previous launch left in shared memory."""
    torch = torch_cuda
    spec = Spec({"KO": 3, "WT": 3}, reads_lo=20, reads_hi=100, f_mod=0.1, seed=33,
                method={"prefiltering": {"method": "t-test", "threshold": 0.1}})
    batch = make_batch(spec, KmerModel.load(), n_positions=60000)
    batch.encode_reads()
    method = xf.Method.from_dict(spec.method)
    db = xf.DeviceBatch(batch)
    out = xf.DeviceResult(db.P, db.G, db.C, db.device)
    L = _lib.lib()
    ref = None
    for pattern in (None, None, 0x7ff8000000000000, 0x4059000000000000, None):
        if pattern is not None:
            _lib.check(L.xpb200_debug_fill_smem(pattern, None))
        xf.fit_device(db, method, out)
        torch.cuda.synchronize()
        got = out.to_host()
        sel = got.selected
        assert not np.isnan(got.fit[sel][:, :9]).any()
        if ref is None:
            ref = got
            continue
        np.testing.assert_array_equal(got.status, ref.status)
        np.testing.assert_array_equal(got.n_iter, ref.n_iter)
        np.testing.assert_array_equal(got.fit[sel], ref.fit[sel])
        np.testing.assert_array_equal(got.stats[sel], ref.stats[sel])


def test_cli_writes_reference_table(torch_cuda, golden_case, tmp_path):
    """The drop-in CLI end to end on the GPU: data.json/data.index -> native ingest -> CUDA path ->
    diffmod.table/diffmod.log, against the rows the unmodified reference emitted (north-star tolerances);
    again through the binary CSR hand-off, which must give the very same file; then `postprocessing`."""
    import os

    from xpore_b200 import cli
    from xpore_b200.batch import Layout
    from xpore_b200.config import Configurator

    from helpers import compare_tables, read_table
    exp = golden_case["expected"]
    cfg = Configurator(golden_case["config_path"])
    out = cfg.get_paths()["out_dir"]
    method = cfg.get_method()
    lay = Layout.from_data_info(cfg.get_data_info(), method["pooling"])
    pm = method["prefiltering"]["method"] if method["prefiltering"] else None
    cli.main(["xpore", "diffmod", "--config", golden_case["config_path"], "--n_processes", "2"])
    table = os.path.join(out, "diffmod.table")
    header, rows = read_table(table)
    assert header == exp["header"]
    compare_tables(rows, {tuple(r[:3]): r for r in exp["rows"]}, header, lay, pm, rtol_param=1e-5, rtol_p=1e-4)
    lines = open(os.path.join(out, "diffmod.log")).read().splitlines()
    assert lines[0] == "--- DIFFMOD ---" and lines[-1] == "--- SUCCESSFULLY FINISHED ---" and lines[1:-1] == exp["ids"]
    first = open(table).read()
    cache = str(tmp_path / "inputs.xpb")
    for _ in range(2):  # builds the cache, then reuses it
        cli.main(["xpore", "diffmod", "--config", golden_case["config_path"], "--csr_cache", cache])
        assert open(table).read() == first
    cli.main(["xpore", "postprocessing", "--diffmod_dir", out])
    kept = open(os.path.join(out, "majority_direction_kmer_diffmod.table")).read().splitlines()
    assert kept[0] == first.splitlines()[0] and set(kept[1:]) <= set(first.splitlines()[1:])


def test_edge_case_positions_vs_oracle(torch_cuda):
    """Hand-made positions at the corners of the input space, one batch, every position against the
    oracle: identical reads (zero variance), two far-apart clusters (|d| beyond the +-500 clip of
    gmm.py:240), reads far from the k-mer mean, the smallest groups the selection lets through, a run
    missing at a position, and read counts on both sides of the 1-warp / 4-warp split (1023 / 1024 for f64 reads,
    1535 / 1536 for the one-position-per-warp tensor-memory kernel of a deep u16 batch)."""
    from xpore_b200.batch import Batch, Layout
    rng = np.random.default_rng(11)
    lay = Layout(["KO", "WT"], ["KO-r1", "KO-r2", "WT-r1", "WT-r2"], [0, 0, 1, 1], False)
    km, ks = 100.0, 3.0
    cases = []

    def add(groups, kmean=km, kstd=ks):
        cases.append(([np.round(np.asarray(g, dtype=np.float64), 2) for g in groups], kmean, kstd))

    add([np.full(20, 101.37)] * 4)                                               # zero variance
    add([np.r_[np.full(15, 60.0), np.full(15, 140.0)]] * 4)                        # two point masses, 80 apart
    add([rng.normal(100, 3, 40), rng.normal(100, 3, 40), rng.normal(160, 1, 40), rng.normal(160, 1, 40)])
    add([rng.normal(30, 2, 25) for _ in range(4)], kmean=120.0, kstd=2.0)        # all reads 45 sigma off the prior
    add([rng.normal(100, 3, 2), rng.normal(100, 3, 2), rng.normal(108, 4, 2), rng.normal(108, 4, 2)])   # 2 reads per run
    add([rng.normal(100, 3, 30), [], rng.normal(106, 4, 30), rng.normal(106, 4, 31)])                    # a run absent
    for n in (1023, 1024, 1025, 1535, 1536, 1537):                                # single-warp / 4-warp team splits
        per = [n // 4 + (1 if j < n % 4 else 0) for j in range(4)]
        add([np.r_[rng.normal(100, 3, m - m // 3), rng.normal(109, 4, m // 3)] if j >= 2 else rng.normal(100, 3, m)
             for j, m in enumerate(per)])
    add([np.r_[rng.normal(95, 2.5, 50), rng.normal(104, 2.5, 50)] for _ in range(4)])                    # both clusters everywhere

    y = np.concatenate([np.concatenate(g) for g, _, _ in cases])
    glen = np.asarray([[len(x) for x in g] for g, _, _ in cases], dtype=np.int32)
    off = np.zeros(len(cases) + 1, dtype=np.int64)
    np.cumsum(glen.sum(axis=1), out=off[1:])
    batch = Batch(lay, y, off, glen, np.asarray([c[1] for c in cases]), np.asarray([c[2] for c in cases]))
    method = xf.Method()  # no prefilter: every position is fitted
    for encode in (False, True):
        if encode:
            assert batch.encode_reads() == 2
        res = xf.fit_batch(batch, method)
        assert res.n_fitted == len(cases)
        for p in range(len(cases)):
            f, present = _oracle_fit(batch, p, method)
            assert res.n_iter[p] == f["n_iterations"], (p, res.n_iter[p], f["n_iterations"])
            assert bool(res.status[p] & xf.CONVERGED) == f["converged"], p
            np.testing.assert_allclose(res.mu[p], f["location"], rtol=1e-5, err_msg=str(p))
            np.testing.assert_allclose(res.alpha[p], f["alpha"], rtol=1e-5, err_msg=str(p))
            np.testing.assert_allclose(res.beta[p], f["beta"], rtol=1e-5, err_msg=str(p))
            np.testing.assert_allclose(res.N[p][present], f["N"], rtol=1e-5, atol=1e-9, err_msg=str(p))


def test_pair_class_boundary_vs_oracle(torch_cuda):
    """A u16 batch without deep positions runs two positions per warp (576 reads each in tensor memory); positions from
    576 reads up go to the 4-warp teams.  Read counts on both sides of that split, an odd number of positions (one warp
    ends up with a single position) and a warp whose two positions stop many sweeps apart - every position against the
    oracle; then the f64 twin (one position per warp in shared memory): bit-identical below the split."""
    from xpore_b200.batch import Batch, Layout
    rng = np.random.default_rng(23)
    lay = Layout(["KO", "WT"], ["KO-r1", "KO-r2", "KO-r3", "WT-r1", "WT-r2", "WT-r3"], [0, 0, 0, 1, 1, 1], False)
    G = 6
    km, ks = 100.0, 3.0
    sizes = [575, 576, 577, 600, 64, 33, 130, 511, 512, 513, 96]
    groups = []
    for q, n in enumerate(sizes):
        per = [n // G + (1 if j < n % G else 0) for j in range(G)]
        shift = 0.4 if q % 3 == 0 else 8.0  # barely separated mixtures take many sweeps, clear ones few
        groups.append([np.round(np.r_[rng.normal(km, ks, m - m // 3), rng.normal(km + shift, ks, m // 3)] if j >= 3
                                else rng.normal(km, ks, m), 2) for j, m in enumerate(per)])
    y = np.concatenate([np.concatenate(g) for g in groups])
    glen = np.asarray([[len(x) for x in g] for g in groups], dtype=np.int32)
    off = np.zeros(len(sizes) + 1, dtype=np.int64)
    np.cumsum(glen.sum(axis=1), out=off[1:])
    batch = Batch(lay, y, off, glen, np.full(len(sizes), km), np.full(len(sizes), ks))
    method = xf.Method()
    ref = xf.fit_batch(batch, method)  # f64: shared-memory kernel
    assert batch.encode_reads() == 2
    res = xf.fit_batch(batch, method)  # u16: tensor-memory kernel, two positions per warp
    assert res.n_fitted == len(sizes)
    assert len(set(res.n_iter.tolist())) > 3
    for p in range(len(sizes)):
        f, present = _oracle_fit(batch, p, method)
        assert res.n_iter[p] == f["n_iterations"], (p, sizes[p], res.n_iter[p], f["n_iterations"])
        assert bool(res.status[p] & xf.CONVERGED) == f["converged"], p
        np.testing.assert_allclose(res.mu[p], f["location"], rtol=1e-5, err_msg=str(p))
        np.testing.assert_allclose(res.alpha[p], f["alpha"], rtol=1e-5, err_msg=str(p))
        np.testing.assert_allclose(res.beta[p], f["beta"], rtol=1e-5, err_msg=str(p))
        np.testing.assert_allclose(res.N[p][present], f["N"], rtol=1e-5, atol=1e-9, err_msg=str(p))
    below = np.asarray(sizes) < 576
    np.testing.assert_array_equal(res.n_iter, ref.n_iter)
    np.testing.assert_array_equal(res.fit[below], ref.fit[below])
    np.testing.assert_array_equal(res.stats[below], ref.stats[below])
    np.testing.assert_allclose(res.fit, ref.fit, rtol=1e-11, atol=0)


def test_fit_host_rows_equals_dense_records(torch_cuda):
    """xpb200_fit_host_rows returns exactly the records of the table's rows (fitted + row filter) that the
    dense xpb200_fit_host arrays hold, ascending by position; too small a buffer is an error, not a truncation."""
    from xpore_b200.dist import unpack_records
    spec = Spec({"KO": 3, "WT": 3}, reads_lo=20, reads_hi=100, f_mod=0.3, seed=19,
                method={"prefiltering": {"method": "t-test", "threshold": 0.1}})
    batch = make_batch(spec, KmerModel.load(), n_positions=150000)  # several chunks of the host pipeline
    batch.encode_reads()
    method = xf.Method.from_dict(spec.method)
    dense = xf.fit_batch(batch, method)
    rows, n_fitted = xf.fit_rows(batch, method)
    assert n_fitted == dense.n_fitted
    keep = np.nonzero(dense.keep_row)[0]
    assert 100 < len(keep) < dense.n_fitted
    idx, status, n_iter, pval, fit, stats = unpack_records(rows, xf.fit_width(6), xf.stats_width(6, 2))
    np.testing.assert_array_equal(idx, keep)
    np.testing.assert_array_equal(status, dense.status[keep])
    np.testing.assert_array_equal(n_iter, dense.n_iter[keep])
    np.testing.assert_array_equal(pval, dense.pval[keep])
    np.testing.assert_array_equal(fit, dense.fit[keep])
    np.testing.assert_array_equal(stats, dense.stats[keep])
    with pytest.raises(RuntimeError, match="rows buffer too small"):
        xf.fit_rows(batch, method, capacity_rows=len(keep) - 1)
    rows2, _ = xf.fit_rows(batch, method, capacity_rows=len(keep))
    np.testing.assert_array_equal(rows2, rows)


def test_chained_host_pipeline(torch_cuda):
    """xpb200_fit_host_rows on a u16 batch runs its chained form (one persistent fit launch walking the chunks' worklists
    while the reads arrive, xp_api.cu:fit_host_impl): the rows must be those of the dense, chunk-by-chunk call record for
    record - also when a few positions belong to the 4-warp teams' class (launches of their own beside the chained one),
    and when the caller understates max_reads (the chained launch is abandoned and the call starts over)."""
    from xpore_b200.dist import unpack_records
    spec = Spec({"KO": 3, "WT": 3}, reads_lo=20, reads_hi=112, f_mod=0.3, seed=23,
                method={"prefiltering": {"method": "t-test", "threshold": 0.1}})
    batch = make_batch(spec, KmerModel.load(), n_positions=90000)
    assert batch.encode_reads() == 2
    n = np.diff(batch.read_off)
    method = xf.Method.from_dict(spec.method)
    dense = xf.fit_batch(batch, method)  # dense results: chunk-by-chunk launches
    team = (n >= 576) & dense.selected  # class split of two positions per warp (xp_api.cu:tm_class_split)
    assert 5 < team.sum() < 0.01 * batch.n_positions and n.max() < 1024
    keep = np.nonzero(dense.keep_row)[0]
    assert team[keep].any()
    # the device arena of the host calls is reused from call to call: leave another batch of the same shape in it, so that
    # anything the persistent launch read too early (a chunk's offsets or status before they had arrived) would differ
    other = batch.subset(np.random.default_rng(5).permutation(batch.n_positions))
    assert other.encode_reads() == 2
    for claim in (None, 560, 300):  # the truth; below the class split; far below
        xf.fit_rows(other, method)
        rows, n_fitted = xf.fit_rows(batch, method, max_reads=claim)
        assert n_fitted == dense.n_fitted
        idx, status, n_iter, pval, fit, stats = unpack_records(rows, xf.fit_width(6), xf.stats_width(6, 2))
        np.testing.assert_array_equal(idx, keep)
        np.testing.assert_array_equal(status, dense.status[keep])
        np.testing.assert_array_equal(n_iter, dense.n_iter[keep])
        np.testing.assert_array_equal(pval, dense.pval[keep])
        np.testing.assert_array_equal(fit, dense.fit[keep])
        np.testing.assert_array_equal(stats, dense.stats[keep])
    assert not (dense.status & xf.NOT_FITTED).any()
    # the asynchronous entry point of the multi-GPU path (rows and counts stay on the device, nothing waits on the host):
    # the same rows, with an index offset, behind work already enqueued on the caller's stream; too small a buffer
    # drops rows but reports the number needed
    import ctypes as C
    from xpore_b200 import _lib
    torch = torch_cuda
    L = _lib.lib()
    W = 4 + xf.fit_width(6) + xf.stats_width(6, 2)
    y, yfmt, yscale = xf._y_buffer(batch)
    arrs = [y, np.ascontiguousarray(batch.read_off, dtype=np.int64), np.ascontiguousarray(batch.grp_len, dtype=np.int32),
            np.ascontiguousarray(batch.kmer_mean), np.ascontiguousarray(batch.kmer_std),
            np.ascontiguousarray(batch.layout.grp_cond, dtype=np.int32)]
    ptr = lambda a: a.ctypes.data_as(C.c_void_p)
    cb = _lib.XpBatch(batch.n_positions, 6, 2, int(n.max()), batch.n_reads, *[ptr(a) for a in arrs], yfmt, 0, yscale)
    cm = method.c_struct()
    # ... and of an f64 batch, which takes the chunk-by-chunk form in both calls
    plain = make_batch(spec, KmerModel.load(), n_positions=90000)
    rows_f64, _ = xf.fit_rows(plain, method)
    yf, yfmt_f, yscale_f = xf._y_buffer(plain)
    assert yfmt_f == 0 and len(rows_f64) == len(keep)
    cb_f = _lib.XpBatch(batch.n_positions, 6, 2, int(n.max()), batch.n_reads, ptr(yf), *[ptr(a) for a in arrs[1:]], yfmt_f, 0, yscale_f)
    rows_dev = torch.full((len(keep) + 3, W), -1.0, dtype=torch.float64, device="cuda:0")
    counts = torch.zeros(2, dtype=torch.int32, device="cuda:0")
    _lib.check(L.xpb200_fit_host_rows_async(0, C.byref(cb_f), C.byref(cm), rows_dev.data_ptr(), len(keep) + 3, 0, counts.data_ptr(),
                                            C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    torch.cuda.synchronize()
    assert counts.tolist() == [len(keep), dense.n_fitted]
    np.testing.assert_array_equal(rows_dev.cpu().numpy()[:len(keep)], rows_f64)
    for cap in (len(keep) + 10, len(keep) - 7):
        rows_dev = torch.full((cap, W), -1.0, dtype=torch.float64, device="cuda:0")
        counts = torch.zeros(2, dtype=torch.int32, device="cuda:0")
        busy = torch.ones(1 << 24, device="cuda:0")
        for _ in range(8):
            busy = busy * 1.0001  # work in front of the call on the caller's stream
        _lib.check(L.xpb200_fit_host_rows_async(0, C.byref(cb), C.byref(cm), rows_dev.data_ptr(), cap, 1000, counts.data_ptr(),
                                                C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        got = rows_dev.clone()  # ordered after the call on the same stream
        torch.cuda.synchronize()
        assert counts.tolist() == [len(keep), dense.n_fitted]
        m_ = min(cap, len(keep))
        exp_rows = rows.copy()
        exp_rows[:, 0] += 1000
        np.testing.assert_array_equal(got.cpu().numpy()[:m_], exp_rows[:m_])


@pytest.mark.parametrize("grouping", ["pooled", "runs"])
def test_deep_positions_vs_oracle(torch_cuda, grouping):
    """Positions deeper than one CTA's shared memory (the reference has no cap on n, gmm.py:93-127; pooled mode keeps
    the reads of conditions that failed the count test, io.py:46-58,70-73) go to the streaming class and must not
    abort the batch: 30k- and 60k-read positions, and read counts on both sides of the boundary between the 4-warp
    teams and the streaming CTAs, beside ordinary positions - every position against the oracle, f64 and u16."""
    from xpore_b200.batch import Batch, Layout
    rng = np.random.default_rng(17)
    if grouping == "pooled":
        lay = Layout(["KO", "WT"], ["KO-r1", "KO-r2", "WT-r1", "WT-r2"], [0, 0, 1, 1], True)
        method = xf.Method(pooling=True)
    else:
        lay = Layout(["KO", "WT"], ["KO-r1", "KO-r2", "WT-r1", "WT-r2"], [0, 0, 1, 1], False)
        method = xf.Method(prefilter_method="t-test", prefilter_threshold=2.0)  # gate open: the generic t-test path runs
    G = lay.n_groups
    # boundary of the 4-warp class for this G (xp_api.cu deep_class_split): the slab of a 4-warp team holds
    # cap reads, 9 bytes each, beside its header
    from xpore_b200 import _lib
    L = _lib.lib()
    L.xpb200_deep_split.restype = C.c_int32
    L.xpb200_deep_split.argtypes = [C.c_int32]
    S = int(L.xpb200_deep_split(G))
    assert 16384 <= S <= 32768
    km, ks = 95.0, 3.0
    sizes = [400, 1500, S - 1, S, S + 1, 30000, 60000, 90, 5000]
    groups = []
    for n in sizes:
        per = [n // G + (1 if j < n % G else 0) for j in range(G)]
        g = []
        for j, m in enumerate(per):
            frac = 0.05 if lay.grp_cond[j] == 0 else 0.5
            k = int(m * frac)
            g.append(np.round(np.r_[rng.normal(km, ks, m - k), rng.normal(km + 7.0, 1.3 * ks, k)], 2))
        groups.append(g)
    y = np.concatenate([np.concatenate(g) for g in groups])
    glen = np.asarray([[len(x) for x in g] for g in groups], dtype=np.int32)
    off = np.zeros(len(sizes) + 1, dtype=np.int64)
    np.cumsum(glen.sum(axis=1), out=off[1:])
    batch = Batch(lay, y, off, glen, np.full(len(sizes), km), np.full(len(sizes), ks))
    oracle = [_oracle_fit(batch, p, method) for p in range(len(sizes))]
    ref = None
    for encode in (False, True):
        if encode:
            assert batch.encode_reads() == 2
        res = xf.fit_batch(batch, method)
        assert res.n_fitted == len(sizes) and not (res.status & xf.NOT_FITTED).any()
        for p, (f, present) in enumerate(oracle):
            assert res.n_iter[p] == f["n_iterations"], (p, sizes[p], res.n_iter[p], f["n_iterations"])
            assert bool(res.status[p] & xf.CONVERGED) == f["converged"], p
            np.testing.assert_allclose(res.mu[p], f["location"], rtol=1e-5, err_msg=str(p))
            np.testing.assert_allclose(res.alpha[p], f["alpha"], rtol=1e-5, err_msg=str(p))
            np.testing.assert_allclose(res.beta[p], f["beta"], rtol=1e-5, err_msg=str(p))
            np.testing.assert_allclose(res.N[p][present], f["N"], rtol=1e-5, atol=1e-9, err_msg=str(p))
        if method.prefilter_method:
            for p in range(len(sizes)):
                a = batch.y[off[p]:off[p + 1]]
                cond = np.repeat(lay.grp_cond, glen[p])
                pv = scipy.stats.ttest_ind(a[cond == 0], a[cond == 1]).pvalue
                np.testing.assert_allclose(res.pval[p], pv, rtol=1e-8, atol=1e-300)
        if ref is None:
            ref = res
        else:  # the encoding changes the bytes moved - and the class of the 1500-read position (u16: one warp with its
            # reads in tensor memory, f64: a 4-warp team), i.e. the order of its sums; nothing else
            np.testing.assert_array_equal(res.n_iter, ref.n_iter)
            same = np.asarray([not (1024 <= n < 1536) for n in sizes])
            np.testing.assert_array_equal(res.fit[same], ref.fit[same])
            np.testing.assert_array_equal(res.stats[same], ref.stats[same])
            np.testing.assert_allclose(res.fit, ref.fit, rtol=1e-11, atol=0)
            np.testing.assert_allclose(res.stats, ref.stats, rtol=1e-9, atol=1e-300, equal_nan=True)
    # the same (u16) batch resident on the device, and as rows
    out = xf.fit_device(xf.DeviceBatch(batch), method)
    torch_cuda.cuda.synchronize()
    dev = out.to_host()
    np.testing.assert_array_equal(dev.n_iter, res.n_iter)
    np.testing.assert_array_equal(dev.fit, res.fit)


def test_understated_max_reads_is_flagged(torch_cuda):
    """A device-pointer caller that understates batch.max_reads gets the affected positions back with
    XPB200_NOT_FITTED (n_iter = -1) instead of a wrong fit or an aborted batch; the others are fitted."""
    torch = torch_cuda
    spec = Spec({"KO": 2, "WT": 2}, reads_lo=20, reads_hi=60, f_mod=0.4, seed=52)
    batch = make_batch(spec, KmerModel.load(), n_positions=500)
    method = xf.Method()
    db = xf.DeviceBatch(batch)
    n = np.diff(batch.read_off)
    good = xf.fit_device(db, method).to_host()
    db.max_reads = int(np.sort(n)[250])  # understated: the slab is sized for the median position
    bad = xf.fit_device(db, method).to_host()
    torch.cuda.synchronize()
    cap = (db.max_reads + 1 + 7) & ~7
    too_deep = n > max(cap, 32)
    assert too_deep.sum() > 100
    assert ((bad.status[too_deep] & xf.NOT_FITTED) != 0).all() and (bad.n_iter[too_deep] == -1).all()
    ok = ~too_deep
    assert not (bad.status[ok] & xf.NOT_FITTED).any()
    np.testing.assert_array_equal(bad.fit[ok], good.fit[ok])


def test_cli_two_ranks_nccl(torch_cuda, tmp_path):
    """The drop-in CLI under torchrun with one rank per GPU (needs two GPUs): ids split by data.index byte spans,
    every rank runs its pipeline, the numeric row records are gathered on rank 0 over NCCL (dist.gather_records) -
    the table must hold exactly the rows of the single-rank run, the log every id."""
    import os
    import subprocess
    import sys

    from conftest import load_case
    from helpers import assert_same_rows, read_table
    from xpore_b200 import cli
    if torch_cuda.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for name in ("g1_runs", "g5_ragged", "g7_six_conditions"):
        (tmp_path / name).mkdir()
        case = load_case(name, tmp_path / name)
        out = os.path.join(os.path.dirname(case["config_path"]), "out")
        cli.main(["xpore", "diffmod", "--config", case["config_path"]])
        header1, rows1 = read_table(os.path.join(out, "diffmod.table"))
        os.rename(os.path.join(out, "diffmod.table"), os.path.join(out, "single.table"))
        r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                            "--master-addr", "127.0.0.1", "--master-port", "29631", "-m", "xpore_b200.cli", "diffmod",
                            "--config", case["config_path"], "--n_processes", "2"], cwd=repo, capture_output=True, text=True,
                           timeout=600)
        assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
        header2, rows2 = read_table(os.path.join(out, "diffmod.table"))
        assert header2 == header1 and len(rows1) > 0
        assert_same_rows(rows1, rows2)
        lines = open(os.path.join(out, "diffmod.log")).read().splitlines()
        assert sorted(lines[1:-1]) == sorted(case["expected"]["ids"]) and lines[-1] == "--- SUCCESSFULLY FINISHED ---"


@pytest.mark.gpu
def test_torch_custom_op_equals_fit_device(torch_cuda):
    """torch.ops.xpore_b200.fit (xpore_b200/torch_op.py) returns what fit_device returns, for f64 and u16 reads."""
    import xpore_b200.torch_op  # noqa: F401
    torch = torch_cuda
    spec = Spec({"KO": 3, "WT": 3}, reads_lo=20, reads_hi=100, f_mod=0.3, seed=77,
                method={"prefiltering": {"method": "t-test", "threshold": 0.1}})
    batch = make_batch(spec, KmerModel.load(), n_positions=4000)
    method = xf.Method.from_dict(spec.method)
    for encode in (False, True):
        if encode:
            assert batch.encode_reads() == 2
        db = xf.DeviceBatch(batch)
        ref = xf.fit_device(db, method).to_host()
        d = db.dev
        m = method.c_struct()
        status, n_iter, pval, fit, stats, n_fitted = torch.ops.xpore_b200.fit(
            d["y"], d["read_off"], d["grp_len"], d["kmer_mean"], d["kmer_std"], d["grp_cond"], db.C, db.max_reads,
            m.max_iters, m.stopping_criteria, m.prefilter, m.prefilter_threshold, m.dirichlet_conc)
        torch.cuda.synchronize()
        assert int(n_fitted.item()) == ref.n_fitted > 100
        np.testing.assert_array_equal(status.cpu().numpy().view(np.uint32), ref.status)
        np.testing.assert_array_equal(n_iter.cpu().numpy(), ref.n_iter)
        np.testing.assert_array_equal(pval.cpu().numpy(), ref.pval)
        sel = ref.selected
        np.testing.assert_array_equal(fit.cpu().numpy()[sel], ref.fit[sel])
        np.testing.assert_array_equal(stats.cpu().numpy()[sel], ref.stats[sel])
