"""`--save_models` (io.save_models_to_hdf5, io.py:99-139): the node dictionaries this package writes, against
the node values of the unmodified reference (tests/golden/*/models.json, make_golden_models.py).

CPU: the host assembly (`xpore_b200.models`) fed with the emulator's fit and the oracle's z node.
GPU (`-m gpu`): fit kernel coefficients -> xpb200_posterior_device -> the same comparison at the north-star
tolerances, and the CLI flag end to end."""
import json
import os

import numpy as np
import pytest

from oracle import xpore_oracle as orc
from xpore_b200 import models as xm
from xpore_b200 import fit as xf

from conftest import GOLDEN, load_case
from helpers import emul_fit_batch, load_emulator, pack_case

CASES = ["g1_runs", "g3_pooled"]


def _golden(name):
    return json.load(open(os.path.join(GOLDEN, name, "models.json")))


def _check_nodes(nodes, m, lay_pooling, rtol, atol_p, atol_ln):
    assert list(nodes["x"]["group_names"]) == m["x_group_names"]
    np.testing.assert_array_equal(np.argmax(nodes["x"]["data"], axis=1), m["x_group_of_read"])
    assert nodes["x"]["data"].sum(axis=1).tolist() == [1] * len(m["y_data"])
    np.testing.assert_array_equal(nodes["y"]["data"], m["y_data"])
    np.testing.assert_allclose(nodes["y"]["N"], m["y_N"], rtol=rtol, atol=1e-9)
    np.testing.assert_allclose(nodes["y"]["mean"], m["y_mean"], rtol=rtol)
    np.testing.assert_allclose(nodes["y"]["variance"], m["y_variance"], rtol=rtol)
    np.testing.assert_allclose(nodes["z"]["prob"], m["z_prob"], rtol=rtol, atol=atol_p)
    np.testing.assert_allclose(nodes["z"]["ln_prob"], m["z_ln_prob"], rtol=rtol, atol=atol_ln)
    np.testing.assert_allclose(nodes["w"]["concentration"], m["w_concentration"], rtol=rtol, atol=1e-9)
    for k in ("location", "lambda", "alpha", "beta"):
        np.testing.assert_allclose(nodes["mu_tau"][k], m["mu_tau"][k], rtol=rtol)
    assert list(nodes["y_condition_names"]["data"]) == m["y_condition_names"]
    if not lay_pooling:
        assert list(nodes["y_run_names"]["data"]) == m["y_run_names"]


def _position_of(batch, key):
    for p in range(batch.n_positions):
        if [batch.ids[p], str(batch.positions[p]), batch.kmers[p]] == key:
            return p
    raise KeyError(key)


@pytest.mark.parametrize("name", CASES)
def test_host_assembly_against_reference_nodes(name, tmp_path):
    case = load_case(name, tmp_path)
    batch, method, _ = pack_case(case["config_path"], case["expected"]["ids"])
    m = xf.Method.from_dict(method)
    res = emul_fit_batch(load_emulator(), batch, m)
    prob0 = np.full(batch.n_reads, np.nan)
    ln_prob = np.full((batch.n_reads, 2), np.nan)
    gold = _golden(name)
    for g in gold:
        p = _position_of(batch, g["key"])
        a, b = batch.read_off[p], batch.read_off[p + 1]
        y = batch.y[a:b]
        present = np.nonzero(batch.grp_len[p] > 0)[0]
        X = np.zeros((len(y), len(present)), dtype=bool)
        X[np.arange(len(y)), np.repeat(np.arange(len(present)), batch.grp_len[p][present])] = True
        f = orc.fit_position(y, X, batch.kmer_mean[p], batch.kmer_std[p], rng=np.random.RandomState(0), trace=True)
        prob0[a:b], ln_prob[a:b] = f["z_prob"][:, 0], f["z_ln_prob"]
        _check_nodes(xm.position_nodes(batch, res, p, prob0, ln_prob, m.dirichlet_conc), g, batch.layout.pooling,
                     rtol=1e-7, atol_p=1e-300, atol_ln=1e-9)
    # the writer: one file per id, flat keys <position>/nodes/<node>/<field>; unfitted reads are not needed
    prob0[np.isnan(prob0)] = 0.5
    ln_prob[np.isnan(ln_prob)] = np.log(0.5)
    paths = xm.save_models(str(tmp_path / "models"), batch, res, prob0, ln_prob, m.dirichlet_conc)
    assert sorted(os.path.basename(p) for p in paths) == sorted(i + ".npz" for i in set(batch.ids))
    z = np.load(os.path.join(str(tmp_path / "models"), gold[0]["key"][0] + ".npz"))
    pos = gold[0]["key"][1]
    assert str(z[pos + "/kmer"]) == gold[0]["kmer"]
    np.testing.assert_allclose(z[pos + "/nodes/z/prob"], gold[0]["z_prob"], rtol=1e-7, atol=1e-300)
    assert {k.split("/")[2] for k in z.files if k.startswith(pos + "/nodes/")} >= {"x", "y", "z", "w", "mu_tau",
                                                                                    "y_condition_names"}


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_gpu_z_node_and_cli_flag(name, tmp_path):
    import torch
    assert torch.cuda.is_available()
    from xpore_b200 import cli
    case = load_case(name, tmp_path)
    batch, method, _ = pack_case(case["config_path"], case["expected"]["ids"])
    m = xf.Method.from_dict(method)
    gold = _golden(name)
    for encode in (False, True):
        if encode:
            assert batch.encode_reads() == 2
        db = xf.DeviceBatch(batch)
        out = xf.fit_device(db, m, xf.DeviceResult(db.P, db.G, db.C, db.device, with_coef=True))
        prob0, ln_prob = xf.posterior_device(db, out)
        res, prob0, ln_prob = out.to_host(), prob0.cpu().numpy(), ln_prob.cpu().numpy()
        plain = xf.fit_batch(batch, m)  # keeping the coefficients must not change the fit
        np.testing.assert_array_equal(res.fit[res.selected], plain.fit[plain.selected])
        for g in gold:
            p = _position_of(batch, g["key"])
            _check_nodes(xm.position_nodes(batch, res, p, prob0, ln_prob, m.dirichlet_conc), g, batch.layout.pooling,
                         rtol=1e-5, atol_p=1e-12, atol_ln=1e-9)
    cli.main(["xpore", "diffmod", "--config", case["config_path"], "--save_models"])
    mdir = os.path.join(os.path.dirname(case["config_path"]), "out", "models")
    z = np.load(os.path.join(mdir, gold[0]["key"][0] + ".npz"))
    for g in gold:
        pos = g["key"][1]
        np.testing.assert_allclose(z[pos + "/nodes/z/prob"], g["z_prob"], rtol=1e-5, atol=1e-12)
        np.testing.assert_allclose(z[pos + "/nodes/z/ln_prob"], g["z_ln_prob"], rtol=1e-5, atol=1e-9)
        np.testing.assert_allclose(z[pos + "/nodes/y/variance"], g["y_variance"], rtol=1e-5)
        assert str(z[pos + "/kmer"]) == g["kmer"]
