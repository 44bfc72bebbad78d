"""CPU checks of the kernels' arithmetic through the host emulator (tests/emul/xp_emul.cpp links
the same xp_math.cuh / xp_core.cuh as the CUDA kernels)."""
import numpy as np
import pytest
import scipy.special
import scipy.stats

from xpore_b200 import fit as xf

from helpers import compare_with_golden, pack_case
from helpers import emul_fit_batch, load_emulator


@pytest.fixture(scope="session")
def emul():
    return load_emulator()


def special(L, which, x, aux=None):
    x = np.ascontiguousarray(x, dtype=np.float64)
    aux = np.ascontiguousarray(aux if aux is not None else np.zeros_like(x), dtype=np.float64)
    out = np.empty_like(x)
    L.emul_special(which, x.ctypes.data, aux.ctypes.data, out.ctypes.data, len(x))
    return out


def test_special_functions(emul):
    x = np.concatenate([np.geomspace(1e-3, 1e5, 4000), np.linspace(0.001, 30, 3000)])
    np.testing.assert_allclose(special(emul, 0, x), scipy.special.digamma(x), rtol=2e-13, atol=2e-14)
    np.testing.assert_allclose(special(emul, 1, x), scipy.special.gammaln(x), rtol=2e-13, atol=2e-14)
    xl = np.geomspace(1e-300, 1e300, 5000)
    np.testing.assert_allclose(special(emul, 2, xl), np.log(xl), rtol=4e-16, atol=0)
    xe = np.linspace(-700, 700, 20001)
    np.testing.assert_allclose(special(emul, 3, xe), np.exp(xe), rtol=1e-15)
    d = np.concatenate([np.linspace(-60, 60, 4001), [-1e4, -600, -500.5, 500.5, 600, 1e4, 0.0, 1e-300, -1e-300]])
    r1 = special(emul, 5, d)
    np.testing.assert_allclose(r1, scipy.special.expit(np.clip(d, -500, 500)), rtol=1e-15, atol=1e-300)
    hq = special(emul, 6, d)
    rm = scipy.special.expit(-np.abs(d))  # minor responsibility (accurate); the entropy is even in d
    ent = scipy.special.xlogy(rm, rm) + (1 - rm) * np.log1p(-rm)  # r ln r + (1-r) ln(1-r)
    np.testing.assert_allclose(hq, ent, rtol=1e-12, atol=1e-15)


def test_student_t(emul):
    rng = np.random.default_rng(0)
    t = np.concatenate([rng.normal(0, 3, 4000), rng.normal(0, 30, 500), [0.0, 1e-200, 1e3]])
    df = np.concatenate([rng.integers(2, 2000, 4500), [10, 10, 10]]).astype(np.float64)
    ref = scipy.stats.t.sf(np.abs(t), df) * 2
    got = special(emul, 4, t, df)
    ok = ref > 1e-290
    np.testing.assert_allclose(got[ok], ref[ok], rtol=5e-12)
    assert (got[~ok] < 1e-280).all()


def test_emulator_matches_reference(emul, golden_case):
    exp = golden_case["expected"]
    batch, method, _ = pack_case(golden_case["config_path"], exp["ids"])
    res = emul_fit_batch(emul, batch, xf.Method.from_dict(method))
    compare_with_golden(batch, res, method, exp, rtol_param=1e-8, rtol_p=1e-9)


def test_closed_form_elbo_equals_the_nine_term_sum(emul):
    """The kernels evaluate a component's ELBO share in closed form after the first sweep
    (xp_core.cuh:elbo_component_closed).  Sweep by sweep the emulator's ELBO must equal the oracle's, which
    adds the reference's nine terms (gmm.py:77-91) in the reference's order."""
    import ctypes as C

    from oracle import xpore_oracle as orc
    from xpore_b200.kmer_model import KmerModel
    from xpore_b200.synth import Spec, make_batch

    spec = Spec({"KO": 2, "WT": 3}, reads_lo=5, reads_hi=80, f_mod=0.5, seed=11)
    batch = make_batch(spec, KmerModel.load(), n_positions=60)
    G = batch.layout.n_groups
    checked = 0
    for p in range(batch.n_positions):
        y = np.ascontiguousarray(batch.y[batch.read_off[p]:batch.read_off[p + 1]])
        gl = np.ascontiguousarray(batch.grp_len[p])
        X = np.zeros((len(y), G), dtype=bool)
        X[np.arange(len(y)), np.repeat(np.arange(G), gl)] = True
        ref = orc.fit_position(y, X, batch.kmer_mean[p], batch.kmer_std[p], rng=np.random.RandomState(0), trace=True)
        fit = np.empty(xf.fit_width(G))
        elbos = np.full(500, np.nan)
        nit, st = C.c_int(0), C.c_uint(0)
        emul.emul_fit(len(y), y.ctypes.data, G, gl.ctypes.data, batch.kmer_mean[p], batch.kmer_std[p], 500, 1e-5, xf.Method().dirichlet_conc,
                      fit.ctypes.data, C.byref(nit), C.byref(st), elbos.ctypes.data)
        assert nit.value == ref["n_iterations"]
        if st.value & xf.ELBO_DECREASED:
            continue  # the reference breaks before recording the decreased value
        want = np.array(ref["elbos"][1:] + [ref["elbo_last"]])  # ELBO after sweep 1, 2, ... (elbos[0] is the random start)
        np.testing.assert_allclose(elbos[:nit.value], want, rtol=2e-12)
        checked += nit.value
    assert checked > 1000
