"""CPU checks of the kernels' arithmetic through the host emulator (tests/emul/xp_emul.cpp links
the same xp_math.cuh / xp_core.cuh as the CUDA kernels)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest
import scipy.special
import scipy.stats

from xpore_b200 import fit as xf

from helpers import compare_with_golden, pack_case

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "emul", "xp_emul.cpp")
SO = os.path.join(HERE, "emul", "libxp_emul.so")


@pytest.fixture(scope="session")
def emul():
    deps = [SRC] + [os.path.join(HERE, "..", "xpore_b200", "csrc", f) for f in ("xp_math.cuh", "xp_core.cuh")]
    if not os.path.exists(SO) or any(os.path.getmtime(d) > os.path.getmtime(SO) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-shared", "-fPIC", "-x", "c++",
                               SRC, "-o", SO])
    L = C.CDLL(SO)
    L.emul_prefilter.restype = C.c_double
    L.emul_prefilter.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_double]
    L.emul_fit.restype = None
    L.emul_fit.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_double, C.c_double, C.c_int, C.c_double,
                           C.c_double, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.emul_stats.restype = C.c_uint
    L.emul_stats.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_double,
                             C.c_void_p, C.c_void_p]
    L.emul_special.restype = None
    L.emul_special.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong]
    return L


def special(L, which, x, aux=None):
    x = np.ascontiguousarray(x, dtype=np.float64)
    aux = np.ascontiguousarray(aux if aux is not None else np.zeros_like(x), dtype=np.float64)
    out = np.empty_like(x)
    L.emul_special(which, x.ctypes.data, aux.ctypes.data, out.ctypes.data, len(x))
    return out


def emul_fit_batch(L, batch, method: xf.Method) -> xf.FitResult:
    """Run the emulator position by position and fill the same records the C ABI returns."""
    P, G, Cn = batch.n_positions, batch.layout.n_groups, batch.layout.n_conditions
    res = xf.FitResult(G, Cn, np.zeros(P, np.uint32), np.zeros(P, np.int32), np.full(P, np.nan),
                       np.full((P, xf.fit_width(G)), np.nan), np.full((P, xf.stats_width(G, Cn)), np.nan), 0)
    gc = np.ascontiguousarray(batch.layout.grp_cond, dtype=np.int32)
    for p in range(P):
        y = np.ascontiguousarray(batch.y[batch.read_off[p]:batch.read_off[p + 1]])
        gl = np.ascontiguousarray(batch.grp_len[p])
        sel = True
        if method.prefilter_method is not None:
            pv = np.nan
            if method.prefilter_method == "t-test":
                pv = L.emul_prefilter(len(y), y.ctypes.data, G, gl.ctypes.data, gc.ctypes.data, batch.kmer_mean[p])
            res.pval[p] = pv
            sel = bool(np.isnan(pv) or pv < method.prefilter_threshold)
        if not sel:
            continue
        nit, st = C.c_int(0), C.c_uint(0)
        L.emul_fit(len(y), y.ctypes.data, G, gl.ctypes.data, batch.kmer_mean[p], batch.kmer_std[p],
                   method.max_iters, method.stopping_criteria, method.dirichlet_conc, res.fit[p].ctypes.data,
                   C.byref(nit), C.byref(st), None)
        fl = L.emul_stats(G, Cn, gc.ctypes.data, gl.ctypes.data, method.dirichlet_conc, batch.kmer_mean[p],
                          batch.kmer_std[p], res.fit[p].ctypes.data, res.stats[p].ctypes.data)
        res.n_iter[p] = nit.value
        res.status[p] = st.value | fl
        res.n_fitted += 1
    return res


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
