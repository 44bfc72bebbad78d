"""CPU oracle for the xPore diffmod hot path — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A numpy/scipy restatement of the reference algorithm, used only by ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs as the checker / the timed CPU
arm.  The product path (``xpore_b200``) never imports this module.

Pinned: ``tests/test_oracle_golden.py`` checks every function here against outputs of the
UNMODIFIED reference captured live by ``tests/golden/make_golden.py`` (the reference ships no
tests or golden vectors of its own).

Third-party arithmetic the reference relies on and that is therefore used here too (not
vendored in /root/reference; ``setup.py:24-26`` lower-bounds only): numpy (installed 2.3.5:
``np.percentile`` linear method, pairwise ``np.sum``), scipy (installed 1.18.1:
``scipy.special.{digamma,gammaln,logsumexp}``, ``scipy.stats.{ttest_ind,t.sf,norm.sf,norm.cdf}``),
CPython ``math.{erf,sqrt,log,fabs}``.

All citations are ``/root/reference/`` paths.  Arithmetic is float64 and follows the reference's
operation order wherever that can change a bit, so the restatement reproduces the reference's
numbers exactly on the golden fixtures; the structure (plain functions over arrays instead of a
graph of node classes) is this repository's own.
"""
from __future__ import annotations

import itertools
import math
from collections import OrderedDict

import numpy as np
import scipy.special
import scipy.stats

LOG_2PI = np.log(2 * np.pi)


# --------------------------------------------------------------------------------------------
# A3  position assembly / selection                                   xpore/diffmod/io.py:12-90
# --------------------------------------------------------------------------------------------
def one_hot(labels):
    """``get_dummies`` (io.py:12-18): columns in sorted label order."""
    names = sorted(set(labels))
    return np.array([labels == nm for nm in names]).T, names


def select_positions(idx, data_dict, min_count, max_count, pooling=False):
    """``load_data`` (io.py:21-90).  Returns ``OrderedDict{(idx,pos,kmer): {...}}`` sorted by
    (int(pos), kmer) — the reference iterates a ``set`` (arbitrary order)."""
    pair_lists = []
    for (_cond, _run), d in data_dict.items():
        if d is not None:  # io.py:33
            pair_lists.append([(pos, kmer) for pos in d[idx] for kmer in d[idx][pos]])
    common = set(pair_lists[0]).intersection(*pair_lists)  # io.py:39

    def order(pk):
        try:
            return (0, int(pk[0]), pk[1])
        except ValueError:
            return (1, 0, pk[0] + pk[1])

    out = OrderedDict()
    cond_names_all = {c for c, _ in data_dict.keys()}
    for pos, kmer in sorted(common, key=order):
        y, cond_labels, run_labels = [], [], []
        counts = {}
        for (cond, run), d in data_dict.items():
            if d is None:
                continue
            reads = d[idx][pos][kmer]
            n = len(reads)
            if (not pooling) and (n < min_count or n > max_count):  # io.py:51-52
                continue
            counts.setdefault(cond, []).append(n)
            y += reads
            cond_labels += [cond] * n
            run_labels += [run] * n
        if len(y) == 0:  # io.py:66-67
            continue
        included = 0
        for cond in cond_names_all:
            c = np.array(counts.get(cond, []))
            if pooling:  # io.py:70-73
                included += int(min_count <= sum(counts.get(cond, [])) <= max_count)
            else:  # io.py:75-77
                included += int((c >= min_count).any() and (c <= max_count).any())
        if included < 2:  # io.py:79-80
            continue
        y = np.array(y)
        cond_labels, run_labels = np.array(cond_labels), np.array(run_labels)
        x, cnames = one_hot(cond_labels)
        r, rnames = one_hot(run_labels)
        out[(idx, pos, kmer)] = {"y": y, "x": x, "r": r, "condition_names": cnames, "run_names": rnames}
    return out


# --------------------------------------------------------------------------------------------
# T1-T3  prefilter                                             xpore/diffmod/statstest.py:4-21
# --------------------------------------------------------------------------------------------
def prefilter_pvalue(y, x, method="t-test"):
    """Two-sample pooled-variance t-test between the two condition columns; NaN unless exactly
    two condition columns and ``method == 't-test'`` (statstest.py:5-21)."""
    if x.shape[1] != 2 or method != "t-test":
        return np.nan
    a, b = y[x[:, 0] == 1], y[x[:, 1] == 1]
    t, _ = scipy.stats.ttest_ind(a, b)
    return scipy.stats.t.sf(np.abs(t), len(a) + len(b) - 2) * 2


# --------------------------------------------------------------------------------------------
# A4  priors                      xpore/scripts/diffmod.py:23-48, diffmod/configurator.py:62-77
# --------------------------------------------------------------------------------------------
def make_priors(kmer_mean, kmer_std, n_groups):
    tau = 1.0 / (kmer_std ** 2)
    return {
        "location": np.array([kmer_mean, kmer_mean]),
        "lambda": np.array([1, 1]),
        "alpha": np.array([tau, tau]),
        "beta": np.array([0.5 * 1.0 / tau, 0.5 * 1.0 / tau]),
        "conc": np.full((n_groups, 2), 0.001),
    }


# --------------------------------------------------------------------------------------------
# G1-G7  variational EM                                           xpore/diffmod/gmm.py:11-370
# --------------------------------------------------------------------------------------------
def _e_log_w(conc):  # gmm.py:283-284
    res = scipy.special.digamma(conc)
    res -= scipy.special.digamma(np.sum(conc, axis=-1))[..., None]
    return res


def _log_c(conc):  # gmm.py:296-297
    return scipy.special.gammaln(np.sum(conc, axis=-1)) - np.sum(scipy.special.gammaln(conc), axis=-1)


def _quad(y, q):
    """ln-density core shared by the E-step and the likelihood term (gmm.py:164,236)."""
    m, v = q["location"], 1.0 / (q["lambda"] * (q["alpha"] / q["beta"]))  # gmm.py:356,360
    e_tau = q["alpha"] / q["beta"]
    e_log_tau = scipy.special.digamma(q["alpha"]) - np.log(q["beta"])  # gmm.py:350
    yc = y[:, None]
    return 0.5 * (e_log_tau - np.log(2 * np.pi) - e_tau * (yc ** 2 - 2 * yc * m + m ** 2 + v))


def _normal_gamma_terms(q, prior):
    """E[ln p(mu|tau)] + E[ln p(tau)] - E[ln q(mu|tau)] - E[ln q(tau)]  (gmm.py:329-345)."""
    e_tau = q["alpha"] / q["beta"]
    e_log_tau = scipy.special.digamma(q["alpha"]) - np.log(q["beta"])
    m = q["location"]
    v = 1.0 / (q["lambda"] * e_tau)

    def normal(p):
        return np.sum(0.5 * (e_log_tau + np.log(p["lambda"]) - np.log(2 * np.pi)
                             - p["lambda"] * e_tau * (m ** 2 + v - 2 * m * p["location"] + p["location"] ** 2)))

    def gamma(p):
        return np.sum(p["alpha"] * np.log(p["beta"]) - scipy.special.gammaln(p["alpha"])
                      + (p["alpha"] - 1) * e_log_tau - p["beta"] * e_tau)

    return normal(prior), gamma(prior), normal(q), gamma(q)


def elbo(y, X, prob, ln_prob, q, conc, prior):
    """Nine-term ELBO in the reference's accumulation order (gmm.py:77-91)."""
    has_group = (np.sum(X, axis=-1)[..., None] > 0)
    e_log_w = _e_log_w(conc)
    like = prob * _quad(y, q)  # gmm.py:164
    like *= has_group  # gmm.py:165
    val = np.sum(np.sum(like, axis=-1), axis=-1)
    val += np.sum(np.sum(prob * np.sum(X[..., None] * e_log_w[None, :, :], axis=-2), axis=-1), axis=-1)  # :214
    val -= np.sum(np.sum(prob * np.sum(X[:, None, :] * ln_prob[..., None], axis=-1), axis=-1), axis=-1)  # :218
    e_log_w_prior = e_log_w
    val += np.sum(_log_c(prior["conc"]) + np.sum((prior["conc"] - 1) * e_log_w_prior, axis=-1), axis=-1)  # :274
    val -= np.sum(_log_c(conc) + np.sum((conc - 1) * e_log_w, axis=-1), axis=-1)  # :278
    pn, pg, qn, qg = _normal_gamma_terms(q, prior)
    val += pn
    val += pg
    val -= qn
    val -= qg
    return val


def initial_locations(y, loc0):
    """gmm.py:39-44."""
    if loc0 < np.percentile(y, q=50):
        loc1 = np.percentile(y, q=75)
    else:
        loc1 = np.percentile(y, q=25)
    return np.array([loc0, loc1])


def fit_position(y, X, kmer_mean, kmer_std, max_iters=500, stopping_criteria=1e-5, rng=None, trace=False):
    """``GMM(...).fit()`` (gmm.py:11-127).  ``X`` is the one-hot group matrix (runs, or
    conditions when pooling).  Returns the posterior parameters the table reads."""
    G = X.shape[1]
    prior = make_priors(kmer_mean, kmer_std, G)
    q = {"location": initial_locations(y, prior["location"][0]), "lambda": prior["lambda"],
         "alpha": prior["alpha"], "beta": prior["beta"]}  # gmm.py:50
    conc = prior["conc"].copy()  # gmm.py:261-263
    n = len(y)
    rng = rng or np.random
    prob = rng.random((n, 2))  # gmm.py:207-209 (unseeded in the reference; only feeds ELBO_0)
    prob /= np.sum(prob, axis=1)[:, None]
    ln_prob = np.log(prob)
    elbo_old = elbo(y, X, prob, ln_prob, q, conc, prior)  # gmm.py:100
    elbos = []
    converged = False
    diff = -1
    iteration = 0
    N = None
    for iteration in range(1, max_iters):  # gmm.py:103  (at most max_iters-1 sweeps)
        # E-step (gmm.py:227-243)
        ln_rho = _quad(y, q)
        ln_rho += np.sum(X[..., None] * _e_log_w(conc)[None, :, :], axis=-2)
        d = ln_rho[:, 1] - ln_rho[:, 0]
        p0 = 1.0 / (1.0 + np.exp(np.clip(d, -500, 500)))
        prob = np.stack([p0, 1 - p0], axis=-1)
        ln_prob = ln_rho - scipy.special.logsumexp(ln_rho, axis=-1)[..., None]
        # sufficient statistics (gmm.py:177-189)
        N = np.sum(X[..., None] * prob[:, None, :], axis=-3)
        Nk = np.sum(N, axis=-2)
        Nk[Nk < 1e-30] = 0.0
        Ninv = np.divide(1.0, Nk, out=np.zeros_like(Nk), where=Nk != 0.0)
        yc = y[:, None]
        mean = Ninv * np.sum(prob * yc, axis=-2)
        resid = (np.sum(X, axis=-1) > 0)[..., None] * (yc - mean[None, :])
        var = Ninv * np.sum(prob * (resid ** 2), axis=-2)
        # w (gmm.py:292-294) and mu,tau (gmm.py:364-370)
        conc = prior["conc"] + N
        Ns = np.sum(N, axis=-2)
        lam = prior["lambda"] + Ns
        q = {"lambda": lam,
             "location": (1.0 / lam) * (Ns * mean + prior["lambda"] * prior["location"]),
             "alpha": prior["alpha"] + 0.5 * Ns,
             "beta": prior["beta"] + 0.5 * (Ns * var + ((prior["lambda"] * Ns) / (prior["lambda"] + Ns))
                                            * (mean - prior["location"]) ** 2)}
        elbo_new = elbo(y, X, prob, ln_prob, q, conc, prior)
        if np.round(elbo_new - elbo_old, 8) < 0:  # gmm.py:110-115
            break
        elbos.append(elbo_old)
        diff = (elbo_new - elbo_old) / np.abs(elbo_old)
        elbo_old = elbo_new
        if diff < stopping_criteria:  # gmm.py:120-122
            converged = True
            break
    out = {"location": q["location"], "lambda": q["lambda"], "alpha": q["alpha"], "beta": q["beta"],
           "conc": conc, "N": N, "n_iterations": iteration, "converged": converged,
           "elbo_prev": elbos[-1] if elbos else None, "n_elbos": len(elbos), "convergence_ratio": diff,
           "elbo_last": elbo_old}
    if trace:
        out["elbos"] = elbos
        # the z and y nodes as io.save_models_to_hdf5 (io.py:125-137) would write them: responsibilities of the
        # last E-step (gmm.py:241-243) and the sufficient statistics taken from them (gmm.py:179-189)
        out["z_prob"], out["z_ln_prob"] = prob, ln_prob
        out["y_mean"], out["y_variance"] = mean, var
    return out


# --------------------------------------------------------------------------------------------
# S1-S6  statistics and row emission      xpore/diffmod/io.py:230-374, xpore/utils/stats.py:7-66
# --------------------------------------------------------------------------------------------
def z_test(w1, w2, n1, n2):  # stats.py:7-15
    p1, p2 = w1.mean(), w2.mean()
    n1, n2 = n1.mean().round(), n2.mean().round()
    se = np.sqrt(p1 * (1 - p1) / n1 + p2 * (1 - p2) / n2)
    z = (p1 - p2) / se
    return z, scipy.stats.norm.sf(abs(z)) * 2


def _ncdf(x, mu, sigma):  # stats.py:35-36
    return 0.5 * (1.0 + math.erf((x - mu) / (sigma * math.sqrt(2.0))))


def overlap(means, variances):
    """``calc_prob_overlapping`` + ``NormalDist.overlap`` (stats.py:18-66)."""
    s1, s2 = np.sqrt(variances)
    m1, m2 = means
    if not (s1 != s2 and m1 != m2):
        return np.nan, [np.nan] * 4
    X, Y = (m1, s1), (m2, s2)
    if (Y[1], Y[0]) < (X[1], X[0]):
        X, Y = Y, X
    xv, yv = X[1] ** 2.0, Y[1] ** 2.0
    dv = yv - xv
    dm = math.fabs(Y[0] - X[0])
    if not dv:  # unreachable given s1 != s2, kept for fidelity (stats.py:53-54)
        return 1.0 - math.erf(dm / (2.0 * X[1] * math.sqrt(2.0)))
    a = X[0] * yv - Y[0] * xv
    b = X[1] * Y[1] * math.sqrt(dm ** 2.0 + dv * math.log(yv / xv))
    x1, x2 = (a + b) / dv, (a - b) / dv
    cx1, cy1, cx2, cy2 = _ncdf(x1, *X), _ncdf(x1, *Y), _ncdf(x2, *X), _ncdf(x2, *Y)
    return 1.0 - (math.fabs(cy1 - cx1) + math.fabs(cy2 - cx2)), [cx1, cy1, cx2, cy2]


def cluster_confidence(mu, kmer_mean, kmer_std):  # io.py:372-374
    return scipy.stats.norm.cdf(kmer_mean - abs(kmer_mean - mu), loc=kmer_mean, scale=kmer_std) * 2


def result_row(key, fit, group_names, kmer_mean, kmer_std, data_info, pooling, prefilter=None):
    """One ``diffmod.table`` row and its keep flag (io.py:255-367)."""
    condition_names = sorted(data_info.keys())
    run_names = sum([list(r.keys()) for r in data_info.values()], [])
    mu = fit["location"]
    sigma2 = 1.0 / (fit["alpha"] / fit["beta"])
    w = fit["conc"] / np.sum(fit["conc"], axis=-1)[..., None]  # gmm.py:287
    coverage = np.sum(fit["N"], axis=-1)
    p_ov, cdfs = overlap(mu, sigma2)
    conf = [cluster_confidence(mu[0], kmer_mean, kmer_std), cluster_confidence(mu[1], kmer_mean, kmer_std)]
    unmod, mod = (0, 1) if conf[0] > conf[1] else (1, 0)  # io.py:277-282
    w_mod = w[:, mod]
    assignment = ["higher", "lower"][int(mu[0] < mu[1]) ^ mod]  # io.py:288
    gnames = np.array(group_names)

    def members(cond):
        return [cond] if pooling else list(data_info[cond].keys())

    def test(sel1, sel2):
        z, p = z_test(w_mod[sel1], w_mod[sel2], coverage[sel1], coverage[sel2])
        return [np.mean(w_mod[sel1]) - np.mean(w_mod[sel2]), p, z]

    row = list(key)
    for c1, c2 in itertools.combinations(condition_names, 2):  # io.py:293-309
        s1, s2 = np.isin(gnames, members(c1)), np.isin(gnames, members(c2))
        row += test(s1, s2) if (s1.any() and s2.any()) else [None, None, None]
    if len(condition_names) > 2:  # io.py:311-330
        for c in condition_names:
            s1 = np.isin(gnames, members(c))
            row += test(s1, ~s1) if s1.any() else [None, None, None]
    names = condition_names if pooling else run_names
    row += [w_mod[gnames == nm][0] if nm in group_names else None for nm in names]  # io.py:337-344
    row += [coverage[gnames == nm][0] if nm in group_names else None for nm in names]
    row += [mu[unmod], mu[mod], sigma2[unmod], sigma2[mod], conf[unmod], conf[mod], assignment]
    if prefilter is not None:
        row += [prefilter]
    cx1, cy1, cx2, cy2 = cdfs
    t = 0.1
    not_inside = ((cy1 < t) & (cx1 < t)) | ((cy2 < t) & (cx2 < t)) | (((1 - cy1) < t) & ((1 - cx1) < t)) \
        | (((1 - cy2) < t) & ((1 - cx2) < t))  # io.py:365
    keep = bool((p_ov <= 0.5) and not_inside)
    return row, keep, {"p_overlap": p_ov, "cdfs": cdfs, "conf": conf, "mod": mod, "w": w, "coverage": coverage}


def table_header(data_info, pooling, prefilter_method=None):
    """``get_result_table_header`` (io.py:189-228)."""
    condition_names = sorted(data_info.keys())
    run_names = sum([list(r.keys()) for r in data_info.values()], [])
    head = ["id", "position", "kmer"]
    for a, b in itertools.combinations(condition_names, 2):
        head += ["%s_%s_vs_%s" % (s, a, b) for s in ("diff_mod_rate", "pval", "z_score")]
    if len(condition_names) > 2:
        for c in condition_names:
            head += ["%s_%s_vs_all" % (s, c) for s in ("diff_mod_rate", "pval", "z_score")]
    names = condition_names if pooling else run_names
    head += ["mod_rate_%s" % nm for nm in names] + ["coverage_%s" % nm for nm in names]
    head += ["mu_unmod", "mu_mod", "sigma2_unmod", "sigma2_mod", "conf_mu_unmod", "conf_mu_mod", "mod_assignment"]
    if prefilter_method:
        head += [prefilter_method]
    return head


# --------------------------------------------------------------------------------------------
# L2 worker: one gene                                          xpore/scripts/diffmod.py:15-77
# --------------------------------------------------------------------------------------------
def run_gene(idx, data_dict, data_info, method, criteria, kmer_lookup, rng=None):
    """``execute`` without the file writes: returns (fits keyed by position key, table rows).
    ``kmer_lookup(kmer) -> (mean, std)``."""
    pooling = bool(method.get("pooling", False))
    pre = method.get("prefiltering", False)
    data = select_positions(idx, data_dict, criteria["readcount_min"], criteria["readcount_max"], pooling)
    fits, rows = OrderedDict(), []
    for key, d in data.items():
        km, ks = kmer_lookup(key[2])
        pval = None
        if pre:
            pval = prefilter_pvalue(d["y"], d["x"], pre["method"])
            if not (np.isnan(pval) | (pval < pre["threshold"])):  # diffmod.py:54
                continue
        X = d["x"] if pooling else d["r"]
        names = d["condition_names"] if pooling else d["run_names"]
        fit = fit_position(d["y"], X, km, ks, method.get("max_iters", 500),
                           method.get("stopping_criteria", 1e-5), rng=rng)
        fit["group_names"], fit["prefilter"], fit["kmer_mean"], fit["kmer_std"] = names, pval, km, ks
        fit["n_reads"] = len(d["y"])
        fits[key] = fit
        row, keep, _ = result_row(key, fit, names, km, ks, data_info, pooling, pval)
        if keep:
            rows.append(row)
    return data, fits, rows
