// xp_core.cuh — per-position arithmetic of the diffmod path, shared by the CUDA kernels and by
// the host emulator the CPU tests link (tests/emul).  No parallel orchestration in here.
//
// Coordinates: reads are centred on the unmodified k-mer mean, y' = y - m0 (exact for doubles of
// the same binade), so the prior location is 0 and every sufficient statistic is a plain moment
//   N_gk = sum_n r_nk [n in g],  S1_k = sum_n r_nk y'_n,  S2_k = sum_n r_nk y'_n^2.
// The reference's two-pass mean/variance (gmm.py:177-189) and its M-step (gmm.py:364-370)
// reduce to   lambda = 1+N,  m' = S1/lambda,  alpha = a0+N/2,  beta = b0 + (S2 - S1^2/lambda)/2,
// because N s^2 + N/(1+N) ybar'^2 = S2 - S1^2/(1+N).  (lambda0 = 1, configurator.py:67.)
#pragma once
#include "xp_math.cuh"

namespace xpc {

using namespace xpm;

// ---- layout of the per-position result records (f64) ------------------------------------
// fit record:  mu0 mu1 lam0 lam1 alpha0 alpha1 beta0 beta1 elbo N[g][k]...
constexpr int FIT_MU = 0, FIT_LAM = 2, FIT_ALPHA = 4, FIT_BETA = 6, FIT_ELBO = 8, FIT_N = 9;
XP_HD int fit_width(int G) { return FIT_N + 2 * G; }
// stats record: mu_unmod mu_mod s2_unmod s2_mod conf_unmod conf_mod p_overlap cdf[4]
//               mod_rate[G] coverage[G] pairwise[npairs][3] one_vs_all[C][3] (only if C > 2)
constexpr int ST_MU = 0, ST_S2 = 2, ST_CONF = 4, ST_POV = 6, ST_CDF = 7, ST_RATE = 11;
XP_HD int stats_width(int G, int C) { return ST_RATE + 2 * G + 3 * (C * (C - 1) / 2) + (C > 2 ? 3 * C : 0); }

// status bits
constexpr uint32_t STATUS_SELECTED = 1u;       // passed the prefilter gate (diffmod.py:54) => fitted
constexpr uint32_t STATUS_CONVERGED = 2u;      // gmm.py:120-122
constexpr uint32_t STATUS_ELBO_DECREASED = 4u; // gmm.py:110-115 (reference prints ERROR and breaks)
constexpr uint32_t STATUS_HIT_MAX = 8u;        // loop ran max_iters-1 sweeps without stopping
constexpr uint32_t STATUS_KEEP_ROW = 16u;      // row filter io.py:363-367
constexpr uint32_t STATUS_MOD_IS_1 = 32u;      // modified cluster is component 1 (io.py:277-282)
constexpr uint32_t STATUS_HIGHER = 64u;        // mod_assignment == 'higher' (io.py:288)
constexpr uint32_t STATUS_TTEST_VALID = 128u;  // exactly two conditions present (statstest.py:20-21)
constexpr uint32_t STATUS_NOT_FITTED = 256u;   // selected but not fitted: the caller understated batch.max_reads (n_iter = -1)
constexpr uint32_t STATUS_NEAR_THRESHOLD = 512u;  // prefilter p-value within 1e-9 (relative) of the threshold: see xp_prefilter_kernel

struct Comp {
    double mp, lam, alpha, beta;  // centred location m', lambda, alpha, beta of q(mu_k, tau_k)
    double il;                    // 1/lambda
};

// ---- E-step for one read (gmm.py:236-243) --------------------------------------------------
// d = ln rho_1 - ln rho_0.  The reference forms r0 = 1/(1+exp(clip(d,+-500))), r1 = 1-r0
// (gmm.py:240-242); here t = exp(-|d|), inv = 1/(1+t) is the responsibility of the likelier
// component and rmin = t inv that of the other one (neither loses relative precision);
// `pos` (d > 0) says component 1 is the likelier one.  h = -(r0 ln r0 + r1 ln r1) = rmin |d| + log1p(t)
// is this read's share of -E[ln q(z)] (gmm.py:217-219).
//
// Table-driven: exp(x) = 2^m 2^(j/256) e^r with |r| <= ln2/512 (degree-4 series), and
// ln(1+t) = ln1p(u/c_i - 1) + ln c_i with u = 1+t in [1,2] and c_i the centre of its 1/128-wide
// interval (degree-5 series, |u/c_i - 1| <= 1/256; entry 128 = {1/2, ln 2} serves u == 2).  `Tab`
// supplies the tables: constant arrays here, shared-memory copies in xp_fit.cuh
// (xpk::estep_dev is the same arithmetic).
struct HostTab {
    XP_HD double exp2(int j) const { return XPT(kExp2)[j]; }
    XP_HD void logc(int i, double& ic, double& lc) const {
        ic = XPT(kInvC)[i];
        lc = XPT(kLogC)[i];
    }
};

template <class Tab>
XP_HD void estep_read(const Tab& tab, double d, double& inv, double& rmin, double& h, bool& pos) {
    const int dhi = hi32(d);
    int ahi = dhi & 0x7fffffff;
    ahi = ahi < 0x407f4000 ? ahi : 0x407f4000;  // |d| clipped at 500 (low word kept: [500, 500 + 2^-12))
    const double ax = mk64(ahi, lo32(d));
    const double magic = XPT(kMisc)[5];
    const double t0 = fma(-ax, XPT(kE256)[0], magic);
    const double kf = t0 - magic;
    const int k = lo32(t0);  // low word of the magic sum = rint(-256 |d| / ln 2), two's complement
    double r = fma(-kf, XPT(kE256)[1], -ax);
    r = fma(-kf, XPT(kE256)[2], r);
    double p = fma(kExpC4, r, XPT(kE3)[0]);
    p = fma(p, r, 0.5);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    double t = tab.exp2(k & 255) * p;
    t = mk64(hi32(t) + (k >> 8) * (1 << 20), lo32(t));  // * 2^m, result stays normal (|d| <= 500)
    const double u = 1.0 + t;                           // [1, 2]
    const int i = (hi32(u) - 0x3ff00000) >> 13;         // top seven mantissa bits; u == 2 gives 128
    double ic, lc;
    tab.logc(i, ic, lc);
    const double rr = fma(u, ic, -1.0);
    double q = fma(kLogC5, rr, -0.25);
    q = fma(q, rr, XPT(kE3)[1]);
    q = fma(q, rr, -0.5);
    q = fma(q, rr, 1.0);
    const double l1p = fma(q, rr, lc);  // ln(1+t)
    inv = xrcp3(u);
    rmin = t * inv;
    pos = dhi >= 0;  // d == +-0 gives inv == rmin == 1/2 either way
    h = fma(rmin, ax, l1p);  // |d| > 500 only when rmin < 1e-217: the clip does not show
}

// The z node as the reference stores it (gmm.py:241-243): prob0 = q(z=0) and ln q(z=k).  ln q of the likelier
// component is -log1p(exp(-|d|)), of the other one -|d| - log1p(exp(-|d|)); |d| is NOT clipped there
// (gmm.py:243 takes ln_rho - logsumexp), only inside the sigmoid (gmm.py:240-241).
template <class Tab>
XP_HD void posterior_read(const Tab& tab, double d, double& prob0, double& lnp0, double& lnp1) {
    double inv, rmin, h;
    bool pos;
    estep_read(tab, d, inv, rmin, h, pos);
    const double ad = fabs(d);
    const double adc = ad < 500.0 ? ad : 500.0;
    const double l1p = h - rmin * adc;  // h = rmin |d|_clipped + log1p(t)
    prob0 = pos ? rmin : inv;
    const double ln_big = -l1p, ln_small = -ad - l1p;
    lnp0 = pos ? ln_small : ln_big;
    lnp1 = pos ? ln_big : ln_small;
}

// ---- psi / lgamma of one posterior parameter inside the sweeps (xp_math.cuh: shift10, xlog_tab) -------
// Returns psi(x) and lgamma(x); den is the shift product (1 for x >= 10), whose logarithm comes from the same table.
template <class Tab>
XP_HD void psi_lgamma_parts(const Tab& tab, double x, double* psi, double* lgam) {
    double z, num, den;
    shift10(x, z, num, den);
    const double lz = xlog_tab(tab, z);
    const double lden = xlog_tab(tab, den);
    const double q = xrcp3(z * den);
    double lg;
    psi_lgamma_shifted(z, lz, q * den, num * (q * z), psi, &lg);
    *lgam = lg - lden;
}

// ---- M-step for one component (gmm.py:364-370 in centred moments) ------------------------
XP_HD Comp mstep(double Nk, double S1, double S2, double a0, double b0) {
    Comp c;
    c.lam = 1.0 + Nk;
    c.il = xdiv3(1.0, c.lam);
    c.mp = S1 * c.il;
    c.alpha = a0 + 0.5 * Nk;
    c.beta = b0 + 0.5 * (S2 - S1 * S1 * c.il);
    return c;
}

// ---- ELBO share of one component (gmm.py:156-166 and :329-345) ----------------------------
//   E[ln p(y|z)]_k + E[ln p(mu_k|tau_k)] + E[ln p(tau_k)] - E[ln q(mu_k|tau_k)] - E[ln q(tau_k)]
// psi_a = psi(alpha), lg_a = lgamma(alpha), ln_b = ln beta, ln_lam = ln lambda,
// prior_gamma = a0 ln b0 - lgamma(a0).  Also returns E[tau] and E[ln tau] for the next E-step.
XP_HD double elbo_component(const Comp& c, double Nk, double S1, double S2, double a0, double b0,
                            double prior_gamma, double psi_a, double lg_a, double ln_b, double ln_lam,
                            double& tbar, double& ltau) {
    tbar = xdiv3(c.alpha, c.beta);
    ltau = psi_a - ln_b;
    const double il = c.il;
    const double m2 = c.mp * c.mp;
    const double quad = fma(m2, Nk, fma(-2.0 * c.mp, S1, S2));  // sum_n r_nk (y'-m')^2
    const double t1 = 0.5 * Nk * (ltau - kLn2Pi - il) - 0.5 * tbar * quad;
    const double t6 = 0.5 * (ltau - kLn2Pi - tbar * m2 - il);
    const double t7 = prior_gamma + (a0 - 1.0) * ltau - b0 * tbar;
    const double t8 = 0.5 * (ltau + ln_lam - kLn2Pi - 1.0);
    const double t9 = c.alpha * ln_b - lg_a + (c.alpha - 1.0) * ltau - c.alpha;
    return t1 + t6 + t7 - t8 - t9;
}

// The same share once the component's posterior is the M-step of the very statistics the ELBO is evaluated
// on (every sweep but the first, whose locations are the percentile start): with alpha = a0 + N/2,
// lambda = 1 + N, m' = S1/lambda, beta = b0 + (S2 - S1 m')/2 the E[ln tau] terms cancel
// (N/2 + a0 - alpha = 0), the quadratic terms sum to E[tau] beta = alpha, and what is left is the ratio of the
// Normal-Gamma normalisers:
//   lgamma(alpha) - alpha ln beta - ln(lambda)/2  +  [a0 ln b0 - lgamma(a0)]  -  (N/2) ln 2 pi
// The last two terms are per-position constants (sum_k N_k = n): the caller adds elbo_closed_const once.
XP_HD double elbo_component_closed(const Comp& c, double psi_a, double lg_a, double ln_b, double ln_lam,
                                   double& tbar, double& ltau) {
    tbar = xdiv3(c.alpha, c.beta);
    ltau = psi_a - ln_b;
    return lg_a - fma(c.alpha, ln_b, 0.5 * ln_lam);
}
XP_HD double elbo_closed_const(double prior_gamma, int n) { return fma(-0.5 * kLn2Pi, (double)n, 2.0 * prior_gamma); }

// Coefficients of d_n = A_g + y'(B + C y') from the two components (gmm.py:236-238 expanded):
//   A_g = base + (E ln w_g1 - E ln w_g0)
XP_HD void estep_coefficients(const Comp& c0, const Comp& c1, double tb0, double tb1, double lt0, double lt1,
                              double& base, double& B, double& C) {
    const double il0 = c0.il, il1 = c1.il;
    base = 0.5 * (lt1 - lt0) - 0.5 * (il1 - il0) - 0.5 * (tb1 * c1.mp * c1.mp - tb0 * c0.mp * c0.mp);
    B = tb1 * c1.mp - tb0 * c0.mp;
    C = -0.5 * (tb1 - tb0);
}

// ---- stopping rule (gmm.py:110-123) -------------------------------------------------------
// returns 0 = continue, 1 = converged, 2 = ELBO decreased.  np.round(x, 8) < 0  <=>  rint(x*1e8) < 0
// (np.round multiplies by 10^8, rounds half-even, divides).
XP_HD int stop_rule(double elbo_new, double elbo_old, double stopping) {
    const double delta = elbo_new - elbo_old;
    if (rint(delta * 1e8) < 0.0) return 2;
    if (elbo_old == -INFINITY) return 0;  // first sweep: the reference compares with the ELBO of a random q(z)
    const double diff = xdiv3(delta, fabs(elbo_old));
    return (diff < stopping) ? 1 : 0;
}

// ---- np.percentile(method='linear') from order statistics (numpy _function_base_impl._lerp)
XP_HD double np_lerp(double a, double b, double t) {
    const double d = b - a;
    return (t >= 0.5) ? (b - d * (1.0 - t)) : (a + d * t);
}
// gmm.py:39-44 picks the second start location by `loc0 < np.percentile(y, 50)`.  When the median of the reads equals
// the k-mer mean mathematically (2-decimal reads around a 2-decimal model mean: it happens), the reference's choice
// between P75 and P25 hangs on the rounding of numpy's lerp of the UNCENTRED order statistics - so that very expression
// is evaluated here, on the uncentred values ya <= yb (for reads within a factor two of m0, y' + m0 restores y exactly;
// any other median is far from the tie).
XP_HD bool prior_below_median(double ya, double yb, double gamma, double m0) { return m0 < np_lerp(ya, yb, gamma); }
// virtual index (n-1)*q is exact for q in {.25,.5,.75}
XP_HD void percentile_index(int n, double q, int& lo, int& hi, double& gamma) {
    const double h = (double)(n - 1) * q;
    const double fl = floor(h);
    lo = (int)fl;
    hi = (lo + 1 < n) ? lo + 1 : n - 1;
    gamma = h - fl;
}

// ---- prefilter: pooled two-sample t statistic from centred moments (statstest.py:15-16) ----
XP_HD void ttest_from_moments(double na, double s1a, double s2a, double nb, double s1b, double s2b,
                              double& t, double& df) {
    df = na + nb - 2.0;
    if (na < 2.0 || nb < 2.0) {
        t = NAN;
        return;
    }
    const double ma = s1a / na, mb = s1b / nb;
    const double ssa = s2a - s1a * ma, ssb = s2b - s1b * mb;
    const double svar = (ssa + ssb) / df;
    const double denom = sqrt(svar * (1.0 / na + 1.0 / nb));
    t = (ma - mb) / denom;
}

// The same statistic from exact integer sums of a scaled-integer encoding: s1 = sum q (exact),
// nss = n sum q^2 - (sum q)^2 (exact; = n times the side's sum of squared deviations).  t is invariant
// under the common scale and origin of q, so neither enters.
XP_HD void ttest_from_exact_sums(double na, double s1a, double nssa, double nb, double s1b, double nssb,
                                 double& t, double& df) {
    df = na + nb - 2.0;
    if (na < 2.0 || nb < 2.0) {
        t = NAN;
        return;
    }
    const double svar = (nssa / na + nssb / nb) / df;
    const double denom = sqrt(svar * (1.0 / na + 1.0 / nb));
    t = ((s1a * nb - s1b * na) / (na * nb)) / denom;  // s1a nb - s1b na < 2^47: exact
}

// ---- statistics + row filter (io.py:255-367, utils/stats.py) -------------------------------
XP_HD double norm_cdf(double x, double mu, double sigma) {  // stats.py:35-36
    return 0.5 * (1.0 + xerf((x - mu) / (sigma * sqrt(2.0))));
}

// Mean modification rate and banker's-rounded mean coverage of the present groups selected by
// `want(g)`; returns the number of groups used.
// The reference rounds the mean of coverage_g = sum_n sum_k r_nk, a float sum whose exact value
// is the read count n_g (r_n0 + r_n1 = 1).  With an even number of groups the mean can be
// exactly x.5, where the reference's rounding direction is decided by sub-ulp summation noise
// (e.g. 94.99999999999994 vs 95.0).  Here n uses the exact counts: deterministic, and equal to
// the reference whenever its sums carry no noise.
template <class Want>
XP_HD int cond_summary(int G, const int32_t* glen, const double* w_mod, const double* cov, Want want,
                       double& p, double& n) {
    double sw = 0.0, sc = 0.0;
    int cnt = 0;
    for (int g = 0; g < G; ++g)
        if (glen[g] > 0 && want(g)) {
            sw += w_mod[g];
            sc += (double)glen[g];
            ++cnt;
        }
    p = sw / (double)cnt;
    n = rint(sc / (double)cnt);  // .round() = half-to-even (stats.py:11-12)
    return cnt;
}

XP_HD void z_test(double p1, double n1, double p2, double n2, double* out3) {  // stats.py:7-15
    const double se = sqrt(p1 * (1.0 - p1) / n1 + p2 * (1.0 - p2) / n2);
    const double z = (p1 - p2) / se;
    out3[0] = p1 - p2;
    out3[1] = xerfc(fabs(z) * kInvSqrt2);  // 2*norm.sf(|z|)
    out3[2] = z;
}

// w_mod / cov are caller scratch of G doubles each.  Returns status bits KEEP_ROW|MOD_IS_1|HIGHER.
XP_HD uint32_t compute_stats(int G, int C, const int32_t* grp_cond, const int32_t* glen, double conc0,
                             double km, double ks, const double* fit, double* out, double* w_mod, double* cov) {
    const double mu0 = fit[FIT_MU], mu1 = fit[FIT_MU + 1];
    const double s20 = 1.0 / (fit[FIT_ALPHA] / fit[FIT_BETA]);  // io.py:258
    const double s21 = 1.0 / (fit[FIT_ALPHA + 1] / fit[FIT_BETA + 1]);
    // cluster assignment (io.py:274-288, :372-374)
    const double x0 = km - fabs(km - mu0), x1 = km - fabs(km - mu1);
    const double conf0 = xerfc(-((x0 - km) / ks) * kInvSqrt2);  // 2*Phi(z) = erfc(-z/sqrt2)
    const double conf1 = xerfc(-((x1 - km) / ks) * kInvSqrt2);
    const int mod = (conf0 > conf1) ? 1 : 0;
    const int unmod = 1 - mod;
    uint32_t flags = 0;
    if (mod) flags |= STATUS_MOD_IS_1;
    if ((((mu0 < mu1) ? 1 : 0) ^ mod) == 0) flags |= STATUS_HIGHER;
    out[ST_MU] = unmod ? mu1 : mu0;
    out[ST_MU + 1] = mod ? mu1 : mu0;
    out[ST_S2] = unmod ? s21 : s20;
    out[ST_S2 + 1] = mod ? s21 : s20;
    out[ST_CONF] = unmod ? conf1 : conf0;
    out[ST_CONF + 1] = mod ? conf1 : conf0;
    // mixing weights and coverage (gmm.py:287, io.py:266)
    for (int g = 0; g < G; ++g) {
        const double n0 = fit[FIT_N + 2 * g], n1 = fit[FIT_N + 2 * g + 1];
        const double c0 = conc0 + n0, c1 = conc0 + n1;
        w_mod[g] = (mod ? c1 : c0) / (c0 + c1);
        cov[g] = n0 + n1;
        out[ST_RATE + g] = glen[g] > 0 ? w_mod[g] : NAN;
        out[ST_RATE + G + g] = glen[g] > 0 ? cov[g] : NAN;
    }
    // overlap of the two fitted Gaussians (stats.py:18-66)
    double pov = NAN, cx1 = NAN, cy1 = NAN, cx2 = NAN, cy2 = NAN;
    {
        const double sg0 = sqrt(s20), sg1 = sqrt(s21);
        if (sg0 != sg1 && mu0 != mu1) {
            double Xm = mu0, Xs = sg0, Ym = mu1, Ys = sg1;
            if (Ys < Xs || (Ys == Xs && Ym < Xm)) {
                Xm = mu1; Xs = sg1; Ym = mu0; Ys = sg0;
            }
            const double xv = Xs * Xs, yv = Ys * Ys;  // python float ** 2.0
            const double dv = yv - xv;
            const double dm = fabs(Ym - Xm);
            if (dv != 0.0) {
                const double a = Xm * yv - Ym * xv;
                const double b = Xs * Ys * sqrt(dm * dm + dv * xlog(yv / xv));
                const double i1 = (a + b) / dv, i2 = (a - b) / dv;
                cx1 = norm_cdf(i1, Xm, Xs);
                cy1 = norm_cdf(i1, Ym, Ys);
                cx2 = norm_cdf(i2, Xm, Xs);
                cy2 = norm_cdf(i2, Ym, Ys);
                pov = 1.0 - (fabs(cy1 - cx1) + fabs(cy2 - cx2));
            }
        }
    }
    out[ST_POV] = pov;
    out[ST_CDF] = cx1;
    out[ST_CDF + 1] = cy1;
    out[ST_CDF + 2] = cx2;
    out[ST_CDF + 3] = cy2;
    const double th = 0.1;
    const bool not_inside = ((cy1 < th) && (cx1 < th)) || ((cy2 < th) && (cx2 < th)) ||
                            (((1.0 - cy1) < th) && ((1.0 - cx1) < th)) || (((1.0 - cy2) < th) && ((1.0 - cx2) < th));
    if (pov <= 0.5 && not_inside) flags |= STATUS_KEEP_ROW;
    // pairwise z-tests over sorted condition pairs (io.py:293-309)
    double* o = out + ST_RATE + 2 * G;
    for (int a = 0; a < C; ++a)
        for (int b = a + 1; b < C; ++b) {
            double p1, n1, p2, n2;
            const int k1 = cond_summary(G, glen, w_mod, cov, [&](int g) { return grp_cond[g] == a; }, p1, n1);
            const int k2 = cond_summary(G, glen, w_mod, cov, [&](int g) { return grp_cond[g] == b; }, p2, n2);
            if (k1 > 0 && k2 > 0) z_test(p1, n1, p2, n2, o);
            else o[0] = o[1] = o[2] = NAN;
            o += 3;
        }
    // one-vs-all (io.py:311-330): "all" = every other group present at this position
    if (C > 2)
        for (int a = 0; a < C; ++a) {
            double p1, n1, p2, n2;
            const int k1 = cond_summary(G, glen, w_mod, cov, [&](int g) { return grp_cond[g] == a; }, p1, n1);
            const int k2 = cond_summary(G, glen, w_mod, cov, [&](int g) { return grp_cond[g] != a; }, p2, n2);
            if (k1 > 0) {
                if (k2 > 0) z_test(p1, n1, p2, n2, o);
                else { o[0] = NAN; o[1] = NAN; o[2] = NAN; }
            } else o[0] = o[1] = o[2] = NAN;
            o += 3;
        }
    return flags;
}

}  // namespace xpc
