// xp_emul.cpp — host emulator of the kernels' per-position arithmetic (TEST INFRASTRUCTURE).
//
// Links the same xp_math.cuh / xp_core.cuh the CUDA kernels use, with a serial driver in place
// of the warp orchestration, so the CPU test-suite (no GPU in the build container) can check
// the arithmetic restatement — centred moments, ELBO from sufficient statistics, special
// functions, stopping rule, statistics — against the oracle before anything runs on a B200.
// Never loaded by the product (xpore_b200/ has no reference to it).
//
//   g++ -O2 -std=c++17 -ffp-contract=off -shared -fPIC -x c++ xp_emul.cpp -o libxp_emul.so
#include <algorithm>
#include <vector>

#include "../../xpore_b200/csrc/xp_core.cuh"

using namespace xpc;

extern "C" {

int emul_fit_width(int G) { return fit_width(G); }
int emul_stats_width(int G, int C) { return stats_width(G, C); }

void emul_special(int which, const double* x, const double* aux, double* out, long long n) {
    for (long long i = 0; i < n; ++i) {
        double ps, lg, inv, rmin, h;
        bool pos;
        switch (which) {
            case 0: psi_lgamma(x[i], &ps, &lg); out[i] = ps; break;
            case 1: psi_lgamma(x[i], &ps, &lg); out[i] = lg; break;
            case 2: out[i] = xlog(x[i]); break;
            case 3: out[i] = xexp(x[i]); break;
            case 4: out[i] = student_t_two_sided(x[i], aux[i]); break;
            case 5: estep_read(HostTab(), x[i], inv, rmin, h, pos); out[i] = pos ? inv : rmin; break;  // r_n1
            case 6: estep_read(HostTab(), x[i], inv, rmin, h, pos); out[i] = -h; break;  // sum_k r ln r
            default: out[i] = NAN;
        }
    }
}

// t-test of one position; returns p (NaN if not exactly two conditions present)
double emul_prefilter(int n, const double* y, int G, const int* glen, const int* grp_cond, double km) {
    unsigned cmask = 0;
    for (int g = 0; g < G; ++g)
        if (glen[g] > 0) cmask |= 1u << grp_cond[g];
    if (__builtin_popcount(cmask) != 2) return NAN;
    const int ca = __builtin_ctz(cmask);
    double na = 0, s1a = 0, s2a = 0, nb = 0, s1b = 0, s2b = 0;
    int i = 0;
    for (int g = 0; g < G; ++g)
        for (int j = 0; j < glen[g]; ++j, ++i) {
            const double v = y[i] - km;
            if (grp_cond[g] == ca) { na += 1; s1a += v; s2a = fma(v, v, s2a); }
            else { nb += 1; s1b += v; s2b = fma(v, v, s2b); }
        }
    (void)n;
    double t, df;
    ttest_from_moments(na, s1a, s2a, nb, s1b, s2b, t, df);
    return student_t_two_sided(t, df);
}

// VB-EM of one position, same phases as xp_fit_kernel.  elbos (optional) gets the ELBO after
// every sweep (index it-1), capacity max_iters.
void emul_fit(int n, const double* y, int G, const int* glen, double km, double ks, int max_iters, double stopping,
              double conc0, double* fit, int* n_iter, unsigned* status, double* elbos) {
    std::vector<double> yv(n);
    for (int i = 0; i < n; ++i) yv[i] = y[i] - km;
    std::vector<int> goff(G + 1, 0);
    for (int g = 0; g < G; ++g) goff[g + 1] = goff[g] + glen[g];
    const double tau = 1.0 / (ks * ks);
    const double a0 = tau, b0 = 0.5 * (1.0 / tau);
    const double lg_prior_w = xlgamma(2.0 * conc0) - 2.0 * xlgamma(conc0);
    std::vector<double> N0g(G, 0.0), dpsi(G), Asm(G);   // N_g0; psi(c_g1) - psi(c_g0)
    double Kconst = 0.0;
    for (int g = 0; g < G; ++g) {
        double ps, lg;
        psi_lgamma(2.0 * conc0 + (double)glen[g], &ps, &lg);
        if (glen[g] > 0) Kconst += lg_prior_w - lg;
    }
    const double prior_gamma = a0 * xlog(b0) - xlgamma(a0);
    // totals: component 1 follows from them because r_n1 = 1 - r_n0 (gmm.py:242)
    double T1 = 0.0, T2 = 0.0;
    for (int i = 0; i < n; ++i) { T1 += yv[i]; T2 = fma(yv[i], yv[i], T2); }
    // initial locations
    std::vector<double> sorted(yv);
    std::sort(sorted.begin(), sorted.end());
    auto pct = [&](double q) {
        int lo, hi;
        double gamma;
        percentile_index(n, q, lo, hi, gamma);
        return np_lerp(sorted[lo], sorted[hi], gamma);
    };
    bool up;
    {
        int lo, hi;
        double gamma;
        percentile_index(n, 0.5, lo, hi, gamma);
        up = prior_below_median(sorted[lo] + km, sorted[hi] + km, gamma, km);  // as the kernel's generic path
    }
    const double loc1 = pct(up ? 0.75 : 0.25);

    Comp c[2];
    c[0] = mstep(0, 0, 0, a0, b0);
    c[1] = c[0];
    c[1].mp = loc1;
    double Nk[2] = {0, 0}, S1[2] = {0, 0}, S2[2] = {0, 0}, Hent = 0;
    double elbo_old = -INFINITY, elbo = -INFINITY;
    int it = 0;
    unsigned st = STATUS_SELECTED;
    double tb[2], lt[2];
    bool first = true;
    for (;;) {
        double e = 0.0;
        for (int g = 0; g < G; ++g) {
            double ps[2], lg[2];
            for (int k = 0; k < 2; ++k) {
                const double N = first ? 0.0 : (k ? (double)glen[g] - N0g[g] : N0g[g]);
                psi_lgamma_parts(HostTab(), conc0 + N, &ps[k], &lg[k]);
                if (glen[g] > 0) e += lg[k];
            }
            dpsi[g] = ps[1] - ps[0];
        }
        for (int k = 0; k < 2; ++k) {
            double ps, lg;
            psi_lgamma_parts(HostTab(), c[k].alpha, &ps, &lg);
            const double lnb = xlog_tab(HostTab(), c[k].beta), lnl = xlog_tab(HostTab(), c[k].lam);
            if (first) e += elbo_component(c[k], Nk[k], S1[k], S2[k], a0, b0, prior_gamma, ps, lg, lnb, lnl, tb[k], lt[k]);
            else e += elbo_component_closed(c[k], ps, lg, lnb, lnl, tb[k], lt[k]);
        }
        elbo = e + (first ? Kconst : Kconst + elbo_closed_const(prior_gamma, n)) + Hent;
        if (it >= 1) {
            if (elbos) elbos[it - 1] = elbo;
            const int rule = stop_rule(elbo, elbo_old, stopping);
            if (rule == 2) { st |= STATUS_ELBO_DECREASED; break; }
            elbo_old = elbo;
            if (rule == 1) { st |= STATUS_CONVERGED; break; }
        }
        if (it >= max_iters - 1) {
            if (it >= 1) st |= STATUS_HIT_MAX;
            break;
        }
        ++it;
        first = false;
        double base, B, C;
        estep_coefficients(c[0], c[1], tb[0], tb[1], lt[0], lt[1], base, B, C);
        for (int g = 0; g < G; ++g) Asm[g] = base + dpsi[g];
        double s10 = 0.0, s20 = 0.0, n0tot = 0.0;
        Hent = 0.0;
        for (int g = 0; g < G; ++g) {
            double n0 = 0;
            for (int i = goff[g]; i < goff[g + 1]; ++i) {
                const double yy = yv[i];
                const double d = fma(yy, fma(C, yy, B), Asm[g]);
                double inv, rmin, h;
                bool pos;
                estep_read(HostTab(), d, inv, rmin, h, pos);
                const double r0 = pos ? rmin : inv;
                const double w = r0 * yy;
                n0 += r0;
                s10 += w;
                s20 = fma(w, yy, s20);
                Hent += h;
            }
            N0g[g] = n0;
            n0tot += n0;
        }
        Nk[0] = n0tot; Nk[1] = (double)n - n0tot;
        S1[0] = s10; S1[1] = T1 - s10;
        S2[0] = s20; S2[1] = T2 - s20;
        c[0] = mstep(Nk[0], S1[0], S2[0], a0, b0);
        c[1] = mstep(Nk[1], S1[1], S2[1], a0, b0);
    }
    fit[FIT_MU] = km + c[0].mp; fit[FIT_MU + 1] = km + c[1].mp;
    fit[FIT_LAM] = c[0].lam; fit[FIT_LAM + 1] = c[1].lam;
    fit[FIT_ALPHA] = c[0].alpha; fit[FIT_ALPHA + 1] = c[1].alpha;
    fit[FIT_BETA] = c[0].beta; fit[FIT_BETA + 1] = c[1].beta;
    fit[FIT_ELBO] = elbo;
    for (int g = 0; g < G; ++g) {
        fit[FIT_N + 2 * g] = (it == 0) ? 0.0 : N0g[g];
        fit[FIT_N + 2 * g + 1] = (it == 0) ? 0.0 : (double)glen[g] - N0g[g];
    }
    *n_iter = it;
    *status = st;
}

unsigned emul_stats(int G, int C, const int* grp_cond, const int* glen, double conc0, double km, double ks,
                    const double* fit, double* out) {
    std::vector<double> w(G), cov(G);
    return compute_stats(G, C, grp_cond, glen, conc0, km, ks, fit, out, w.data(), cov.data());
}

}  // extern "C"
