// xp_math.cuh — FP64 special functions for the diffmod kernels (host+device).
//
// Everything here is written against IEEE-754 double with explicit fma(), so the same source
// gives the same bits under nvcc (device) and g++ (host emulator used by the CPU tests), except
// xrcp() which uses the hardware reciprocal seed on the device.
//
// Replaces the third-party calls the reference makes from its hot loop
// (/root/reference/xpore/diffmod/gmm.py): np.exp :241, scipy.special.logsumexp :243,
// scipy.special.digamma :283-284,350, scipy.special.gammaln :297,334,343, np.log :350; and
// from the statistics (statstest.py:15-17 scipy.stats.ttest_ind / t.sf; utils/stats.py:15
// norm.sf; io.py:373 norm.cdf; stats.py:36 math.erf).
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define XP_HD __host__ __device__ __forceinline__
#else
#define XP_HD inline
#endif

// Polynomial coefficients live in __constant__ memory on the device so that DFMA takes them as
// constant-bank operands (64-bit literals would be rebuilt with two UMOVs per use inside the
// hot loops); the host build reads the same values from static arrays.
#if defined(__CUDACC__)
#define XP_TABLE(name, n, ...)                        \
    static const double h_##name[n] = {__VA_ARGS__};  \
    __constant__ double d_##name[n] = {__VA_ARGS__};
#else
#define XP_TABLE(name, n, ...) static const double h_##name[n] = {__VA_ARGS__};
#endif
#if defined(__CUDA_ARCH__)
#define XPT(name) d_##name
#else
#define XPT(name) h_##name
#endif

namespace xpm {

// log1p kernel R(z), fdlibm/msun e_log.c Lg7..Lg1
XP_TABLE(kLg, 7, 1.479819860511658591e-01, 1.531383769920937332e-01, 1.818357216161805012e-01,
         2.222219843214978396e-01, 2.857142874366239149e-01, 3.999999999940941908e-01, 6.666666666666735130e-01)
// 1/13! .. 1/2!
XP_TABLE(kExp, 12, 1.6059043836821613e-10, 2.08767569878681e-09, 2.505210838544172e-08, 2.755731922398589e-07,
         2.7557319223985893e-06, 2.48015873015873e-05, 1.984126984126984e-04, 1.388888888888889e-03,
         8.333333333333333e-03, 4.1666666666666664e-02, 1.6666666666666666e-01, 0.5)
// B_2k/(2k): 1/12 -691/32760 1/132 -1/240 1/252 -1/120 1/12   (highest order first)
XP_TABLE(kPsi, 7, 8.33333333333333333333e-2, -2.10927960927960927961e-2, 7.57575757575757575758e-3,
         -4.16666666666666666667e-3, 3.96825396825396825397e-3, -8.33333333333333333333e-3,
         8.33333333333333333333e-2)
// B_2k/(2k(2k-1)): 1/156 -691/360360 1/1188 -1/1680 1/1260 -1/360 1/12
XP_TABLE(kLgam, 7, 6.41025641025641025641e-3, -1.91752691752691752692e-3, 8.41750841750841750842e-4,
         -5.95238095238095238095e-4, 7.93650793650793650794e-4, -2.77777777777777777778e-3,
         8.33333333333333333333e-2)
// ln2 hi/lo, log2(e), ln2, sqrt2, round-to-int magic
XP_TABLE(kMisc, 6, 6.93147180369123816490e-01, 1.90821492927058770002e-10, 1.44269504088896338700e+00,
         6.931471805599453094e-01, 1.41421356237309504880e+00, 6755399441055744.0)

#include "xp_tables.inc"  // kExp2[256], kInvC[129], kLogC[129], kE256[3] (tools/make_tables.py)
// full-precision series coefficients of the E-step: 1/3! (exp) and 1/3 (log1p)
XP_TABLE(kE3, 2, 1.6666666666666666e-01, 3.3333333333333331e-01)

// The highest-order coefficient of each E-step series rounded to a double whose low word is zero: DFMA
// takes it as a 32-bit immediate.  Its rounding (rel. 2^-21) is far below the weight of the term it
// multiplies (r^4/24 <= 1.4e-13, rr^5/5 <= 1.9e-13 relative to the results).
constexpr double kExpC4 = 0x1.5555500000000p-5;   // ~ 1/24
constexpr double kLogC5 = 0x1.9999900000000p-3;   // ~ 1/5

XP_HD uint64_t d2u(double x) {
#if defined(__CUDA_ARCH__)
    return (uint64_t)__double_as_longlong(x);
#else
    uint64_t u;
    memcpy(&u, &x, 8);
    return u;
#endif
}
XP_HD double u2d(uint64_t u) {
#if defined(__CUDA_ARCH__)
    return __longlong_as_double((long long)u);
#else
    double x;
    memcpy(&x, &u, 8);
    return x;
#endif
}

// high / low 32-bit words of a double
XP_HD int hi32(double x) {
#if defined(__CUDA_ARCH__)
    return __double2hiint(x);
#else
    return (int)(uint32_t)(d2u(x) >> 32);
#endif
}
XP_HD int lo32(double x) {
#if defined(__CUDA_ARCH__)
    return __double2loint(x);
#else
    return (int)(uint32_t)d2u(x);
#endif
}
XP_HD double mk64(int hi, int lo) {
#if defined(__CUDA_ARCH__)
    return __hiloint2double(hi, lo);
#else
    return u2d(((uint64_t)(uint32_t)hi << 32) | (uint64_t)(uint32_t)lo);
#endif
}

constexpr double kLn2Hi = 6.93147180369123816490e-01;
constexpr double kLn2Lo = 1.90821492927058770002e-10;
constexpr double kLn2 = 6.931471805599453094e-01;
constexpr double kLog2e = 1.44269504088896338700e+00;
constexpr double kSqrt2 = 1.41421356237309504880e+00;
constexpr double kLn2Pi = 1.83787706640934548356e+00;
constexpr double kHalfLn2Pi = 0.91893853320467274178e+00;
constexpr double kInvSqrt2 = 0.70710678118654752440e+00;

// 1/x for normal positive x.  Device: 20-bit hardware seed + two Newton steps (<= 1 ulp).
XP_HD double xrcp(double x) {
#if defined(__CUDA_ARCH__)
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double e = fma(-x, y, 1.0);
    y = fma(y, e, y);
    e = fma(-x, y, 1.0);
    y = fma(y, e, y);
    return y;
#else
    return 1.0 / x;
#endif
}

// 1/x for normal positive x (the E-step calls it with x in [1, 2]).  Device: hardware seed (rel. error <= 2^-20) + one cubic step
// y (1 + e + e^2), e = 1 - x y  (error e^3, <= 1 ulp).
XP_HD double xrcp3(double x) {
#if defined(__CUDA_ARCH__)
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double e = fma(-x, y, 1.0);
    const double g = fma(e, e, e);
    return fma(y, g, y);
#else
    return 1.0 / x;
#endif
}

// a/b: exact IEEE on the host, a * xrcp(b) (<= 2 ulp) on the device.
XP_HD double xdiv(double a, double b) {
#if defined(__CUDA_ARCH__)
    return a * xrcp(b);
#else
    return a / b;
#endif
}
// the same with the one-step cubic reciprocal (the hardware seed is good to 2^-20 for every normal
// operand, so the cubic step leaves <= 2^-60): one DFMA less, used once per sweep and call site.
XP_HD double xdiv3(double a, double b) {
#if defined(__CUDA_ARCH__)
    return a * xrcp3(b);
#else
    return a / b;
#endif
}

// log(1+f) kernel for f in [sqrt(1/2)-1, sqrt(2)-1], given s = f/(2+f).
// Classic atanh form: log(1+f) = 2 atanh(s) = f - f^2/2 + s*(f^2/2 + R(s^2)); R is the degree-7
// minimax polynomial of the freely published fdlibm/msun e_log.c (|err| < 2^-58.45).
XP_HD double log1p_core(double f, double s) {
    const double z = s * s;
    double R = XPT(kLg)[0];
#pragma unroll
    for (int i = 1; i < 7; ++i) R = fma(R, z, XPT(kLg)[i]);
    R = R * z;
    const double hfsq = 0.5 * f * f;
    return f - (hfsq - s * (hfsq + R));
}

// Natural log of a positive, normal double.
XP_HD double xlog(double x) {
    uint64_t u = d2u(x);
    int e = (int)(u >> 52) - 1023;
    u = (u & 0x000FFFFFFFFFFFFFull) | 0x3FF0000000000000ull;
    double m = u2d(u);  // [1,2)
    if (m > kSqrt2) {
        m *= 0.5;
        e += 1;
    }
    const double f = m - 1.0;
    const double s = xdiv(f, 2.0 + f);
    const double dk = (double)e;
    const double l = log1p_core(f, s);
    return fma(dk, XPT(kMisc)[0], fma(dk, XPT(kMisc)[1], l));
}

// log(1+u) for u > -0.29 (accurate for tiny u).
XP_HD double xlog1p(double u) {
    if (u < 0.41 && u > -0.29) return log1p_core(u, xdiv(u, 2.0 + u));
    return xlog(1.0 + u);
}

// exp(r) for |r| <= 0.35: Taylor to degree 13 (truncation < 5e-18 relative).
XP_HD double exp_poly(double r) {
    double p = XPT(kExp)[0];
#pragma unroll
    for (int i = 1; i < 12; ++i) p = fma(p, r, XPT(kExp)[i]);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    return p;
}

// exp(x) for x in [-708, 708] (result stays a normal double).
XP_HD double xexp(double x) {
    const double magic = XPT(kMisc)[5];  // 1.5 * 2^52: round-to-nearest-integer by addition
    const double t = fma(x, XPT(kMisc)[2], magic);
    const double kf = t - magic;
    const int64_t k = (int64_t)(int32_t)(uint32_t)d2u(t);
    double r = fma(-kf, XPT(kMisc)[0], x);
    r = fma(-kf, XPT(kMisc)[1], r);
    const double p = exp_poly(r);
    return u2d(d2u(p) + ((uint64_t)k << 52));
}

// psi(x) and lgamma(x) for x > 0 in one pass: shift to z >= 10 by the recurrences
//   psi(x) = psi(x+m) - sum_{j<m} 1/(x+j),   lgamma(x) = lgamma(x+m) - log prod_{j<m}(x+j)
// (the sum is kept as one fraction num/den), then the Stirling series with Bernoulli terms
// through B14.  |err| ~ 1e-15 relative.  Split in three steps so the kernels can run the two
// logarithms of all lanes' arguments through one uniform code path.
XP_HD void shift_to_10(double x, double& z, double& num, double& den) {
    num = 0.0;
    den = 1.0;
    z = x;
    while (z < 10.0) {
        num = fma(num, z, den);
        den = den * z;
        z += 1.0;
    }
}
// The same shift for the sweeps' scalar phase, without a data-dependent loop: every x < 10 moves up by
// exactly ten, z = x + 10, and the two products are the fixed polynomials
//   den = P(x) = x (x+1) ... (x+9) = sum_k [10 k] x^k      (unsigned Stirling numbers of the first kind)
//   num = P'(x)                                            (so num/den = sum_j 1/(x+j))
// evaluated by two independent Horner chains (all coefficients are small integers, x > 0: no cancellation).
XP_HD void shift10(double x, double& z, double& num, double& den) {
    double p = x + 45.0;
    double d = fma(10.0, x, 405.0);
    p = fma(p, x, 870.0);
    d = fma(d, x, 6960.0);
    p = fma(p, x, 9450.0);
    d = fma(d, x, 66150.0);
    p = fma(p, x, 63273.0);
    d = fma(d, x, 379638.0);
    p = fma(p, x, 269325.0);
    d = fma(d, x, 1346625.0);
    p = fma(p, x, 723680.0);
    d = fma(d, x, 2894720.0);
    p = fma(p, x, 1172700.0);
    d = fma(d, x, 3518100.0);
    p = fma(p, x, 1026576.0);
    d = fma(d, x, 2053152.0);
    p = fma(p, x, 362880.0);
    d = fma(d, x, 362880.0);
    p = p * x;
    const bool sh = x < 10.0;
    z = sh ? x + 10.0 : x;
    num = sh ? d : 0.0;
    den = sh ? p : 1.0;
}

// ln x of a positive normal double through the E-step's {1/c, ln c} table: x = 2^e u, u in [1, 2),
// ln u = ln c_i + ln1p(u / c_i - 1) with the degree-5 series (absolute error < 1e-15).
template <class Tab>
XP_HD double xlog_tab(const Tab& tab, double x) {
    const int hx = hi32(x);
    const int i = (hx >> 13) & 127;
    const double u = mk64((hx & 0x000fffff) | 0x3ff00000, lo32(x));
    const double e = (double)((hx >> 20) - 1023);
    double ic, lc;
    tab.logc(i, ic, lc);
    const double rr = fma(u, ic, -1.0);
    double q = fma(kLogC5, rr, -0.25);
    q = fma(q, rr, XPT(kE3)[1]);
    q = fma(q, rr, -0.5);
    q = fma(q, rr, 1.0);
    const double lu = fma(q, rr, lc);
    return fma(e, XPT(kMisc)[0], fma(e, XPT(kMisc)[1], lu));
}

// psi and lgamma + ln den from the shifted argument: lz = ln z, iz = 1/z, sden = num/den.  The caller
// subtracts ln den itself (the scalar phase takes one logarithm of the product of all its den).
XP_HD void psi_lgamma_shifted(double z, double lz, double iz, double sden, double* psi, double* lgam_plus_lden) {
    const double w = iz * iz;
    double a = XPT(kPsi)[0];
    double b = XPT(kLgam)[0];
#pragma unroll
    for (int i = 1; i < 7; ++i) {
        a = fma(a, w, XPT(kPsi)[i]);
        b = fma(b, w, XPT(kLgam)[i]);
    }
    *psi = lz - 0.5 * iz - a * w - sden;
    *lgam_plus_lden = fma(z - 0.5, lz, -z) + kHalfLn2Pi + b * iz;
}

// lz = ln z, lden = ln den, iz = 1/z, sden = num/den
XP_HD void psi_lgamma_finish(double z, double lz, double lden, double iz, double sden, double* psi, double* lgam) {
    const double w = iz * iz;
    double a = XPT(kPsi)[0];
    double b = XPT(kLgam)[0];
#pragma unroll
    for (int i = 1; i < 7; ++i) {
        a = fma(a, w, XPT(kPsi)[i]);
        b = fma(b, w, XPT(kLgam)[i]);
    }
    *psi = lz - 0.5 * iz - a * w - sden;
    *lgam = fma(z - 0.5, lz, -z) + kHalfLn2Pi + b * iz - lden;
}
XP_HD void psi_lgamma(double x, double* psi, double* lgam) {
    double z, num, den;
    shift_to_10(x, z, num, den);
    const double q = xdiv(1.0, z * den);
    psi_lgamma_finish(z, xlog(z), (den != 1.0) ? xlog(den) : 0.0, q * den, num * (q * z), psi, lgam);
}

XP_HD double xlgamma(double x) {
    double p, l;
    psi_lgamma(x, &p, &l);
    return l;
}

XP_HD double xerfc(double x) { return erfc(x); }
XP_HD double xerf(double x) { return erf(x); }

// Continued fraction of the regularised incomplete beta function I_x(a,b) (the one of Numerical Recipes' betacf):
//   h = 1 / (1 + d_1 / (1 + d_2 / (1 + ...))),  d_1 = -(a+b) x / (a+1),
//   d_2m = m (b-m) x / ((a+2m-1)(a+2m)),  d_2m+1 = -(a+m)(a+b+m) x / ((a+2m)(a+2m+1)).
// Evaluated by the forward recurrence on numerators and denominators, A_n = A_(n-1) + d_n A_(n-2), with every step
// multiplied through by the denominator D of d_n = N / D - (A_(n-2), A_(n-1)) -> (D A_(n-1), D A_(n-1) + N A_(n-2)) -
// so that no division is executed per step (the modified Lentz form needs five per iteration; this loop is a
// third of the prefilter kernel's instructions).  N and D are scaled by the power of two 2^(-2 e), e the exponent of
// a + 2m (exact), which keeps D in [1/2, 4); all four terms are brought back to unit scale every 64 iterations.
// Stops when two successive convergents agree to 1e-14 (the recurrence carries a few ulp of noise, and the
// coefficients of a ~ 1e3 are themselves rounded at that level).
XP_HD double betacf(double a, double b, double x) {
    const double qab = a + b, qap = a + 1.0;
    double A0 = 1.0, A1 = 1.0, B0 = 1.0, B1 = 1.0 - xdiv3(qab * x, qap);
    int m = 1;
    for (int blk = 0; blk < 8; ++blk) {
        bool done = false;
        for (int i = 0; i < 64; ++i, ++m) {
            const double dm = (double)m;
            const double u = a + 2.0 * dm;
            // 2^(-2 e): u >= 2.5, so the biased exponent 3069 - 2 e stays far inside the normal range
            const int e = (int)((d2u(u) >> 52) & 0x7ff);
            const double s = u2d((uint64_t)(3069 - 2 * e) << 52);
            const double us = u * s, xs = x * s;
            double N = (dm * (b - dm)) * xs, D = (u - 1.0) * us;
            double T = D * A1;
            double A2 = fma(N, A0, T);
            A0 = T;
            T = D * B1;
            double B2 = fma(N, B0, T);
            B0 = T;
            N = -((a + dm) * (qab + dm)) * xs;
            D = (u + 1.0) * us;
            T = D * A2;
            A1 = fma(N, A0, T);
            A0 = T;
            T = D * B2;
            B1 = fma(N, B0, T);
            B0 = T;
            const double t2 = A0 * B1;
            if (fabs(fma(A1, B0, -t2)) < 1e-14 * fabs(t2)) {
                done = true;
                break;
            }
        }
        if (done) break;
        const double sc = u2d((uint64_t)(2046 - (int)((d2u(B1) >> 52) & 0x7ff)) << 52);  // 2^-(exponent of B1)
        A0 *= sc;
        A1 *= sc;
        B0 *= sc;
        B1 *= sc;
    }
    return xdiv3(A1, B1);
}

// Stirling correction  sum_k B_2k / (2k(2k-1) z^(2k-1))  for z >= 10.
XP_HD double stirling_corr(double z) {
    const double iz = 1.0 / z, w = iz * iz;
    double b = XPT(kLgam)[0];
#pragma unroll
    for (int i = 1; i < 7; ++i) b = fma(b, w, XPT(kLgam)[i]);
    return b * iz;
}

// lgamma(a+b) - lgamma(a) without the cancellation of two large lgammas (a >= 10):
//   (a-1/2) log1p(b/a) + b ln(a+b) - b + corr(a+b) - corr(a)
XP_HD double lgamma_diff(double a, double b) {
    if (a < 10.0) return xlgamma(a + b) - xlgamma(a);
    return (a - 0.5) * xlog1p(b / a) + b * xlog(a + b) - b + (stirling_corr(a + b) - stirling_corr(a));
}

// comp: evaluate 1 - I_xc(b, a) instead of I_x(a, b).  One call of the continued fraction with the chosen arguments:
// the lanes of a warp that take different sides run the same loop.
XP_HD double betainc(double a, double b, double x, double xc, double lnx, double lnxc, bool comp) {
    if (x <= 0.0) return 0.0;
    if (xc <= 0.0) return 1.0;
    const double lbeta = lgamma_diff(a, b) - xlgamma(b);
    const double lfront = lbeta + a * lnx + b * lnxc;
    const double front = (lfront < -700.0) ? 0.0 : xexp(lfront);
    const double ca = comp ? b : a, cb = comp ? a : b, cx = comp ? xc : x;
    const double r = front * betacf(ca, cb, cx) / ca;
    return comp ? 1.0 - r : r;
}

// Two-sided Student-t p-value 2*sf(|t|, df) = I_{df/(df+t^2)}(df/2, 1/2)   (statstest.py:17).
// The continued fraction in x converges within O(sqrt(a)) iterations below x = (a+1)/(a+b+2) and ever more slowly
// towards it (df = 358: 50 iterations at |t| = 1.73, 27 at 2.5), the one of the complement in a handful (8 and 11
// there); the complement costs digits through 1 - q only where p is small.  So the complement is taken up to
// t^2 = 8 for df >= 20 (p >= 0.005: at most two and a half digits, measured 3e-12 against scipy), which keeps the
// longest lane of a warp at ~25 iterations instead of ~50.
XP_HD double student_t_two_sided(double t, double df) {
    if (t != t || df != df) return NAN;
    if (!(df > 0.0)) return NAN;
    const double t2 = t * t;
    if (t2 == INFINITY) return 0.0;
    if (t2 < 1e-280) return 1.0;
    const double x = df / (df + t2);
    const double xc = t2 / (df + t2);  // 1-x without cancellation
    const double lnx = -xlog1p(t2 / df);
    const double lnxc = xlog(xc);
    const double a = 0.5 * df;
    const bool comp = !(x < (a + 1.0) / (a + 2.5)) || (t2 < 8.0 && df >= 20.0);
    return betainc(a, 0.5, x, xc, lnx, lnxc, comp);
}

}  // namespace xpm
