// acro_fastquad.cuh -- the production evaluation of the three Planck-weighted kernel integrals (a9, a10, a11).
//
// Same integrals, same limits as the reference (cascade.py:406-435, 507-540, 562-598) and as the plain Gauss-Legendre
// version in acro_kernels.cu (kernel_mode = ACRO_KERNELS_PLAIN_GL, kept as the in-library cross-check); what changes
// is how they are evaluated:
//
//  * STEEP regime  x_a/T > e  and  (x_b - x_a)/T >= 32  (the lower limit sits on the Wien tail of the thermal photons):
//    with s = (x - x_a)/T the integral is  T e^{-x_a/T}/pi^2 * int_0^inf e^{-s} g(s) ds,  g smooth on the scale x_a/T > e,
//    so a Gauss-Laguerre rule of 24 ... 4 nodes (by x_a/T) is exact to < 4e-10 -- and needs NO exponential per node
//    (e^{-s_k} are constants).  Nodes beyond x_b (weight < e^-32) are masked.  Shorter spans: GL-32 in s.
//  * log and exp: a 128-entry table-driven log (flog_tab) and Estrin-evaluated polynomials (ACRO_FQ_LOGTAB, ACRO_FQ_ESTRIN;
//    0 selects the atanh-series log / Horner forms they replaced: same results to 2e-16, 5 % slower kernels).
//  * otherwise the Q-point Gauss-Legendre rule in log(eps_bar) as before, but with the integrands reduced
//    algebraically (q = x_a/x for the photon kernel so that log q = log x_a - y; Gamma*q = A/B with A + B = E' for
//    the electron kernel; one shared reciprocal for the pair-creation function), division-free reciprocals and a
//    branch-free exp.  The range tests of F and G are dropped: inside the integration limits they are identities
//    ("CHECKED to never happen", cascade.py:33-35,63-66); kernel_mode = PLAIN_GL keeps them.
#pragma once
#include "acro_physics.cuh"

#ifndef ACRO_FQ_UNROLL
#define ACRO_FQ_UNROLL 1
#endif
#ifndef ACRO_FQ_ESTRIN
#define ACRO_FQ_ESTRIN 1
#endif
#ifndef ACRO_FQ_LOGTAB
#define ACRO_FQ_LOGTAB 1
#endif
#define ACRO_PRAGMA(x) _Pragma(#x)
#define ACRO_UNROLL(n) ACRO_PRAGMA(unroll n)

namespace acro {

// Gauss-Laguerre rules n = 4, 6, 8, 12, 16, 24 packed back to back (offsets 0, 4, 10, 18, 30, 46): nodes, weights
// (include e^{-s}), e^{-s_k}
constexpr int kLagTotal = 80;     // Gauss-Laguerre rules of order 4, 6, 8, 12, 16, 24, 10 back to back
__constant__ double c_lag_s[kLagTotal];
__constant__ double c_lag_w[kLagTotal];
__constant__ double c_lag_es[kLagTotal];

#ifndef ACRO_FQ_LAG_UNROLL
#define ACRO_FQ_LAG_UNROLL 2
#endif
constexpr double kHalfRange = 8.0;              // log-range [e-folds] up to which Q/2 nodes are used (a9, a11 only)
constexpr double kSteepX = 2.718281828459045;   // x_a/T above which the Laguerre form is used (u_a > 1)
constexpr double kSteepSpan = 32.0;             // and the span (x_b - x_a)/T must exceed this (truncation e^-32)

// 1/x to full double precision: MUFU.RCP64H seed + two Newton steps, no special-case handling (x finite, normal, != 0)
__device__ __forceinline__ double frcp(double x) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
#if ACRO_FQ_ESTRIN
    const double e = fma(-x, y, 1.0);      // seed error ~2^-20: one cubic step y (1 + e + e^2) leaves e^3
    return fma(y, fma(e, e, e), y);
#else
    double e = fma(-x, y, 1.0);
    y = fma(y, e, y);
    e = fma(-x, y, 1.0);
    return fma(y, e, y);
#endif
}

// Polynomial coefficients live in constant memory so that every FMA takes its constant straight from the constant bank
// (as 64-bit immediates they cost two register moves each inside the node loops: ~25 % of the issued instructions).
__constant__ double c_exp[14] = {1.0, 1.0, 0.5, 1.6666666666666666e-01, 4.1666666666666664e-02, 8.3333333333333332e-03,
                                       1.3888888888888889e-03, 1.9841269841269841e-04, 2.4801587301587302e-05,
                                       2.7557319223985888e-06, 2.7557319223985893e-07, 2.5052108385441720e-08,
                                       2.0876756987868100e-09, 1.6059043836821613e-10};
__constant__ double c_exp_rr[3] = {1.4426950408889634, -6.93147180369123816490e-01, -1.90821492927058770002e-10};
// log: log(m) = 2 atanh(f), f = (m-1)/(m+1), m in [sqrt(1/2), sqrt(2)): 2f (1 + f^2/3 + ... + f^18/19), |f| <= 0.1716
__constant__ double c_log[10] = {2.0, 2.0 / 3.0, 2.0 / 5.0, 2.0 / 7.0, 2.0 / 9.0, 2.0 / 11.0, 2.0 / 13.0, 2.0 / 15.0,
                                       2.0 / 17.0, 2.0 / 19.0};
__constant__ double c_ln2[2] = {6.93147180369123816490e-01, 1.90821492927058770002e-10};
__constant__ double c_bose[4] = {1.0 / 12.0, -1.0 / 720.0, 1.0 / 30240.0, -1.0 / 1209600.0};

// exp(x) for -700 < x < 700 (degree-13 Taylor polynomial on [-ln2/2, ln2/2], remainder 4e-18), no range handling
__device__ __forceinline__ double fexp(double x) {
    const double t = fma(x, c_exp_rr[0], 6755399441055744.0);
    const int k = __double2loint(t);
    const double kf = t - 6755399441055744.0;
    double r = fma(kf, c_exp_rr[1], x);
    r = fma(kf, c_exp_rr[2], r);
#if ACRO_FQ_ESTRIN
    const double r2 = r * r, r4 = r2 * r2, r8 = r4 * r4;   // Estrin: dependency depth 4 instead of 13
    const double a0 = fma(c_exp[1], r, c_exp[0]), a1 = fma(c_exp[3], r, c_exp[2]), a2 = fma(c_exp[5], r, c_exp[4]),
                 a3 = fma(c_exp[7], r, c_exp[6]), a4 = fma(c_exp[9], r, c_exp[8]), a5 = fma(c_exp[11], r, c_exp[10]),
                 a6 = fma(c_exp[13], r, c_exp[12]);
    const double b0 = fma(a1, r2, a0), b1 = fma(a3, r2, a2), b2 = fma(a5, r2, a4);
    const double p = fma(fma(a6, r4, b2), r8, fma(b1, r4, b0));
#else
    double p = c_exp[13];
#pragma unroll
    for (int i = 12; i >= 0; --i) p = fma(p, r, c_exp[i]);
#endif
    return __hiloint2double(__double2hiint(p) + (k << 20), __double2loint(p));
}

// log(x) for positive normal x (no special cases): exponent/mantissa split, atanh series, error ~2e-16 relative to |log|+1
__device__ __forceinline__ double flog(double x) {
    int hi = __double2hiint(x);
    const int lo = __double2loint(x);
    int e = (hi >> 20) - 1023;
    hi = (hi & 0x000fffff) | 0x3ff00000;          // mantissa in [1, 2)
    if (hi >= 0x3ff6a09e) { hi -= 0x00100000; e += 1; }   // m >= sqrt(2): halve it
    const double m = __hiloint2double(hi, lo);
    const double f = (m - 1.0) * frcp(m + 1.0);
    const double f2 = f * f;
#if ACRO_FQ_ESTRIN
    const double z2 = f2 * f2, z4 = z2 * z2;    // polynomial in z = f^2: depth 4 instead of 9
    const double a0 = fma(c_log[1], f2, c_log[0]), a1 = fma(c_log[3], f2, c_log[2]), a2 = fma(c_log[5], f2, c_log[4]),
                 a3 = fma(c_log[7], f2, c_log[6]), a4 = fma(c_log[9], f2, c_log[8]);
    const double p = fma(fma(a4, z4, fma(a3, z2, a2)), z4, fma(a1, z2, a0));
#else
    double p = c_log[9];
#pragma unroll
    for (int i = 8; i >= 0; --i) p = fma(p, f2, c_log[i]);
#endif
    const double ef = (double)e;
    return fma(ef, c_ln2[0], fma(p, f, ef * c_ln2[1]));
}

// Table-driven log for the node loops: x = 2^e m, m in [1,2) falls into one of 128 intervals with centre c_i; with
// (u_i, v_i) = (1/c_i, -log u_i - [i >= 53] ln 2) -- the second term moves m >= sqrt(2) to the exponent above so that nothing
// cancels around x = 1 -- log x = (e + [i >= 53]) ln 2 + v_i + log1p(m u_i - 1), |m u_i - 1| < 4e-3, series to r^7 (< 2e-20).
// 10 FP64 operations on a dependency chain of 9 (atanh form above: 17 on a chain of 19, plus the reciprocal seed).
constexpr int kLogTab = 128, kLogTabSplit = 53;
__device__ double2 g_logtab[kLogTab];
__constant__ double c_l1p[6] = {-1.0 / 2.0, 1.0 / 3.0, -1.0 / 4.0, 1.0 / 5.0, -1.0 / 6.0, 1.0 / 7.0};

__device__ __forceinline__ double flog_tab(double x, const double2* __restrict__ lt) {
    const int hi = __double2hiint(x);
    const int i = (hi >> 13) & (kLogTab - 1);
    const int e = (hi >> 20) - 1023 + (i >= kLogTabSplit ? 1 : 0);
    const double m = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, __double2loint(x));
    const double2 uv = lt[i];
    const double r = fma(m, uv.x, -1.0);
    double p = c_l1p[5];
#pragma unroll
    for (int k = 4; k >= 0; --k) p = fma(p, r, c_l1p[k]);
    const double ef = (double)e;
    return fma(ef, c_ln2[0], (uv.y + fma(r * r, p, r)) + ef * c_ln2[1]);
}

// 1/(e^v - 1), v > 0
__device__ __forceinline__ double bose_inv(double v) {
    if (v < 0.125) {   // Laurent series: 1/v - 1/2 + v/12 - v^3/720 + v^5/30240 - v^7/1209600 (rel. error < 1e-15 here)
        const double v2 = v * v;
        double p = fma(v2, c_bose[3], c_bose[2]);
        p = fma(p, v2, c_bose[1]);
        p = fma(p, v2, c_bose[0]);
        return fma(p, v, frcp(v) - 0.5);
    }
    const double e = fexp(-v);
    return e * frcp(1.0 - e);
}

enum { KIND_IC_PHOTON = 0, KIND_IC_ELECTRON = 1, KIND_PC_ELECTRON = 2 };

struct PairConst {   // per-(E,E') constants of the three integrands
    double E, Ep, dE;
    double c9;        // r^2/(2+2r), r = E/(E'-E)
    double c10;       // me^2/(4 E')
    double i2Ep;      // 1/(2 E')
#if ACRO_FQ_LOGTAB
    const double2* lt;   // log table (shared-memory copy of g_logtab where the kernel stages one)
#endif
};

#if ACRO_FQ_LOGTAB
#define ACRO_NODE_LOG(c, v) flog_tab((v), (c).lt)
#else
#define ACRO_NODE_LOG(c, v) flog(v)
#endif

// integrand WITHOUT the Bose factor and without the power of x that goes with it:
//   IC kernels:  F(...)            (the measure supplies x^2 dlogx = x dx)
//   pair kernel: G(...)            (the measure supplies x dlogx = dx)
template <int KIND>
__device__ __forceinline__ double integrand_core(const PairConst& c, double x, double x_a, double lnx_minus_lnxa /* log(x/x_a) */,
                                                 bool have_log) {
    if (KIND == KIND_IC_PHOTON) {
        const double q = x_a * frcp(x);
        const double lq = have_log ? -lnx_minus_lnxa : ACRO_NODE_LOG(c, q);
        const double omq = 1.0 - q;
        return fma(2.0 * q, lq, fma(fma(2.0, q, 1.0), omq, c.c9 * omq));
    } else if (KIND == KIND_IC_ELECTRON) {
        const double A = c.dE + x, B = c.E - x;
        const double ixb = frcp(x * B);
        const double q = c.c10 * A * ixb;
        const double omq = 1.0 - q;
        return fma(2.0 * q, ACRO_NODE_LOG(c, q), fma(fma(2.0, q, 1.0), omq, A * A * omq * c.i2Ep * (x * ixb)));
    } else {
        const double s = c.Ep + x, Eep = s - c.E, ee = c.E * Eep;
        const double R = frcp(ee * s * x);
        const double iee = R * s * x, is = R * ee * x, ix = R * ee * s;
        const double s2 = s * s;
        double sud = 4.0 * s2 * ACRO_NODE_LOG(c, 4.0 * x * ee * is * (1.0 / kMe2)) * iee;
        sud = fma(fma(kMe2 * ix, is, -1.0), (s2 * s2) * (iee * iee), sud);
        sud = fma(2.0 * fma(2.0 * x, s, -kMe2), s2 * iee * (1.0 / kMe2), sud);
        return fma(-8.0 * (1.0 / kMe2), x * s, sud);
    }
}

// int_{x_a}^{x_b} core(x) * x^{P-1} / (pi^2 (e^{x/T}-1)) dx   with P = 2 (IC) or 1 (pair creation)
// STEEP_ONLY: the caller guarantees x_a > e T (kernels_wien_kernel: x_a > wien_from T, wien_from >= e), so the plain
// Gauss-Legendre branch is not instantiated
//
// x_kin (kernels_wien_kernel only): the kinematic upper limit of the integrand, x_kin >= x_b.  Where the range ends at the
// thermal cut-off x_b = Ephb_T_max * T < x_kin (so x_a/T > Ephb_T_max - 32 = 168) the integrand stays analytic beyond x_b, and
//     int_0^S f(s) e^{-s} ds  =  L[f from x_a] - e^{-S} L[f from x_b]
// with the 4-node Gauss-Laguerre rule L on both (the order the long-span rule uses for x_a/T > e^4) is exact to 2e-13 for
// S >= 2 (tools/proto/wien_short_span.py: 1500 random pairs per integral); for S < 2 an 8-node Gauss-Legendre rule on [0, S]
// reaches 7e-13.  8 integrand evaluations instead of 32 -- these combinations are 4 % of the items of kernels_wien_kernel in
// the benchmark scan but were 14 % of its nodes, run at half-empty warps.
constexpr double kLag4Top = 9.395070912301133;   // largest node of the 4-point Gauss-Laguerre rule
constexpr double kCutSplit = 2.0;

template <int KIND, int Q, bool STEEP_ONLY = false>
__device__ __forceinline__ double thermal_integral(const PairConst& c, double T, double x_a, double x_b, double x_kin = -1.0) {
    const double iT = frcp(T);
    const double va = x_a * iT;
    if ((STEEP_ONLY || va > kSteepX) && (x_b - x_a) * iT >= kSteepSpan) {
        // Gauss-Laguerre in s = (x - x_a)/T.  The further out on the Wien tail, the smoother the integrand on the scale of
        // the weight, so the order follows va = x_a/T (worst relative error of each band in tools/proto/wien_orders.py, all
        // <= 4e-10):   va in (e, e^1.5]: 24   (e^1.5, e^2]: 16   (e^2, e^2.5]: 10   (e^2.5, e^3]: 8   (e^3, e^4]: 6   > e^4: 4
        const double Ea = fexp(-va);
        int n, off;
        if (va > 54.598150033144236) { n = 4; off = 0; }
        else if (va > 20.085536923187668) { n = 6; off = 4; }
        else if (va > 12.182493960703473) { n = 8; off = 10; }
        else if (va > 7.38905609893065) { n = 10; off = 70; }
        else if (va > 4.4816890703380645) { n = 16; off = 30; }
        else { n = 24; off = 46; }
        double sum = 0.0;
#if ACRO_FQ_LAG_UNROLL > 1
        // all orders are even: two nodes per trip as independent chains (branch-free: a node beyond x_b is computed and
        // dropped by a select, whatever its value)
        double sum2 = 0.0;
#pragma unroll 1
        for (int kk = 0; kk < n; kk += 2) {
            const int k = off + kk;
            const double x0 = fma(c_lag_s[k], T, x_a), x1 = fma(c_lag_s[k + 1], T, x_a);
            double g0 = integrand_core<KIND>(c, x0, x_a, 0.0, false), g1 = integrand_core<KIND>(c, x1, x_a, 0.0, false);
            if (KIND != KIND_PC_ELECTRON) { g0 *= x0; g1 *= x1; }
            const double a0 = fma(c_lag_w[k] * g0, frcp(fma(-Ea, c_lag_es[k], 1.0)), sum);
            const double a1 = fma(c_lag_w[k + 1] * g1, frcp(fma(-Ea, c_lag_es[k + 1], 1.0)), sum2);
            sum = x0 < x_b ? a0 : sum;
            sum2 = x1 < x_b ? a1 : sum2;
        }
        sum += sum2;
#else
#pragma unroll 1
        for (int kk = 0; kk < n; ++kk) {
            const int k = off + kk;
            const double x = fma(c_lag_s[k], T, x_a);
            if (x < x_b) {
                double g = integrand_core<KIND>(c, x, x_a, 0.0, false);
                if (KIND != KIND_PC_ELECTRON) g *= x;
                sum = fma(c_lag_w[k] * g, frcp(fma(-Ea, c_lag_es[k], 1.0)), sum);
            }
        }
#endif
        return sum * T * Ea * (1.0 / kPi2);
    }
    if (STEEP_ONLY || va > kSteepX) {
        // Wien tail with a short span S = (x_b - x_a)/T < 32: Gauss-Legendre (32 nodes) in s on [0, S] with the explicit
        // weight e^{-s}; no logarithmic variable needed because the integrand varies on the scale x_a/T > e.
        const double S = (x_b - x_a) * iT;
        const double Ea = fexp(-va);
        const double hs = 0.5 * S;
        double sum = 0.0;
        if (STEEP_ONLY && x_b < x_kin && va > 54.598150033144236) {
            if (S >= kCutSplit && fma(kLag4Top, T, x_b) < x_kin) {
                const double eS = fexp(-S), Eb = Ea * eS;
                double sum_b = 0.0;
#pragma unroll 1
                for (int k = 0; k < 4; ++k) {
                    const double x0 = fma(c_lag_s[k], T, x_a), x1 = fma(c_lag_s[k], T, x_b);
                    double g0 = integrand_core<KIND>(c, x0, x_a, 0.0, false), g1 = integrand_core<KIND>(c, x1, x_a, 0.0, false);
                    if (KIND != KIND_PC_ELECTRON) { g0 *= x0; g1 *= x1; }
                    sum = fma(c_lag_w[k] * g0, frcp(fma(-Ea, c_lag_es[k], 1.0)), sum);
                    sum_b = fma(c_lag_w[k] * g1, frcp(fma(-Eb, c_lag_es[k], 1.0)), sum_b);
                }
                return fma(-eS, sum_b, sum) * T * Ea * (1.0 / kPi2);
            }
            if (S < kCutSplit) {
#pragma unroll 1
                for (int k = 0; k < 8; ++k) {
                    const double sk = fma(hs, c_glx[gl_offset(8) + k], hs);
                    const double x = fma(sk, T, x_a);
                    const double es = fexp(-sk);
                    double g = integrand_core<KIND>(c, x, x_a, 0.0, false);
                    if (KIND != KIND_PC_ELECTRON) g *= x;
                    sum = fma(c_glw[gl_offset(8) + k] * g * es, frcp(fma(-Ea, es, 1.0)), sum);
                }
                return sum * hs * T * Ea * (1.0 / kPi2);
            }
        }
#pragma unroll 1
        for (int k = 0; k < 32; ++k) {
            const double sk = fma(hs, c_glx[gl_offset(32) + k], hs);
            const double x = fma(sk, T, x_a);
            const double es = fexp(-sk);
            double g = integrand_core<KIND>(c, x, x_a, 0.0, false);
            if (KIND != KIND_PC_ELECTRON) g *= x;
            sum = fma(c_glw[gl_offset(32) + k] * g * es, frcp(fma(-Ea, es, 1.0)), sum);
        }
        return sum * hs * T * Ea * (1.0 / kPi2);
    }
    const double a = flog(x_a), b = flog(x_b);
    const double h = 0.5 * (b - a), m = 0.5 * (b + a);
    // node count by the width of the log-range: the Q-point rule, or Q/2 points when the range is short.  Measured on a
    // 168-point slice sample against Q=96: with ranges <= 8 e-folds at Q/2 = 32 the photon (a9) and pair (a11) kernels
    // stay within 1e-9; the electron kernel (a10) does not (2e-7 already at 4.5 e-folds) and always uses Q.
    const bool half = (Q >= 32) && (KIND != KIND_IC_ELECTRON) && (b - a <= kHalfRange);
    const int n = half ? Q / 2 : Q;
    const double* gx = c_glx + (half ? gl_offset(Q / 2 < 8 ? 8 : Q / 2) : gl_offset(Q));
    const double* gw = c_glw + (half ? gl_offset(Q / 2 < 8 ? 8 : Q / 2) : gl_offset(Q));
    double sum = 0.0;
    ACRO_UNROLL(ACRO_FQ_UNROLL)
    for (int k = 0; k < n; ++k) {
        const double y = fma(h, gx[k], m);
        const double x = fexp(y);
        double g = integrand_core<KIND>(c, x, x_a, y - a, true) * x;
        if (KIND != KIND_PC_ELECTRON) g *= x;
        sum = fma(gw[k] * g, bose_inv(x * iT), sum);
    }
    return sum * h * (1.0 / kPi2);
}

}  // namespace acro
