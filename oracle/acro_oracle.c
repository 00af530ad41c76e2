/*
 * acro_oracle.c -- CPU restatement of the ACROPOLIS v1.3.1 hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this; the product
 * (acropolis_b200/) never does.  It follows the reference's algorithm function by function (file:line cited at each
 * definition, paths relative to hep-mh/acropolis): scalar integrands, ADAPTIVE quadrature at relative tolerance `eps`
 * (the reference: scipy.integrate.quad/dblquad = QUADPACK QAGS, epsrel=eps=1e-3, epsabs=0), the sequential
 * back-substitution, the log-log interpolants, and Pade-13 scaling-and-squaring for the matrix exponential.
 *
 * Third-party arithmetic that is NOT under /root/reference and is restated from its published algorithm:
 *   - QUADPACK (SciPy 1.18.1 in the build container; reference pins only scipy>=1.5.2, setup.py:42-46): 21-point
 *     Gauss-Kronrod rule with QUADPACK's error heuristic (qk21) and globally adaptive bisection of the worst
 *     interval (qag/qags).  QAGS's Wynn-epsilon extrapolation is NOT restated: for the analytic integrands of this
 *     path it never changes the accepted result beyond the tolerance; tests/test_oracle.py pins the difference to
 *     scipy on the reference's own integrands and to the golden fixtures generated from the reference.
 *   - LAPACK dgesv (3x3): Gaussian elimination with partial pivoting.
 *   - scipy.linalg.expm: Pade-13 scaling and squaring (Higham 2005).
 * Parity is pinned: see tests/test_oracle.py (golden vectors from tests/golden/make_golden.py, and the reference's own
 * tools/data/ (.dat) rows).
 *
 * Build: make -C oracle  ->  oracle/liboracle.so  (gcc -O2 -pthread; libgomp is not in this image, so the scan driver
 * uses a plain pthread work queue).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <pthread.h>
#include <unistd.h>

/* ---- constants (acropolis/params.py:15-49) ---- */
#define ALPHA (1.0 / 137.036)
#define ME 0.511
#define ME2 (0.511 * 0.511)
#define RE (ALPHA / ME)
#define PI 3.14159265358979323846
#define PI2 (PI * PI)
#define ZETA3 1.2020569031595942
#define HBAR 6.582119514e-22
#define TAU_N 8.794e2
#define TAU_T 5.605e8
#define NX 3
#define NY 9
#define NR 17

typedef struct {
    /* tables */
    const double* ratedb; /* [nT*nE][2] log10 rates (db.py:95-110) */
    int nT, nE;
    double Elog_min, Elog_max, Tlog_min, Tlog_max;
    const double* cosmo_lT; /* log10 T column, as stored (descending) */
    const double* cosmo_ld; /* log10 |dT/dt| column */
    int n_cosmo;
    double eta, Yp, YHe4;
    double Y0[27];
    /* algorithm parameters (params.py:78-128) */
    double eps, Emin, approx_zero, Ephb_T_max, E_EC_max;
    int use_db;
    long neval; /* integrand evaluations (statistics) */
} oracle_t;

/* ------------------------------------------------------------------------------------------------------------
 * adaptive Gauss-Kronrod 21 (QUADPACK qk21 + qag-style bisection), epsabs = 0
 * ---------------------------------------------------------------------------------------------------------- */
typedef double (*integrand_t)(double x, void* ctx);

static const double XGK[11] = {0.995657163025808080735527280689003, 0.973906528517171720077964012084452,
                               0.930157491355708226001207180059508, 0.865063366688984510732096688423493,
                               0.780817726586416897063717578345042, 0.679409568299024406234327365114874,
                               0.562757134668604683339000099272694, 0.433395394129247190799265943165784,
                               0.294392862701460198131126603103866, 0.148874338981631210884826001129720, 0.0};
static const double WGK[11] = {0.011694638867371874278064396062192, 0.032558162307964727478818972459390,
                               0.054755896574351996031381300244580, 0.075039674810919952767043140916190,
                               0.093125454583697605535065465083366, 0.109387158802297641899210590325805,
                               0.123491976262065851077958109585166, 0.134709217311473325928054001771707,
                               0.142775938577060080797094273138717, 0.147739104901338491374841515972068,
                               0.149445554002916905664936468389821};
static const double WG[5] = {0.066671344308688137593568809893332, 0.149451349150580593145776339657697,
                             0.219086362515982043995534934228163, 0.269266719309996355091226921569469,
                             0.295524224714752870173815619188769};

static void qk21(integrand_t f, void* ctx, double a, double b, double* result, double* abserr, long* neval) {
    const double epmach = 2.220446049250313e-16, uflow = 2.2250738585072014e-308;
    const double centr = 0.5 * (a + b), hlgth = 0.5 * (b - a), dhlgth = fabs(hlgth);
    double fv1[10], fv2[10];
    const double fc = f(centr, ctx);
    double resg = 0.0, resk = WGK[10] * fc, resabs = fabs(resk);
    for (int j = 0; j < 5; ++j) {
        const int jtw = 2 * j + 1;
        const double absc = hlgth * XGK[jtw];
        const double f1 = f(centr - absc, ctx), f2 = f(centr + absc, ctx);
        fv1[jtw] = f1; fv2[jtw] = f2;
        resg += WG[j] * (f1 + f2);
        resk += WGK[jtw] * (f1 + f2);
        resabs += WGK[jtw] * (fabs(f1) + fabs(f2));
    }
    for (int j = 0; j < 5; ++j) {
        const int jtwm1 = 2 * j;
        const double absc = hlgth * XGK[jtwm1];
        const double f1 = f(centr - absc, ctx), f2 = f(centr + absc, ctx);
        fv1[jtwm1] = f1; fv2[jtwm1] = f2;
        resk += WGK[jtwm1] * (f1 + f2);
        resabs += WGK[jtwm1] * (fabs(f1) + fabs(f2));
    }
    const double reskh = resk * 0.5;
    double resasc = WGK[10] * fabs(fc - reskh);
    for (int j = 0; j < 10; ++j) resasc += WGK[j] * (fabs(fv1[j] - reskh) + fabs(fv2[j] - reskh));
    *result = resk * hlgth;
    resabs *= dhlgth;
    resasc *= dhlgth;
    double err = fabs((resk - resg) * hlgth);
    if (resasc != 0.0 && err != 0.0) {
        const double t = pow(200.0 * err / resasc, 1.5);
        err = resasc * (t < 1.0 ? t : 1.0);
    }
    if (resabs > uflow / (50.0 * epmach)) {
        const double t = 50.0 * epmach * resabs;
        if (t > err) err = t;
    }
    *abserr = err;
    *neval += 21;
}

/* returns the integral; *ier = 1 if `limit` subdivisions were not enough (scipy would warn, nucl.py:453) */
static double quad(oracle_t* o, integrand_t f, void* ctx, double a, double b, double epsrel, int limit, int* ier) {
    enum { MAXLIM = 400 };
    double alist[MAXLIM], blist[MAXLIM], rlist[MAXLIM], elist[MAXLIM];
    if (limit > MAXLIM) limit = MAXLIM;
    if (ier) *ier = 0;
    double sign = 1.0;
    if (b < a) { const double t = a; a = b; b = t; sign = -1.0; }
    if (a == b) return 0.0;
    long ne = 0;
    qk21(f, ctx, a, b, &rlist[0], &elist[0], &ne);
    alist[0] = a; blist[0] = b;
    int n = 1;
    double result = rlist[0], errsum = elist[0];
    while (errsum > epsrel * fabs(result) && n < limit) {
        int w = 0;
        for (int i = 1; i < n; ++i) if (elist[i] > elist[w]) w = i;
        const double a1 = alist[w], b2 = blist[w], mid = 0.5 * (a1 + b2);
        if (!(a1 < mid && mid < b2)) break;   /* interval exhausted in floating point */
        double r1, e1, r2, e2;
        qk21(f, ctx, a1, mid, &r1, &e1, &ne);
        qk21(f, ctx, mid, b2, &r2, &e2, &ne);
        alist[w] = a1; blist[w] = mid; rlist[w] = r1; elist[w] = e1;
        alist[n] = mid; blist[n] = b2; rlist[n] = r2; elist[n] = e2;
        ++n;
        result = 0.0; errsum = 0.0;
        for (int i = 0; i < n; ++i) { result += rlist[i]; errsum += elist[i]; }
    }
    if (errsum > epsrel * fabs(result) && ier) *ier = 1;
    o->neval += ne;
    return sign * result;
}

/* ------------------------------------------------------------------------------------------------------------
 * integrand functions (cascade.py:28-164)
 * ---------------------------------------------------------------------------------------------------------- */
/* _JIT_F, cascade.py:29-43 */
static double fn_F(double Eph, double Ee, double Ephb) {
    if (!(Ephb <= Eph && Eph <= 4.0 * Ephb * Ee * Ee / (ME2 + 4.0 * Ephb * Ee))) return 0.0;
    const double G = 4.0 * Ephb * Ee / ME2;
    const double q = Eph / (G * (Ee - Eph));
    return 2.0 * q * log(q) + (1.0 + 2.0 * q) * (1.0 - q) + pow(G * q, 2.0) * (1.0 - q) / (2.0 + 2.0 * G * q);
}

/* _JIT_G, cascade.py:47-80 */
static double fn_G(double Ee, double Eph, double Ephb) {
    const double Eep = Eph + Ephb - Ee;
    const double dE_sqrt = (Eph - Ephb) * sqrt(1.0 - ME2 / (Eph * Ephb));
    const double lim_m = (Eph + Ephb - dE_sqrt) / 2.0, lim_p = (Eph + Ephb + dE_sqrt) / 2.0;
    if (!(ME < lim_m && lim_m <= Ee && Ee <= lim_p)) return 0.0;
    const double s = Ee + Eep;
    double sud = 0.0;
    sud += 4.0 * (s * s) * log((4.0 * Ephb * Ee * Eep) / (ME2 * s)) / (Ee * Eep);
    sud += (ME2 / (Ephb * s) - 1.0) * pow(s, 4.0) / ((Ee * Ee) * (Eep * Eep));
    sud += 2.0 * (2.0 * Ephb * s - ME2) * (s * s) / (ME2 * Ee * Eep);
    sud += -8.0 * Ephb * s / ME2;
    return sud;
}

/* _JIT_dsdE_Z2, cascade.py:139-164 */
static double fn_dsdE_Z2(double Ee, double Eph) {
    const double Em = Ee, Ep = Eph - Ee;
    const double pm = sqrt(Em * Em - ME2), pp = sqrt(Ep * Ep - ME2);
    const double L = log((Ep * Em + pp * pm + ME2) / (Ep * Em - pp * pm + ME2));
    const double lm = log((Em + pm) / (Em - pm)), lp = log((Ep + pp) / (Ep - pp));
    const double pref = ALPHA * (RE * RE) * pp * pm / pow(Eph, 3.0);
    double sud = 0.0;
    sud += -4.0 / 3.0 - 2.0 * Ep * Em * (pp * pp + pm * pm) / ((pp * pp) * (pm * pm));
    sud += ME2 * (lm * Ep / pow(pm, 3.0) + lp * Em / pow(pp, 3.0) - lp * lm / (pp * pm));
    sud += L * (-8.0 * Ep * Em / (3.0 * pp * pm) +
                Eph * Eph * (pow(Ep * Em, 2.0) + pow(pp * pm, 2.0) - ME2 * Ep * Em) / (pow(pp, 3.0) * pow(pm, 3.0)));
    sud += -L * ME2 * Eph * (lp * (Ep * Em - pp * pp) / pow(pp, 3.0) + lm * (Ep * Em - pm * pm) / pow(pm, 3.0)) /
           (2.0 * pp * pm);
    return pref * sud;
}

/* densities, cascade.py:258-271 */
static double n_b(const oracle_t* o, double T) { return o->eta * (2.0 * ZETA3 / PI2) * pow(T, 3.0); }
static double n_e(const oracle_t* o, double T) { return (o->Yp + 2.0 * o->YHe4) * n_b(o, T); }
static double n_NZ2(const oracle_t* o, double T) { return (o->Yp + 4.0 * o->YHe4) * n_b(o, T); }

/* ---- rates.db (db.py:50-124) ---- */
static int in_rate_db(const oracle_t* o, double E_log, double T_log) {
    return (o->Elog_min <= E_log && E_log <= o->Elog_max) && (o->Tlog_min <= T_log && T_log <= o->Tlog_max);
}

static double interp_rate_db(const oracle_t* o, int c, double E_log, double T_log) {
    const int Enum = o->nE, Tnum = o->nT;
    int iE = (int)((Enum - 1) * (E_log - o->Elog_min) / (o->Elog_max - o->Elog_min));
    int iT = (int)((Tnum - 1) * (T_log - o->Tlog_min) / (o->Tlog_max - o->Tlog_min));
    if (iE == Enum - 1) iE -= 1;
    if (iT == Tnum - 1) iT -= 1;
    const double x = T_log, y = E_log;
    const double x0 = o->Tlog_min + (o->Tlog_max - o->Tlog_min) * iT / (Tnum - 1);
    const double x1 = o->Tlog_min + (o->Tlog_max - o->Tlog_min) * (iT + 1) / (Tnum - 1);
    const double y0 = o->Elog_min + (o->Elog_max - o->Elog_min) * iE / (Enum - 1);
    const double y1 = o->Elog_min + (o->Elog_max - o->Elog_min) * (iE + 1) / (Enum - 1);
    const double c00 = o->ratedb[((size_t)iT * Enum + iE) * 2 + c];
    const double c10 = o->ratedb[((size_t)(iT + 1) * Enum + iE) * 2 + c];
    const double c01 = o->ratedb[((size_t)iT * Enum + iE + 1) * 2 + c];
    const double c11 = o->ratedb[((size_t)(iT + 1) * Enum + iE + 1) * 2 + c];
    const double d = (x0 - x1) * (y0 - y1);
    const double a0 = (c00 * x1 * y1 - c01 * x1 * y0 - c10 * x0 * y1 + c11 * x0 * y0) / d;
    const double a1 = (-c00 * y1 + c01 * y0 + c10 * y1 - c11 * y0) / d;
    const double a2 = (-c00 * x1 + c01 * x1 + c10 * x0 - c11 * x0) / d;
    const double a3 = (c00 - c01 - c10 + c11) / d;
    return pow(10.0, a0 + a1 * x + a2 * y + a3 * x * y);
}

/* ---- photon rates (cascade.py:285-374) ---- */
static double rate_photon_photon(double E, double T) {
    return 0.151348 * pow(ALPHA, 4.0) * ME * pow(E / ME, 3.0) * pow(T / ME, 6.0) * exp(-E * T / ME2);
}

static double rate_compton(const oracle_t* o, double E, double T) {
    const double x = 2.0 * E / ME;
    return (2.0 * PI * (RE * RE) / x) * n_e(o, T) *
           ((1.0 - 4.0 / x - 8.0 / (x * x)) * log(1.0 + x) + 0.5 + 8.0 / x - 1.0 / (2.0 * pow(1.0 + x, 2.0)));
}

static double rate_bethe_heitler(const oracle_t* o, double E, double T) {
    const double k = E / ME;
    if (k < 2.0) return 0.0;
    if (k <= 4.0) {
        const double r = (2.0 * k - 4.0) / (k + 2.0 + 2.0 * sqrt(2.0 * k));
        return (pow(ALPHA, 3.0) / ME2) * n_NZ2(o, T) * (2.0 * PI / 3.0) * pow((k - 2.0) / k, 3.0) *
               (1.0 + r / 2.0 + (23.0 / 40.0) * (r * r) + (11.0 / 60.0) * pow(r, 3.0) + (29.0 / 960.0) * pow(r, 4.0));
    }
    const double l = log(2.0 * k);
    return (pow(ALPHA, 3.0) / ME2) * n_NZ2(o, T) *
           ((28.0 / 9.0) * l - 218.0 / 27.0 +
            pow(2.0 / k, 2.0) * ((2.0 / 3.0) * pow(l, 3.0) - l * l + (6.0 - PI2 / 3.0) * l + 2.0 * ZETA3 + PI2 / 6.0 - 7.0 / 2.0) -
            pow(2.0 / k, 4.0) * ((3.0 / 16.0) * l + 1.0 / 8.0) - pow(2.0 / k, 6.0) * ((29.0 / 2304.0) * l - 77.0 / 13824.0));
}

/* _JIT_ph_rate_pair_creation_ae + dblquad, cascade.py:86-103, 336-358 */
typedef struct { oracle_t* o; double E, T, logx; } pc_ctx_t;
static double pc_inner(double logy, void* vctx) {
    pc_ctx_t* c = (pc_ctx_t*)vctx;
    const double x = exp(c->logx), y = exp(logy);
    const double b = sqrt(1.0 - 4.0 * ME2 / y);
    const double sig = 0.5 * PI * (RE * RE) * (1.0 - b * b) *
                       ((3.0 - pow(b, 4.0)) * log((1.0 + b) / (1.0 - b)) - 2.0 * b * (2.0 - b * b));
    return (1.0 / (PI * PI)) / (exp(x / c->T) - 1.0) * y * sig * (x * y);
}
static double pc_outer(double logx, void* vctx) {
    pc_ctx_t* c = (pc_ctx_t*)vctx;
    pc_ctx_t in = *c;
    in.logx = logx;
    return quad(c->o, pc_inner, &in, log(4.0 * ME2), log(4.0 * c->E) + logx, c->o->eps, 50, NULL);
}
static double rate_pair_creation_ae(oracle_t* o, double E, double T) {
    if (E < ME2 / (50.0 * T)) return 0.0;
    pc_ctx_t c = {o, E, T, 0.0};
    const double I = quad(o, pc_outer, &c, log(ME2 / E), log(o->Ephb_T_max * T), o->eps, 50, NULL);
    return I / (8.0 * E * E);
}
static double rate_pair_creation_ae_db(oracle_t* o, double E, double T) {
    if (E < ME2 / (50.0 * T)) return 0.0;
    const double E_log = log10(E), T_log = log10(T);
    if (!o->use_db || !in_rate_db(o, E_log, T_log)) return rate_pair_creation_ae(o, E, T);
    return interp_rate_db(o, 0, E_log, T_log);
}
static double total_rate_photon(oracle_t* o, double E, double T) {
    return rate_photon_photon(E, T) + rate_compton(o, E, T) + rate_bethe_heitler(o, E, T) + rate_pair_creation_ae_db(o, E, T);
}

/* ---- electron rate: _JIT_el_rate_inverse_compton + dblquad, cascade.py:117-119, 469-498 ---- */
typedef struct { oracle_t* o; double E, T, x; } ic_ctx_t;
static double icr_inner(double y, void* vctx) {
    ic_ctx_t* c = (ic_ctx_t*)vctx;
    return fn_F(y, c->E, c->x) * c->x / ((PI * PI) * (exp(c->x / c->T) - 1.0));
}
static double icr_outer(double x, void* vctx) {
    ic_ctx_t* c = (ic_ctx_t*)vctx;
    if (x <= 0.0) return 0.0;
    ic_ctx_t in = *c;
    in.x = x;
    return quad(c->o, icr_inner, &in, x, 4.0 * x * c->E * c->E / (ME2 + 4.0 * x * c->E), c->o->eps, 50, NULL);
}
static double rate_inverse_compton(oracle_t* o, double E, double T) {
    const double ulim = fmin(E - ME2 / (4.0 * E), o->Ephb_T_max * T);
    ic_ctx_t c = {o, E, T, 0.0};
    const double I = quad(o, icr_outer, &c, 0.0, ulim, o->eps, 50, NULL);
    return 2.0 * PI * (ALPHA * ALPHA) * I / (E * E);
}
static double total_rate_electron(oracle_t* o, double E, double T) {
    const double E_log = log10(E), T_log = log10(T);
    if (!o->use_db || !in_rate_db(o, E_log, T_log)) return rate_inverse_compton(o, E, T);
    return interp_rate_db(o, 1, E_log, T_log);
}

/* ---- kernels (cascade.py:383-455, 507-598, 621-700) ---- */
static double kernel_photon_photon(double E, double Ep, double T) {
    return 1112.0 / (10125.0 * PI) * pow(ALPHA, 4.0) / pow(ME, 8.0) * 8.0 * pow(PI, 4.0) * pow(T, 6.0) / 63.0 * (Ep * Ep) *
           pow(1.0 - E / Ep + pow(E / Ep, 2.0), 2.0) * exp(-Ep * T / ME2);
}
static double kernel_compton_formula(const oracle_t* o, double E, double Ep, double T) {
    if (Ep / (1.0 + 2.0 * Ep / ME) > E) return 0.0;
    return PI * (RE * RE) * ME / (Ep * Ep) * n_e(o, T) *
           (Ep / E + E / Ep + pow(ME / E - ME / Ep, 2.0) - 2.0 * ME * (1.0 / E - 1.0 / Ep));
}
typedef struct { double E, Ep, T; } k_ctx_t;
static double ick_ph_integrand(double logx, void* v) {   /* cascade.py:107-111 */
    k_ctx_t* c = (k_ctx_t*)v;
    const double x = exp(logx);
    return fn_F(c->E, c->Ep, x) * x / (PI2 * (exp(x / c->T) - 1.0)) * x;
}
static double kernel_ic_photon(oracle_t* o, double E, double Ep, double T) {   /* cascade.py:406-435 */
    if (Ep < E + ME) return 0.0;
    const double llim = 0.25 * ME2 * E / (Ep * Ep - Ep * E);
    const double ulim = fmin(fmin(E, Ep - ME2 / (4.0 * Ep)), o->Ephb_T_max * T);
    if (ulim <= llim) return 0.0;
    k_ctx_t c = {E, Ep, T};
    return 2.0 * PI * (ALPHA * ALPHA) * quad(o, ick_ph_integrand, &c, log(llim), log(ulim), o->eps, 50, NULL) / (Ep * Ep);
}
static double ick_el_integrand(double logx, void* v) {   /* cascade.py:123-127 */
    k_ctx_t* c = (k_ctx_t*)v;
    const double x = exp(logx);
    return fn_F(c->Ep + x - c->E, c->Ep, x) * (x / (PI * PI)) / (exp(x / c->T) - 1.0) * x;
}
static double kernel_ic_electron(oracle_t* o, double E, double Ep, double T) {   /* cascade.py:507-540 */
    if (E == Ep) return 0.0;
    const double pf = 0.25 * ME2 / Ep - E, qf = 0.25 * ME2 * (Ep - E) / Ep;
    const double sqrt_d = sqrt(pow(pf / 2.0, 2.0) - qf);
    const double z1 = -pf / 2.0 - sqrt_d, z2 = -pf / 2.0 + sqrt_d;
    const double llim = z1, ulim = fmin(fmin(z2, Ep - ME2 / (4.0 * Ep)), o->Ephb_T_max * T);
    if (ulim <= llim) return 0.0;
    k_ctx_t c = {E, Ep, T};
    return 2.0 * PI * (ALPHA * ALPHA) * quad(o, ick_el_integrand, &c, log(llim), log(ulim), o->eps, 50, NULL) / (Ep * Ep);
}
static double pck_integrand(double logx, void* v) {   /* cascade.py:131-135 */
    k_ctx_t* c = (k_ctx_t*)v;
    const double x = exp(logx);
    return fn_G(c->E, c->Ep, x) / ((PI * PI) * (exp(x / c->T) - 1.0)) * x;
}
static double kernel_pair_creation_ae(oracle_t* o, double E, double Ep, double T) {   /* cascade.py:562-598 */
    if (Ep < ME2 / (50.0 * T)) return 0.0;
    const double dE = Ep - E, E2 = E * E;
    const double z1 = Ep * (ME2 - 2.0 * dE * (sqrt(E2 - ME2) - E)) / (4 * Ep * dE + ME2);
    const double z2 = Ep * (ME2 + 2.0 * dE * (sqrt(E2 - ME2) + E)) / (4 * Ep * dE + ME2);
    const double llim = fmax(ME2 / Ep, z1), ulim = fmin(fmin(z2, Ep), o->Ephb_T_max * T);
    if (ulim <= llim) return 0.0;
    k_ctx_t c = {E, Ep, T};
    return 0.25 * PI * (ALPHA * ALPHA) * ME2 * quad(o, pck_integrand, &c, log(llim), log(ulim), o->eps, 50, NULL) / pow(Ep, 3.0);
}
static double kernel_bethe_heitler(const oracle_t* o, double E, double Ep, double T) {   /* cascade.py:550-558 */
    if (Ep < E + ME) return 0.0;
    return n_NZ2(o, T) * fn_dsdE_Z2(E, Ep);
}

/* the five distinct planes of K[X,Xp] (cascade.py:439-455, 644-660, 684-700) */
static void kernel_planes(oracle_t* o, double E, double Ep, double T, double k[5]) {
    k[0] = kernel_photon_photon(E, Ep, T) + kernel_compton_formula(o, E, Ep, T);   /* gamma -> gamma */
    k[1] = kernel_ic_photon(o, E, Ep, T);                                           /* e+- -> gamma   */
    const double bh = kernel_bethe_heitler(o, E, Ep, T), pc = kernel_pair_creation_ae(o, E, Ep, T);
    k[2] = kernel_compton_formula(o, Ep + ME - E, Ep, T) + bh + pc;                 /* gamma -> e-    */
    k[3] = 0.0 + bh + pc;                                                           /* gamma -> e+    */
    k[4] = kernel_ic_electron(o, E, Ep, T);                                         /* e+- -> e+-     */
}

/* ------------------------------------------------------------------------------------------------------------
 * cascade equation (cascade.py:180-239, 733-771)
 * ---------------------------------------------------------------------------------------------------------- */
static void solve3(double B[3][3], double a[3], double x[3]) {   /* dgesv: LU with partial pivoting */
    int p[3] = {0, 1, 2};
    for (int c = 0; c < 3; ++c) {
        int piv = c;
        for (int r = c + 1; r < 3; ++r) if (fabs(B[p[r]][c]) > fabs(B[p[piv]][c])) piv = r;
        const int t = p[c]; p[c] = p[piv]; p[piv] = t;
        for (int r = c + 1; r < 3; ++r) {
            const double f = B[p[r]][c] / B[p[c]][c];
            for (int cc = c; cc < 3; ++cc) B[p[r]][cc] -= f * B[p[c]][cc];
            a[p[r]] -= f * a[p[c]];
        }
    }
    for (int c = 2; c >= 0; --c) {
        double v = a[p[c]];
        for (int cc = c + 1; cc < 3; ++cc) v -= B[p[c]][cc] * x[cc];
        x[c] = v / B[p[c]][c];
    }
}

/* G[3][NE], K[5][NE][NE] (dense, zero below the diagonal) at temperature T */
int oracle_rates_kernels(oracle_t* o, int NE, const double* E_grid, double T, double* G, double* K) {
    for (int i = 0; i < NE; ++i) {
        G[i] = total_rate_photon(o, E_grid[i], T);
        G[NE + i] = G[2 * NE + i] = total_rate_electron(o, E_grid[i], T);
    }
    if (K) {
        memset(K, 0, sizeof(double) * 5 * NE * NE);
        for (int i = 0; i < NE; ++i)
            for (int j = 0; j < NE; ++j)
                if (E_grid[j] >= E_grid[i]) {
                    double k[5];
                    kernel_planes(o, E_grid[i], E_grid[j], T, k);
                    for (int q = 0; q < 5; ++q) K[((size_t)q * NE + i) * NE + j] = k[q];
                }
    }
    return 0;
}

/* full K[X][Xp] entry from the five planes */
static double Kx(const double* K, int NE, int X, int Xp, int i, int j) {
    static const int plane[3][3] = {{0, 1, 1}, {2, 4, -1}, {3, -1, 4}};
    const int q = plane[X][Xp];
    return q < 0 ? 0.0 : K[((size_t)q * NE + i) * NE + j];
}

/* _JIT_solve_cascade_equation: F[3][NE] (floored at approx_zero) */
void oracle_solve(const oracle_t* o, int NE, const double* E_grid, const double* G, const double* K, const double* S0,
                  const double* SC /* [3][NE] */, double* F) {
    const int last = NE - 1;
    const double dy = log(E_grid[last] / o->Emin) / (NE - 1);
    for (int X = 0; X < NX; ++X) {
        double s = 0.0;
        for (int Xp = 0; Xp < NX; ++Xp) s += Kx(K, NE, X, Xp, last, last) * S0[Xp] / (G[Xp * NE + last] * G[X * NE + last]);
        F[X * NE + last] = SC[X * NE + last] / G[X * NE + last] + s;
    }
    for (int i = last - 1; i >= 0; --i) {
        double B[3][3], a[3], x[3];
        for (int X = 0; X < NX; ++X) {
            for (int Xp = 0; Xp < NX; ++Xp)
                B[X][Xp] = -0.5 * dy * E_grid[i] * Kx(K, NE, X, Xp, i, i) + (X == Xp ? G[X * NE + i] : 0.0);
            a[X] = SC[X * NE + i];
            for (int Xp = 0; Xp < NX; ++Xp) {
                const double kl = Kx(K, NE, X, Xp, i, last);
                a[X] += kl * S0[Xp] / G[Xp * NE + last] + 0.5 * dy * E_grid[last] * kl * F[Xp * NE + last];
                for (int j = i + 1; j < last; ++j) a[X] += dy * E_grid[j] * Kx(K, NE, X, Xp, i, j) * F[Xp * NE + j];
            }
        }
        solve3(B, a, x);
        for (int X = 0; X < NX; ++X) F[X * NE + i] = x[X];
    }
    for (int q = 0; q < NX * NE; ++q) if (F[q] < o->approx_zero) F[q] = o->approx_zero;
}

/* ------------------------------------------------------------------------------------------------------------
 * photodisintegration cross-sections and rates (nucl.py:70-88, 174-183, 209-467)
 * ---------------------------------------------------------------------------------------------------------- */
static const double ETH[NR] = {2.224573, 6.257248, 8.481821, 5.493485, 7.718058, 19.813852, 20.577615, 23.846527, 26.071100,
                               3.698892, 15.794685, 2.467032, 7.249962, 10.948850, 1.586627, 5.605794, 9.304680};

static double generic_expr(double E, double Q, double N, double p1, double p2, double p3) {
    if (E < Q) return 0.0;
    return N * pow(Q, p1) * pow(E - Q, p2) / pow(E, p3);
}
static double exp_term(double E, double N, double Eb, double Ed) { return N * exp(-0.5 * pow((E - Eb) / Ed, 2.0)); }

double oracle_cross_section(int rid, double E) {   /* MeV^-2; rid = 1..17 */
    const double Q = ETH[rid - 1];
    double s = 0.0;
    switch (rid) {
        case 1:
            if (E < Q) return 0.0;
            s = 18.75 * (pow(sqrt(Q * (E - Q)) / E, 3.0) +
                         0.007947 * pow(sqrt(Q * (E - Q)) / E, 2.0) * (pow(sqrt(Q) - sqrt(0.037), 2.0) / (E - Q + 0.037)));
            break;
        case 2: s = generic_expr(E, Q, 9.8, 1.95, 1.65, 3.6); break;
        case 3: s = generic_expr(E, Q, 26.0, 2.6, 2.3, 4.9); break;
        case 4: s = generic_expr(E, Q, 8.88, 1.75, 1.65, 3.4); break;
        case 5: s = generic_expr(E, Q, 16.7, 1.95, 2.3, 4.25); break;
        case 6: s = generic_expr(E, Q, 19.5, 3.5, 1.0, 4.5); break;
        case 7: s = generic_expr(E, Q, 20.7, 3.5, 1.0, 4.5); break;
        case 8: s = generic_expr(E, Q, 10.7, 10.2, 3.4, 13.6); break;
        case 9: s = generic_expr(E, Q, 21.7, 4.0, 3.0, 7.0); break;
        case 10: s = generic_expr(E, Q, 143.0, 2.3, 4.7, 7.0); break;
        case 11:
            s = generic_expr(E, Q, 38.1, 3.0, 2.0, 5.0) *
                (exp_term(E, 3.7, 19.0, 3.5) + exp_term(E, 2.75, 30.0, 3.0) + exp_term(E, 2.2, 43.0, 5.0));
            break;
        case 12: {
            if (E < Q) return 0.0;
            const double Ecm = E - Q;
            const double p = 1.0 + 2.2875 * pow(Ecm, 2.0) - 1.1798 * pow(Ecm, 3.0) + 2.5279 * pow(Ecm, 4.0);
            s = 0.105 * (2371.0 / (E * E)) * exp(-2.5954 / sqrt(Ecm)) * exp(-2.056 * Ecm) * p;
            break;
        }
        case 13:
            if (E < Q) return 0.0;
            s = generic_expr(E, Q, 0.176, 1.51, 0.49, 2.0) + generic_expr(E, Q, 1205.0, 5.5, 5.0, 10.5) +
                0.06 / (1.0 + pow((E - Q - 7.46) / 0.188, 2.0));
            break;
        case 14: s = generic_expr(E, Q, 122.0, 4.0, 3.0, 7.0); break;
        case 15: {
            if (E < Q) return 0.0;
            const double Ecm = E - Q;
            const double p = 1.0 - 0.428 * pow(Ecm, 2.0) + 0.534 * pow(Ecm, 3.0) - 0.115 * pow(Ecm, 4.0);
            if (p < 0.0) return 0.0;
            s = 0.504 * (2371.0 / (E * E)) * exp(-5.1909 / sqrt(Ecm)) * exp(-0.548 * Ecm) * p;
            break;
        }
        case 16: s = generic_expr(E, Q, 32.6, 10.0, 2.0, 12.0) + generic_expr(E, Q, 2.27e6, 8.8335, 13.0, 21.8335); break;
        case 17: s = generic_expr(E, Q, 133.0, 4.0, 3.0, 7.0); break;
        default: return 0.0;
    }
    return 2.56819e-6 * s;
}

/* LogInterp (utils.py:9-68) on precomputed logs, arbitrary base via the caller */
typedef struct { int N; const double* xlog; const double* ylog; } loginterp_t;
static double loginterp_log(const loginterp_t* li, double x_log) {   /* returns log(value) in the grid's base */
    const double xmin = li->xlog[0], xmax = li->xlog[li->N - 1];
    int ix = (int)((x_log - xmin) * (li->N - 1) / (xmax - xmin));
    if (ix == li->N - 1) ix -= 1;
    if (ix < 0) ix = 0;
    if (ix > li->N - 2) ix = li->N - 2;
    const double m = (li->ylog[ix + 1] - li->ylog[ix]) / (li->xlog[ix + 1] - li->xlog[ix]);
    const double b = li->ylog[ix + 1] - m * li->xlog[ix + 1];
    return m * x_log + b;
}

typedef struct { loginterp_t li; int rid; } pdi_ctx_t;
static double pdi_integrand(double log_E, void* v) {   /* Fph_s, nucl.py:429-432 */
    pdi_ctx_t* c = (pdi_ctx_t*)v;
    const double E = exp(log_E);
    return exp(loginterp_log(&c->li, log(E))) * E * oracle_cross_section(c->rid, E);
}

/* _pdi_rates for one temperature (nucl.py:398-467); Fph = F[0][:] ; returns number of quad warnings */
int oracle_pdi_rates(oracle_t* o, int NE, const double* E_grid, const double* Fph, double E0, double T, double S0_photon,
                     double rate_photon_E0, double* rates /* [17] */) {
    const double EC = ME2 / (22.0 * T);
    const double Emax = fmin(E0, o->E_EC_max * EC);
    double* xl = (double*)malloc(sizeof(double) * 2 * NE);
    double* yl = xl + NE;
    for (int i = 0; i < NE; ++i) { xl[i] = log(E_grid[i]); yl[i] = log(Fph[i]); }
    int warn = 0;
    for (int rid = 1; rid <= NR; ++rid) {
        double r = S0_photon * oracle_cross_section(rid, E0) / rate_photon_E0;
        if (Emax > ETH[rid - 1]) {
            pdi_ctx_t c = {{NE, xl, yl}, rid};
            int ier = 0;
            /* limit=50 is scipy's default (nucl.py:451).  In "tight" mode (eps < 1e-6, used by the tests as the converged
             * value of the same integrand) the kinked interpolant needs more subdivisions than that. */
            r += quad(o, pdi_integrand, &c, log(ETH[rid - 1]), log(Emax), o->eps, o->eps < 1e-6 ? 400 : 50, &ier);
            warn += ier;
        }
        rates[rid - 1] = fmax(o->approx_zero, r);
    }
    free(xl);
    return warn;
}

/* ------------------------------------------------------------------------------------------------------------
 * rate matrix T-integral (nucl.py:515-623), cosmo interpolation (input.py:177-226)
 * ---------------------------------------------------------------------------------------------------------- */
static const signed char R_INIT[NR] = {2, 3, 3, 4, 4, 5, 5, 5, 5, 6, 6, 7, 7, 7, 8, 8, 8};
static const signed char R_FIN[NR][NY] = {
    {1, 1, 0, 0, 0, 0, 0, 0, 0}, {1, 0, 1, 0, 0, 0, 0, 0, 0}, {2, 1, 0, 0, 0, 0, 0, 0, 0}, {0, 1, 1, 0, 0, 0, 0, 0, 0},
    {1, 2, 0, 0, 0, 0, 0, 0, 0}, {0, 1, 0, 1, 0, 0, 0, 0, 0}, {1, 0, 0, 0, 1, 0, 0, 0, 0}, {0, 0, 2, 0, 0, 0, 0, 0, 0},
    {1, 1, 1, 0, 0, 0, 0, 0, 0}, {1, 1, 0, 0, 0, 1, 0, 0, 0}, {0, 0, 0, 0, 0, 0, 0, 0, 0}, {0, 0, 0, 1, 0, 1, 0, 0, 0},
    {1, 0, 0, 0, 0, 0, 1, 0, 0}, {2, 1, 0, 0, 0, 1, 0, 0, 0}, {0, 0, 0, 0, 1, 1, 0, 0, 0}, {0, 1, 0, 0, 0, 0, 1, 0, 0},
    {1, 2, 0, 0, 0, 1, 0, 0, 0}};
static const signed char D_INIT[2] = {0, 3};
static const signed char D_FIN[2][NY] = {{0, 1, 0, 0, 0, 0, 0, 0, 0}, {0, 0, 0, 0, 1, 0, 0, 0, 0}};
static const double D_TAU[2] = {TAU_N, TAU_T};

static double pref_ij(int init, const signed char* fin, int i, int j) {   /* nucl.py:544-560 */
    if (init == j && fin[i] != 0) return (double)fin[i];
    if (i == j && init == j) return -1.0;
    return 0.0;
}

/* |dT/dt|(T): log10-log10 interpolation on the cosmo table, binary search (input.py:177-226, :237) */
static double abs_dTdt(const oracle_t* o, double T) {
    const double v = log10(T);
    const double* x = o->cosmo_lT;
    int lo = 0, hi = o->n_cosmo - 2;   /* descending x: find ix with x[ix] >= v > x[ix+1] */
    while (lo < hi) {
        const int mid = (lo + hi) / 2;
        if (x[mid + 1] < v) hi = mid; else lo = mid + 1;
    }
    const int ix = lo;
    const double m = (o->cosmo_ld[ix + 1] - o->cosmo_ld[ix]) / (x[ix + 1] - x[ix]);
    const double b = o->cosmo_ld[ix] - m * x[ix];
    return pow(10.0, m * v + b);
}

typedef struct { oracle_t* o; int NT; const double* xlog; const double* ylog; /* [17][NT] log10 */ int i, j, dcy; } mat_ctx_t;
static double mat_integrand(double y, void* v) {   /* Ik_pdi / Ik_dcy, nucl.py:563-580, 609-610 */
    mat_ctx_t* c = (mat_ctx_t*)v;
    const double T = exp(y);
    double mat = 0.0;
    if (!c->dcy) {
        const double tl = log(T) / log(10.0);
        for (int r = 0; r < NR; ++r) {
            const double p = pref_ij(R_INIT[r], R_FIN[r], c->i, c->j);
            if (p != 0.0) {
                loginterp_t li = {c->NT, c->xlog, c->ylog + (size_t)r * c->NT};
                mat += p * pow(10.0, loginterp_log(&li, tl));
            }
        }
    } else {
        for (int d = 0; d < 2; ++d) mat += pref_ij(D_INIT[d], D_FIN[d], c->i, c->j) * HBAR / D_TAU[d];
    }
    return mat / (-abs_dTdt(c->o, T)) * T;
}

/* get_matp(T_lo): mpdi[9][9], mdcy[9][9] (nucl.py:583-623) */
void oracle_matp(oracle_t* o, int NT, const double* Tr, const double* pdi /* [NT][17] */, double T_lo, double* mpdi, double* mdcy) {
    double* xl = (double*)malloc(sizeof(double) * (NT + (size_t)NR * NT));
    double* yl = xl + NT;
    for (int k = 0; k < NT; ++k) {
        xl[k] = log(Tr[k]) / log(10.0);
        for (int r = 0; r < NR; ++r) yl[(size_t)r * NT + k] = log(pdi[(size_t)k * NR + r]) / log(10.0);
    }
    const double Tmax = Tr[NT - 1];
    for (int i = 0; i < NY; ++i)
        for (int j = 0; j < NY; ++j) {
            mat_ctx_t c = {o, NT, xl, yl, i, j, 0};
            int any = 0;
            for (int r = 0; r < NR; ++r) any |= pref_ij(R_INIT[r], R_FIN[r], i, j) != 0.0;
            mpdi[i * NY + j] = any ? quad(o, mat_integrand, &c, log(Tmax), log(T_lo), o->eps, o->eps < 1e-6 ? 400 : 100, NULL) : 0.0;
            c.dcy = 1;
            any = 0;
            for (int d = 0; d < 2; ++d) any |= pref_ij(D_INIT[d], D_FIN[d], i, j) != 0.0;
            mdcy[i * NY + j] = any ? quad(o, mat_integrand, &c, log(Tmax), log(T_lo), o->eps, o->eps < 1e-6 ? 400 : 100, NULL) : 0.0;
        }
    free(xl);
}

/* ------------------------------------------------------------------------------------------------------------
 * expm (scipy.linalg.expm: Pade-13 scaling & squaring, Higham 2005) and the final abundances (models.py:77-157)
 * ---------------------------------------------------------------------------------------------------------- */
static void mm(const double* A, const double* B, double* C) {
    for (int i = 0; i < NY; ++i)
        for (int j = 0; j < NY; ++j) {
            double s = 0.0;
            for (int k = 0; k < NY; ++k) s += A[i * NY + k] * B[k * NY + j];
            C[i * NY + j] = s;
        }
}

void oracle_expm(const double* Ain, double* R) {
    static const double b[14] = {64764752532480000., 32382376266240000., 7771770303897600., 1187353796428800.,
                                 129060195264000., 10559470521600., 670442572800., 33522128640., 1323241920., 40840800.,
                                 960960., 16380., 182., 1.};
    double A[81], A2[81], A4[81], A6[81], U[81], V[81], W[81], Z[81];
    double nrm = 0.0;
    for (int j = 0; j < NY; ++j) {
        double c = 0.0;
        for (int i = 0; i < NY; ++i) c += fabs(Ain[i * NY + j]);
        if (c > nrm) nrm = c;
    }
    int s = 0;
    if (nrm > 5.371920351148152) s = (int)ceil(log2(nrm / 5.371920351148152));
    if (s < 0) s = 0;
    for (int i = 0; i < 81; ++i) A[i] = ldexp(Ain[i], -s);
    mm(A, A, A2); mm(A2, A2, A4); mm(A4, A2, A6);
    for (int i = 0; i < 81; ++i) W[i] = b[13] * A6[i] + b[11] * A4[i] + b[9] * A2[i];
    mm(A6, W, Z);
    for (int i = 0; i < 81; ++i) Z[i] += b[7] * A6[i] + b[5] * A4[i] + b[3] * A2[i] + (i % 10 == 0 ? b[1] : 0.0);
    mm(A, Z, U);
    for (int i = 0; i < 81; ++i) W[i] = b[12] * A6[i] + b[10] * A4[i] + b[8] * A2[i];
    mm(A6, W, V);
    for (int i = 0; i < 81; ++i) V[i] += b[6] * A6[i] + b[4] * A4[i] + b[2] * A2[i] + (i % 10 == 0 ? b[0] : 0.0);
    for (int i = 0; i < 81; ++i) { W[i] = V[i] - U[i]; R[i] = V[i] + U[i]; }
    for (int c = 0; c < NY; ++c) {   /* LU with partial pivoting, applied to all right-hand sides */
        int piv = c;
        for (int r = c + 1; r < NY; ++r) if (fabs(W[r * NY + c]) > fabs(W[piv * NY + c])) piv = r;
        if (piv != c)
            for (int k = 0; k < NY; ++k) {
                double t = W[c * NY + k]; W[c * NY + k] = W[piv * NY + k]; W[piv * NY + k] = t;
                t = R[c * NY + k]; R[c * NY + k] = R[piv * NY + k]; R[piv * NY + k] = t;
            }
        for (int r = c + 1; r < NY; ++r) {
            const double f = W[r * NY + c] / W[c * NY + c];
            if (f == 0.0) continue;
            for (int k = c; k < NY; ++k) W[r * NY + k] -= f * W[c * NY + k];
            for (int k = 0; k < NY; ++k) R[r * NY + k] -= f * R[c * NY + k];
        }
    }
    for (int c = NY - 1; c >= 0; --c)
        for (int k = 0; k < NY; ++k) {
            double v = R[c * NY + k];
            for (int cc = c + 1; cc < NY; ++cc) v -= W[c * NY + cc] * R[cc * NY + k];
            R[c * NY + k] = v / W[c * NY + c];
        }
    for (int q = 0; q < s; ++q) { mm(R, R, W); memcpy(R, W, sizeof(W)); }
}

/* Yf[9][3] = postd . expm(mpdi+mdcy) . pred . Y0 (models.py:77-89, 133-157) */
void oracle_final_abundances(const oracle_t* o, const double* mpdi, const double* mdcy, double t_at_tmax, double* Yf) {
    double A[81], Ex[81], pred[81], postd[81], T1[81], T2[81];
    for (int i = 0; i < 81; ++i) { A[i] = mpdi[i] + mdcy[i]; pred[i] = postd[i] = (i % 10 == 0) ? 1.0 : 0.0; }
    oracle_expm(A, Ex);
    const double ef = exp(-t_at_tmax / TAU_T);
    pred[0] = 0.0; pred[1 * NY + 0] = 1.0; pred[3 * NY + 3] = ef; pred[4 * NY + 3] = 1.0 - ef;
    postd[0] = 0.0; postd[1 * NY + 0] = 1.0; postd[3 * NY + 3] = 0.0; postd[4 * NY + 3] = 1.0;
    postd[8 * NY + 8] = 0.0; postd[7 * NY + 8] = 1.0;
    mm(Ex, pred, T1); mm(postd, T1, T2);
    for (int c = 0; c < 3; ++c)
        for (int i = 0; i < NY; ++i) {
            double s = 0.0;
            for (int k = 0; k < NY; ++k) s += T2[i * NY + k] * o->Y0[k * 3 + c];
            Yf[i * 3 + c] = s;
        }
}

/* ------------------------------------------------------------------------------------------------------------
 * one full model point: get_pdi_grids + get_final_matp + expm + Yf (models.py:112-130, 43-91)
 *   SC: dense [NT][3][NE] or NULL;  outputs may be NULL except Yf
 * ---------------------------------------------------------------------------------------------------------- */
int oracle_run_point(oracle_t* o, int NE, const double* E_grid, int NT, const double* Tr, double E0, const double* S0 /* [NT][3] */,
                     const double* SC, double t_at_tmax, double* G_out /* [NT][3][NE] */, double* F_out /* [NT][3][NE] */,
                     double* pdi_out /* [NT][17] */, double* mpdi_out, double* mdcy_out, double* Yf) {
    double* G = (double*)malloc(sizeof(double) * (3 * NE + 5 * (size_t)NE * NE + 3 * NE + 3 * NE + (size_t)NT * NR));
    double* K = G + 3 * NE;
    double* F = K + 5 * (size_t)NE * NE;
    double* zero = F + 3 * NE;
    double* pdi = zero + 3 * NE;
    memset(zero, 0, sizeof(double) * 3 * NE);
    int warn = 0;
    for (int it = 0; it < NT; ++it) {
        const double T = Tr[it];
        oracle_rates_kernels(o, NE, E_grid, T, G, K);
        const double* sc = SC ? SC + (size_t)it * 3 * NE : zero;
        oracle_solve(o, NE, E_grid, G, K, S0 + 3 * it, sc, F);
        const double rate_E0 = total_rate_photon(o, E0, T);
        warn += oracle_pdi_rates(o, NE, E_grid, F, E0, T, S0[3 * it], rate_E0, pdi + (size_t)it * NR);
        if (G_out) memcpy(G_out + (size_t)it * 3 * NE, G, sizeof(double) * 3 * NE);
        if (F_out) memcpy(F_out + (size_t)it * 3 * NE, F, sizeof(double) * 3 * NE);
    }
    double mpdi[81], mdcy[81];
    oracle_matp(o, NT, Tr, pdi, Tr[0], mpdi, mdcy);
    oracle_final_abundances(o, mpdi, mdcy, t_at_tmax, Yf);
    if (pdi_out) memcpy(pdi_out, pdi, sizeof(double) * (size_t)NT * NR);
    if (mpdi_out) memcpy(mpdi_out, mpdi, sizeof(mpdi));
    if (mdcy_out) memcpy(mdcy_out, mdcy, sizeof(mdcy));
    free(G);
    return warn;
}

/* Many points on all host threads (the stand-in for scans.py's multiprocessing.Pool when this port is timed as the CPU
 * baseline).  Ragged inputs by offsets; SC separable form only (sc_T [S][3], sc_E [sumNE][3]) or none.
 * Returns the number of threads used. */
typedef struct {
    const oracle_t* proto;
    int n_points;
    const int *pt_ne, *pt_nt, *pt_sys_off;
    const int64_t* pt_e_off;
    const double *pt_e0, *pt_t_at_tmax, *e_grid, *sys_T, *s0, *sc_T, *sc_E;
    double* Yf;
    int next;            /* work-queue cursor (guarded by mu), chunk = 1 point like Pool.map_async(..., 1), scans.py:212 */
    long neval;
    pthread_mutex_t mu;
} scan_job_t;

static void* scan_worker(void* v) {
    scan_job_t* j = (scan_job_t*)v;
    long neval = 0;
    for (;;) {
        pthread_mutex_lock(&j->mu);
        const int p = j->next++;
        pthread_mutex_unlock(&j->mu);
        if (p >= j->n_points) break;
        oracle_t o = *j->proto;
        o.neval = 0;
        const int NE = j->pt_ne[p], NT = j->pt_nt[p];
        const double* E = j->e_grid + j->pt_e_off[p];
        double* SC = NULL;
        if (j->sc_T && j->sc_E) {
            SC = (double*)malloc(sizeof(double) * (size_t)NT * 3 * NE);
            for (int it = 0; it < NT; ++it)
                for (int X = 0; X < 3; ++X)
                    for (int i = 0; i < NE; ++i)
                        SC[((size_t)it * 3 + X) * NE + i] =
                            j->sc_T[(size_t)(j->pt_sys_off[p] + it) * 3 + X] * j->sc_E[(j->pt_e_off[p] + i) * 3 + X];
        }
        oracle_run_point(&o, NE, E, NT, j->sys_T + j->pt_sys_off[p], j->pt_e0[p], j->s0 + (size_t)j->pt_sys_off[p] * 3, SC,
                         j->pt_t_at_tmax[p], NULL, NULL, NULL, NULL, NULL, j->Yf + (size_t)p * 27);
        free(SC);
        neval += o.neval;
    }
    pthread_mutex_lock(&j->mu);
    j->neval += neval;
    pthread_mutex_unlock(&j->mu);
    return NULL;
}

int oracle_run_points(const oracle_t* proto, int n_points, const int* pt_ne, const int* pt_nt, const int64_t* pt_e_off,
                      const int* pt_sys_off, const double* pt_e0, const double* pt_t_at_tmax, const double* e_grid,
                      const double* sys_T, const double* s0, const double* sc_T, const double* sc_E, double* Yf,
                      int n_threads, long* neval_total) {
    if (n_threads <= 0) n_threads = (int)sysconf(_SC_NPROCESSORS_ONLN);
    if (n_threads > n_points) n_threads = n_points;
    if (n_threads < 1) n_threads = 1;
    scan_job_t j = {proto, n_points, pt_ne, pt_nt, pt_sys_off, pt_e_off, pt_e0, pt_t_at_tmax, e_grid, sys_T, s0, sc_T, sc_E,
                    Yf, 0, 0, PTHREAD_MUTEX_INITIALIZER};
    pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * n_threads);
    for (int t = 0; t < n_threads; ++t) pthread_create(&th[t], NULL, scan_worker, &j);
    for (int t = 0; t < n_threads; ++t) pthread_join(th[t], NULL);
    free(th);
    if (neval_total) *neval_total = j.neval;
    return n_threads;
}

/* ---- small accessors for the tests ---- */
double oracle_F(double Eph, double Ee, double Ephb) { return fn_F(Eph, Ee, Ephb); }
double oracle_G(double Ee, double Eph, double Ephb) { return fn_G(Ee, Eph, Ephb); }
double oracle_dsdE_Z2(double Ee, double Eph) { return fn_dsdE_Z2(Ee, Eph); }
double oracle_total_rate(oracle_t* o, int X, double E, double T) { return X == 0 ? total_rate_photon(o, E, T) : total_rate_electron(o, E, T); }
double oracle_rate_piece(oracle_t* o, int which, double E, double T) {
    switch (which) {
        case 0: return rate_photon_photon(E, T);
        case 1: return rate_compton(o, E, T);
        case 2: return rate_bethe_heitler(o, E, T);
        case 3: return rate_pair_creation_ae_db(o, E, T);
        case 4: return total_rate_electron(o, E, T);
    }
    return 0.0;
}
double oracle_kernel_piece(oracle_t* o, int which, double E, double Ep, double T) {
    switch (which) {
        case 0: return kernel_photon_photon(E, Ep, T);
        case 1: return kernel_compton_formula(o, E, Ep, T);
        case 2: return kernel_ic_photon(o, E, Ep, T);
        case 3: return kernel_ic_electron(o, E, Ep, T);
        case 4: return kernel_compton_formula(o, Ep + ME - E, Ep, T);
        case 5: return kernel_bethe_heitler(o, E, Ep, T);
        case 6: return kernel_pair_creation_ae(o, E, Ep, T);
    }
    return 0.0;
}
double oracle_abs_dTdt(const oracle_t* o, double T) { return abs_dTdt(o, T); }
int oracle_sizeof(void) { return (int)sizeof(oracle_t); }
/* quad on a test integrand family, to pin the GK21 port against scipy: f(x) = exp(-a x) * x^p */
typedef struct { double a, p; } tf_ctx_t;
static double tf(double x, void* v) { tf_ctx_t* c = (tf_ctx_t*)v; return exp(-c->a * x) * pow(x, c->p); }
double oracle_quad_test(oracle_t* o, double a, double p, double lo, double hi, double epsrel, int limit) {
    tf_ctx_t c = {a, p};
    return quad(o, tf, &c, lo, hi, epsrel, limit, NULL);
}
