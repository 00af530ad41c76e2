// acro_physics.cuh -- FP64 device functions of the electromagnetic-cascade and photodisintegration physics.
//
// Every function states the reference location (hep-mh/acropolis v1.3.1) whose arithmetic it restates.  The
// formulas are the published ones (Kawasaki & Moroi astro-ph/9412055, Jones Phys.Rev.167.1159, Poulin & Serpico
// 1503.04852, Cyburt et al. astro-ph/0211258, Hufnagel et al. 1808.09324 / ACROPOLIS manual); nothing here is a
// translation of the reference's Python control flow -- the batched kernels in acro_kernels.cu own that.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

namespace acro {

// ---- physical / mathematical constants (reference params.py:15-49) ------------------------------------------
constexpr double kAlpha = 1.0 / 137.036;
constexpr double kMe    = 0.511;
constexpr double kMe2   = 0.511 * 0.511;
constexpr double kRe    = kAlpha / kMe;
constexpr double kPi    = 3.14159265358979323846;
constexpr double kPi2   = kPi * kPi;
constexpr double kZeta3 = 1.2020569031595942;
constexpr double kHbar  = 6.582119514e-22;
constexpr double kTauN  = 8.794e2;
constexpr double kTauT  = 5.605e8;

// first entry of row i in a packed upper-triangular NE x NE plane (row-major, j >= i)
__host__ __device__ __forceinline__ long long tri_row_off(int i, int ne) {
    return (long long)i * ne - ((long long)i * (i - 1)) / 2;
}

struct Densities {
    double ne;     // electron density          (cascade.py:264-266)
    double nNZ2;   // sum_N Z^2 n_N             (cascade.py:269-271)
};

// n_b = eta * 2 zeta(3)/pi^2 * T^3 (cascade.py:258-261)
__host__ __device__ inline Densities densities(double T, double eta, double Yp, double YHe4) {
    const double nb = eta * (2.0 * kZeta3 / kPi2) * (T * T * T);
    Densities d;
    d.ne   = (Yp + 2.0 * YHe4) * nb;
    d.nNZ2 = (Yp + 4.0 * YHe4) * nb;
    return d;
}

// ---- inverse-Compton function F(E_gamma, E_e, eps_bar) (cascade.py:29-43; Jones 1968) -------------------------
__device__ __forceinline__ double ic_F(double Eph, double Ee, double x) {
    const double fxe = 4.0 * x * Ee;
    // validity window  x <= Eph <= 4 x Ee^2 / (me^2 + 4 x Ee), written without the division
    if (!(x <= Eph && Eph * (kMe2 + fxe) <= fxe * Ee)) return 0.0;
    const double Gm = fxe * (1.0 / kMe2);
    const double q  = Eph / (Gm * (Ee - Eph));
    const double Gq = Gm * q;
    return 2.0 * q * log(q) + (1.0 + 2.0 * q) * (1.0 - q) + Gq * Gq * (1.0 - q) / (2.0 + 2.0 * Gq);
}

// ---- pair-creation spectrum function G(E_e, E_gamma, eps_bar) (cascade.py:47-80; Aharonian et al. 1983) -------
__device__ __forceinline__ double pc_G(double Ee, double Eph, double x) {
    const double s   = Eph + x;          // Ee + Eep
    const double Eep = s - Ee;
    const double dEs = (Eph - x) * sqrt(1.0 - kMe2 / (Eph * x));
    const double lim_m = 0.5 * (s - dEs), lim_p = 0.5 * (s + dEs);
    // The reference also requires me < lim_m (cascade.py:63).  lim_m has its minimum at x ~ me/2 where it EQUALS me up to
    // rounding (electron at rest), so that test fails, by rounding, on a set of measure zero inside the integration
    // range; a fixed-order rule whose node lands there would lose the node (measured: 35 of 1e8 entries off by 1-9 %),
    // whereas the reference's adaptive rule bisects around it.  The test is therefore applied with a rounding margin.
    if (!(kMe * (1.0 - 1e-9) < lim_m && lim_m <= Ee && Ee <= lim_p)) return 0.0;
    const double ee  = Ee * Eep;
    const double s2  = s * s;
    const double iee = 1.0 / ee;
    double sud = 4.0 * s2 * log(4.0 * x * ee / (kMe2 * s)) * iee;
    sud += (kMe2 / (x * s) - 1.0) * (s2 * s2) * (iee * iee);
    sud += 2.0 * (2.0 * x * s - kMe2) * s2 * iee * (1.0 / kMe2);
    sud += -8.0 * x * s * (1.0 / kMe2);
    return sud;
}

// ---- Bethe-Heitler differential cross-section / Z^2 (cascade.py:139-164) --------------------------------------
__device__ inline double bh_dsdE_Z2(double Ee, double Eph) {
    const double Em = Ee, Ep = Eph - Ee;
    const double pm = sqrt(Em * Em - kMe2), pp = sqrt(Ep * Ep - kMe2);
    const double EE = Ep * Em, PP = pp * pm;
    const double L  = log((EE + PP + kMe2) / (EE - PP + kMe2));
    const double lm = log((Em + pm) / (Em - pm));
    const double lp = log((Ep + pp) / (Ep - pp));
    const double pref = kAlpha * (kRe * kRe) * PP / (Eph * Eph * Eph);
    const double pp2 = pp * pp, pm2 = pm * pm, pp3 = pp2 * pp, pm3 = pm2 * pm;
    double sud = -4.0 / 3.0 - 2.0 * EE * (pp2 + pm2) / (pp2 * pm2);
    sud += kMe2 * (lm * Ep / pm3 + lp * Em / pp3 - lp * lm / PP);
    sud += L * (-8.0 * EE / (3.0 * PP) + Eph * Eph * (EE * EE + PP * PP - kMe2 * EE) / (pp3 * pm3));
    sud += -L * kMe2 * Eph * (lp * (EE - pp2) / pp3 + lm * (EE - pm2) / pm3) / (2.0 * PP);
    return pref * sud;
}

// ---- closed-form photon rates (cascade.py:285-332) -------------------------------------------------------------
__device__ inline double rate_photon_photon(double E, double T) {
    const double a2 = kAlpha * kAlpha;
    const double e = E / kMe, t = T / kMe, t3 = t * t * t;
    return 0.151348 * (a2 * a2) * kMe * (e * e * e) * (t3 * t3) * exp(-E * T / kMe2);
}

__device__ inline double rate_compton(double E, double ne) {
    const double x = 2.0 * E / kMe;
    const double opx = 1.0 + x;
    return (2.0 * kPi * (kRe * kRe) / x) * ne *
           ((1.0 - 4.0 / x - 8.0 / (x * x)) * log(opx) + 0.5 + 8.0 / x - 1.0 / (2.0 * opx * opx));
}

__device__ inline double rate_bethe_heitler(double E, double nNZ2) {
    const double k = E / kMe;
    if (k < 2.0) return 0.0;
    const double a3me2 = kAlpha * kAlpha * kAlpha / kMe2;
    if (k <= 4.0) {   // small-energy fit (cascade.py:315-320)
        const double r  = (2.0 * k - 4.0) / (k + 2.0 + 2.0 * sqrt(2.0 * k));
        const double f  = (k - 2.0) / k;
        const double r2 = r * r;
        return a3me2 * nNZ2 * (2.0 * kPi / 3.0) * (f * f * f) *
               (1.0 + r / 2.0 + (23.0 / 40.0) * r2 + (11.0 / 60.0) * (r2 * r) + (29.0 / 960.0) * (r2 * r2));
    }
    const double l = log(2.0 * k);   // large-energy expansion to (2/k)^6 (cascade.py:324-332)
    const double u = 2.0 / k, u2 = u * u, u4 = u2 * u2, u6 = u4 * u2;
    return a3me2 * nNZ2 *
           ((28.0 / 9.0) * l - 218.0 / 27.0 +
            u2 * ((2.0 / 3.0) * l * l * l - l * l + (6.0 - kPi2 / 3.0) * l + 2.0 * kZeta3 + kPi2 / 6.0 - 7.0 / 2.0) -
            u4 * ((3.0 / 16.0) * l + 1.0 / 8.0) - u6 * ((29.0 / 2304.0) * l - 77.0 / 13824.0));
}

// ---- closed-form kernels (cascade.py:383-402, 621-639, 550-558) --------------------------------------------------
__device__ inline double kernel_photon_photon(double E, double Ep, double T) {
    const double a2 = kAlpha * kAlpha, m2 = kMe2, m8 = (m2 * m2) * (m2 * m2);
    const double T3 = T * T * T, r = E / Ep, g = 1.0 - r + r * r;
    const double pi4 = kPi2 * kPi2;
    return 1112.0 / (10125.0 * kPi) * (a2 * a2) / m8 * 8.0 * pi4 * (T3 * T3) / 63.0 * (Ep * Ep) * (g * g) *
           exp(-Ep * T / kMe2);
}

// Klein-Nishina differential kernel; E = outgoing energy of the tracked particle after substitution
__device__ inline double kernel_compton_core(double E, double Ep, double ne) {
    if (Ep / (1.0 + 2.0 * Ep / kMe) > E) return 0.0;   // Compton edge
    const double d = kMe / E - kMe / Ep;
    return kPi * (kRe * kRe) * kMe / (Ep * Ep) * ne * (Ep / E + E / Ep + d * d - 2.0 * kMe * (1.0 / E - 1.0 / Ep));
}

// ---- bilinear log-log interpolation of rates.db (db.py:66-124), same expanded a0..a3 form, no FMA contraction ----
struct RateDb {
    const double* tab;   // [nT][nE][2]
    int nT, nE;
    double Elog_min, Elog_max, Tlog_min, Tlog_max;
};

__device__ inline bool in_rate_db(const RateDb& db, double E_log, double T_log) {
    return (db.Elog_min <= E_log && E_log <= db.Elog_max) && (db.Tlog_min <= T_log && T_log <= db.Tlog_max);
}

__device__ inline double interp_rate_db(const RateDb& db, int c, double E_log, double T_log) {
    int iE = (int)((db.nE - 1) * (E_log - db.Elog_min) / (db.Elog_max - db.Elog_min));
    int iT = (int)((db.nT - 1) * (T_log - db.Tlog_min) / (db.Tlog_max - db.Tlog_min));
    if (iE == db.nE - 1) iE -= 1;
    if (iT == db.nT - 1) iT -= 1;
    const double dE = db.Elog_max - db.Elog_min, dT = db.Tlog_max - db.Tlog_min;
    const double x = T_log, y = E_log;
    const double x0 = db.Tlog_min + __ddiv_rn(__dmul_rn(dT, (double)iT), (double)(db.nT - 1));
    const double x1 = db.Tlog_min + __ddiv_rn(__dmul_rn(dT, (double)(iT + 1)), (double)(db.nT - 1));
    const double y0 = db.Elog_min + __ddiv_rn(__dmul_rn(dE, (double)iE), (double)(db.nE - 1));
    const double y1 = db.Elog_min + __ddiv_rn(__dmul_rn(dE, (double)(iE + 1)), (double)(db.nE - 1));
    const double c00 = db.tab[((size_t)iT * db.nE + iE) * 2 + c];
    const double c10 = db.tab[((size_t)(iT + 1) * db.nE + iE) * 2 + c];
    const double c01 = db.tab[((size_t)iT * db.nE + iE + 1) * 2 + c];
    const double c11 = db.tab[((size_t)(iT + 1) * db.nE + iE + 1) * 2 + c];
#define M(a, b) __dmul_rn((a), (b))
#define A(a, b) __dadd_rn((a), (b))
#define S(a, b) __dsub_rn((a), (b))
    const double d  = M(S(x0, x1), S(y0, y1));
    const double a0 = __ddiv_rn(A(S(S(M(M(c00, x1), y1), M(M(c01, x1), y0)), M(M(c10, x0), y1)), M(M(c11, x0), y0)), d);
    const double a1 = __ddiv_rn(S(A(A(M(-c00, y1), M(c01, y0)), M(c10, y1)), M(c11, y0)), d);
    const double a2 = __ddiv_rn(S(A(A(M(-c00, x1), M(c01, x1)), M(c10, x0)), M(c11, x0)), d);
    const double a3 = __ddiv_rn(A(S(S(c00, c01), c10), c11), d);
    const double v  = A(A(A(a0, M(a1, x)), M(a2, y)), M(M(a3, x), y));
#undef M
#undef A
#undef S
    return pow(10.0, v);
}

// ---- photodisintegration cross-sections in MeV^-2 (nucl.py:70-88, 174-183, 209-395) --------------------------------
__constant__ const double kEth[17] = {2.224573, 6.257248, 8.481821, 5.493485, 7.718058, 19.813852, 20.577615, 23.846527,
                                      26.071100, 3.698892, 15.794685, 2.467032, 7.249962, 10.948850, 1.586627, 5.605794,
                                      9.304680};
constexpr double kMbToIMeV2 = 2.56819e-6;
// E at which the closing polynomial of reaction 15 (nucl.py:328) changes sign: Q15 + 3.8259903470277856
constexpr double kKink15 = 5.412617347027785;

// N Q^p1 (E-Q)^p2 / E^p3 above threshold (nucl.py:209-214); lnE = log(E) is supplied by the caller
__device__ __forceinline__ double sig_generic(double E, double lnE, double Q, double N, double p1, double p2, double p3) {
    if (E < Q) return 0.0;
    if (E == Q) return (p2 > 0.0) ? 0.0 : N * exp(p1 * log(Q) - p3 * lnE);
    return N * exp(p1 * log(Q) + p2 * log(E - Q) - p3 * lnE);
}

__device__ inline double cross_section_mb(int rid, double E, double lnE) {   // rid = 1..17, lnE = log(E)
    const double Q = kEth[rid - 1];
    if (E < Q) return 0.0;
    switch (rid) {
        case 1: {   // d+a>n+p (nucl.py:217-224)
            const double r = sqrt(Q * (E - Q)) / E;
            const double sq = sqrt(Q) - sqrt(0.037);
            return 18.75 * (r * r * r + 0.007947 * (r * r) * (sq * sq / (E - Q + 0.037)));
        }
        case 2:  return sig_generic(E, lnE, Q, 9.8, 1.95, 1.65, 3.6);
        case 3:  return sig_generic(E, lnE, Q, 26.0, 2.6, 2.3, 4.9);
        case 4:  return sig_generic(E, lnE, Q, 8.88, 1.75, 1.65, 3.4);
        case 5:  return sig_generic(E, lnE, Q, 16.7, 1.95, 2.3, 4.25);
        case 6:  return sig_generic(E, lnE, Q, 19.5, 3.5, 1.0, 4.5);
        case 7:  return sig_generic(E, lnE, Q, 20.7, 3.5, 1.0, 4.5);
        case 8:  return sig_generic(E, lnE, Q, 10.7, 10.2, 3.4, 13.6);
        case 9:  return sig_generic(E, lnE, Q, 21.7, 4.0, 3.0, 7.0);
        case 10: return sig_generic(E, lnE, Q, 143.0, 2.3, 4.7, 7.0);
        case 11: {  // Li6+a>X: three Gaussians on a power law (nucl.py:267-277)
            const double g1 = (E - 19.0) / 3.5, g2 = (E - 30.0) / 3.0, g3 = (E - 43.0) / 5.0;
            return sig_generic(E, lnE, Q, 38.1, 3.0, 2.0, 5.0) *
                   (3.7 * exp(-0.5 * g1 * g1) + 2.75 * exp(-0.5 * g2 * g2) + 2.2 * exp(-0.5 * g3 * g3));
        }
        case 12: {  // Li7+a>t+He4 (nucl.py:280-300)
            const double Ecm = E - Q;
            if (Ecm <= 0.0) return 0.0;
            const double e2 = Ecm * Ecm;
            const double p = 1.0 + 2.2875 * e2 - 1.1798 * (e2 * Ecm) + 2.5279 * (e2 * e2);
            return 0.105 * (2371.0 / (E * E)) * exp(-2.5954 / sqrt(Ecm) - 2.056 * Ecm) * p;
        }
        case 13: {  // Li7+a>n+Li6: two power laws + Lorentzian (nucl.py:303-310)
            const double z = (E - Q - 7.46) / 0.188;
            return sig_generic(E, lnE, Q, 0.176, 1.51, 0.49, 2.0) + sig_generic(E, lnE, Q, 1205.0, 5.5, 5.0, 10.5) +
                   0.06 / (1.0 + z * z);
        }
        case 14: return sig_generic(E, lnE, Q, 122.0, 4.0, 3.0, 7.0);
        case 15: {  // Be7+a>He3+He4 with the polynomial sign guard (nucl.py:317-337)
            const double Ecm = E - Q;
            if (Ecm <= 0.0) return 0.0;
            const double e2 = Ecm * Ecm;
            const double p = 1.0 - 0.428 * e2 + 0.534 * (e2 * Ecm) - 0.115 * (e2 * e2);
            if (p < 0.0) return 0.0;
            return 0.504 * (2371.0 / (E * E)) * exp(-5.1909 / sqrt(Ecm) - 0.548 * Ecm) * p;
        }
        case 16: return sig_generic(E, lnE, Q, 32.6, 10.0, 2.0, 12.0) + sig_generic(E, lnE, Q, 2.27e6, 8.8335, 13.0, 21.8335);
        case 17: return sig_generic(E, lnE, Q, 133.0, 4.0, 3.0, 7.0);
    }
    return 0.0;
}

__device__ inline double cross_section(int rid, double E) { return kMbToIMeV2 * cross_section_mb(rid, E, log(E)); }

}  // namespace acro
