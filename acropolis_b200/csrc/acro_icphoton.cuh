// acro_icphoton.cuh -- the e -> gamma inverse-Compton kernel plane (reference cascade.py:406-435) in closed form.
//
// With q = x_a/x the factor F(E, E', x) of cascade.py:29-43 is  2 q ln q + (1+2q)(1-q) + c9 (1-q),  c9 = r^2/(2+2r),
// r = E/(E'-E): x enters only through q.  In v = x/T the thermal integral therefore separates into one-variable functions,
//
//     int_{x_a}^{x_b} F x / (pi^2 (e^{x/T} - 1)) dx  =  (T/pi)^2 [ Phi0(v_a, v_b) + c9 Phi1(v_a, v_b) ]
//     Phi1 = B(v_a) - [B(v_b) + (v_b - v_a) L(v_b)]
//     Phi0 = A(v_a) - [B(v_b) + v_b L(v_b) + v_a (1 + 2 ln v_a) L(v_b) - 2 v_a^2 M(v_b) - 2 v_a N(v_b)]
//     L(v) = -ln(1 - e^-v),  B(v) = Li2(e^-v),  M(v) = int_v^inf dt/(t (e^t - 1)),  N(v) = int_v^inf ln t/(e^t - 1) dt,
//     A(a) = int_a^inf [2 q ln q + (1+2q)(1-q)]_{q = a/v} v/(e^v - 1) dv,
//
// tabulated (times e^v, piecewise polynomials of degree 11 in ln v; tools/make_icphoton_tables.py, 30-digit mpmath) so that one
// (pair, T) entry costs two polynomial evaluations and one exponential instead of a quadrature -- and is exact to 3e-14 (the
// generator's check against 30-digit quadrature of the original integral) where the quadratures were good to 1e-9.  The upper-limit
// bracket is dropped when v_b - v_a >= 36 (it is below 4e-14 of the result); where the range ends at the thermal cut-off
// v_b = Ephb_T_max (the common case of a bracket that matters: v_a > 164) its five function values are per-block constants.
// Two corners keep a quadrature (thermal_integral): ranges shorter than 1/4 in v, where the brackets cancel (8-node rule), and
// upper limits v_b < e^2 that still matter (T > E/7.4: hotter than anything photodisintegration is computed for), where M and
// N are not tabulated.
//
// This kernel writes plane 1 completely (zeros where the range is empty); kernels_mma_kernel / kernels_wien_kernel run for the
// e -> e and gamma -> e integrals only.
#pragma once
#include "acro_tshared.cuh"

#define ACRO_ICP_TABLE __device__ const
#include "acro_icphoton_tab.h"

namespace acro {

constexpr int kIcpThreads = 256;
constexpr int kIcpChunk = 64;                 // temperatures staged per pass
constexpr int kIcpRow = ACRO_ICP_DEG + 1;
constexpr double kIcpTailFrom = 36.0;         // v_b - v_a from which the upper-limit bracket is dropped
constexpr double kIcpShort = 0.25;            // v_b - v_a below which the brackets cancel (> 1e3 x rounding): quadrature
constexpr double kIcpVTailMin = 7.38905609893065;   // e^ACRO_ICP_U_TAIL: M, N are tabulated from here
constexpr double kIcpVMax = 244.0;                  // < e^ACRO_ICP_U_MAX: end of all tables (only reachable with Ephb_T_max > 244)

__device__ __forceinline__ double icp_poly(const double* __restrict__ row, double x) {
    double p = row[ACRO_ICP_DEG];
#pragma unroll
    for (int j = ACRO_ICP_DEG - 1; j >= 0; --j) p = fma(p, x, row[j]);
    return p;
}

// interval of the A/B tables that contains u = ln v, and the local variable in [-1, 1] (selects, no branch)
__device__ __forceinline__ int icp_locate(double u, double& x) {
    constexpr int n_lo = (int)((ACRO_ICP_U_SPLIT - ACRO_ICP_U_MIN) / ACRO_ICP_W_LO);
    const bool hi = u >= ACRO_ICP_U_SPLIT;
    const double t = (fmax(u, ACRO_ICP_U_MIN) - (hi ? ACRO_ICP_U_SPLIT : ACRO_ICP_U_MIN)) * (hi ? 1.0 / ACRO_ICP_W_HI : 1.0 / ACRO_ICP_W_LO);
    const int k = min((int)t, hi ? ACRO_ICP_N_MAIN - n_lo - 1 : n_lo - 1);     // below e^-40 the functions are constant
    x = fma(2.0, t - (double)k, -1.0);
    return (hi ? n_lo : 0) + k;
}

struct IcpPair {     // per-pair constants of ic_photon_closed_kernel
    double xa, xb0, lnxa, c9, pre;
};

// The rare (pair, T) combinations: short ranges (quadrature), upper-limit brackets at a kinematic limit (general table
// look-ups) and upper limits outside the M, N tables (quadrature).  Not inlined: keeps the main loop small.
__device__ __noinline__ double icp_slow(const IcpPair& q, double T, double v_cut, bool cut_tab, const double* __restrict__ tA,
                                        const double* __restrict__ tB, const double* __restrict__ tM, const double* __restrict__ tN,
                                        const double2* __restrict__ lt) {
    const double x_cut = v_cut * T;
    const bool capped = x_cut < q.xb0;
    const double xb = capped ? x_cut : q.xb0;
    const double iT = frcp(T);
    const double va = q.xa * iT, vb = capped ? v_cut : xb * iT, S = vb - va;
    if (S < kIcpShort || va > kIcpVMax || vb > kIcpVMax || (S < kIcpTailFrom && (capped ? !cut_tab : vb < kIcpVTailMin))) {
        PairConst c;
        c.c9 = q.c9;
#if ACRO_FQ_LOGTAB
        c.lt = lt;
#endif
        // short ranges under the cut-off sit at v_a > 199: the 8-node rule of the Wien kernel; anything else: generic
        return q.pre * (capped && va > 54.598150033144236 ? thermal_integral<KIND_IC_PHOTON, 64, true>(c, T, q.xa, xb, q.xb0)
                                                          : thermal_integral<KIND_IC_PHOTON, 64>(c, T, q.xa, xb));
    }
    const double ua = q.lnxa - log(T);
    double x;
    const int k = icp_locate(ua, x);
    const double ea = fexp(-va);
    double phi0 = icp_poly(tA + k * kIcpRow, x) * ea;
    double phi1 = icp_poly(tB + k * kIcpRow, x) * ea;
    const double ub = flog(vb);
    const double eb = fexp(-vb);
    double xq;
    const int kb = icp_locate(ub, xq);
    const double Bb = icp_poly(tB + kb * kIcpRow, xq) * eb;
    const double tt = (ub - ACRO_ICP_U_TAIL) * (1.0 / ACRO_ICP_W_HI);
    const int kt = min((int)tt, ACRO_ICP_N_TAIL - 1);
    const double xt = fma(2.0, tt - (double)kt, -1.0);
    const double Mb = icp_poly(tM + kt * kIcpRow, xt) * eb;
    const double Nb = icp_poly(tN + kt * kIcpRow, xt) * eb;
    // L(v_b) = -log1p(-e^-v_b), e^-v_b <= e^-e^2 = 6.2e-4: five terms (the sixth is < 2e-17 of the sum)
    const double Lb = eb * fma(eb, fma(eb, fma(eb, fma(eb, 0.2, 0.25), 1.0 / 3.0), 0.5), 1.0);
    phi0 -= Bb + vb * Lb + va * (fma(2.0, ua, 1.0) * Lb - 2.0 * fma(va, Mb, Nb));
    phi1 -= fma(S, Lb, Bb);
    return q.pre * (T * T * (1.0 / kPi2)) * fma(q.c9, phi1, phi0);
}

#ifndef ACRO_ICP_MINBLOCKS
#define ACRO_ICP_MINBLOCKS 3      // measured (ms per 1000-point slice): 1: 6.0, 2: 4.65, 3: 4.45, 4: 5.0, 6: 10.3 (spills)
#endif
__global__ void __launch_bounds__(kIcpThreads, ACRO_ICP_MINBLOCKS) ic_photon_closed_kernel(acro_batch_t b, double* __restrict__ Kws, acro_params_t p) {
    __shared__ double sA[ACRO_ICP_N_MAIN * kIcpRow], sB[ACRO_ICP_N_MAIN * kIcpRow];
    __shared__ double sM[ACRO_ICP_N_TAIL * kIcpRow], sN[ACRO_ICP_N_TAIL * kIcpRow];
    __shared__ double sT[kIcpChunk], sIT[kIcpChunk], sLnT[kIcpChunk], sPre[kIcpChunk];
    __shared__ double sCut[4];                  // at v = Ephb_T_max: B + v L, L, M, N
    __shared__ int64_t sK[kIcpChunk];
#if ACRO_FQ_LOGTAB
    __shared__ double2 s_lt[kLogTab];
    for (int k = threadIdx.x; k < kLogTab; k += blockDim.x) s_lt[k] = g_logtab[k];
#endif
    const int pt = blockIdx.y;
    const int ne = b.pt_ne[pt], nt = b.pt_nt[pt], s0 = b.pt_sys_off[pt];
    const int64_t tri = (int64_t)ne * (ne + 1) / 2;
    if ((int64_t)blockIdx.x * blockDim.x >= tri) return;
    for (int k = threadIdx.x; k < ACRO_ICP_N_MAIN * kIcpRow; k += blockDim.x) {
        sA[k] = kIcpA[k / kIcpRow][k % kIcpRow];
        sB[k] = kIcpB[k / kIcpRow][k % kIcpRow];
    }
    for (int k = threadIdx.x; k < ACRO_ICP_N_TAIL * kIcpRow; k += blockDim.x) {
        sM[k] = kIcpM[k / kIcpRow][k % kIcpRow];
        sN[k] = kIcpN[k / kIcpRow][k % kIcpRow];
    }
    __syncthreads();
    const double v_cut = p.Ephb_T_max;
    const bool cut_tab = v_cut >= kIcpVTailMin && v_cut <= kIcpVMax;   // inside the M, N tables (e^2 .. e^5.5)
    if (threadIdx.x == 0 && cut_tab) {
        const double ub = log(v_cut), eb = exp(-v_cut);
        double xq;
        const int kb = icp_locate(ub, xq);
        const double tt = (ub - ACRO_ICP_U_TAIL) * (1.0 / ACRO_ICP_W_HI);
        const int kt = min((int)tt, ACRO_ICP_N_TAIL - 1);
        const double xt = fma(2.0, tt - (double)kt, -1.0);
        const double Lb = -log1p(-eb);
        sCut[0] = icp_poly(sB + kb * kIcpRow, xq) * eb + v_cut * Lb;
        sCut[1] = Lb;
        sCut[2] = icp_poly(sM + kt * kIcpRow, xt) * eb;
        sCut[3] = icp_poly(sN + kt * kIcpRow, xt) * eb;
    }
    const int64_t loc = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = loc < tri;
    const int64_t locc = valid ? loc : tri - 1;
    const double tn = 2.0 * ne + 1.0;
    int i = (int)floor(0.5 * (tn - sqrt(tn * tn - 8.0 * (double)locc)));
    i = max(0, min(i, ne - 1));
    while (i > 0 && tri_row_off(i, ne) > locc) --i;
    while (i < ne - 1 && tri_row_off(i + 1, ne) <= locc) ++i;
    const int j = i + (int)(locc - tri_row_off(i, ne));
    const double* eg = b.e_grid + b.pt_e_off[pt];
    const double E = eg[i], Ep = eg[j];
    // limits as in pair_limits() (cascade.py:415-427): nothing below the threshold E' >= E + me
    const bool has = !(Ep < E + kMe);
    IcpPair q;
    q.xa = has ? 0.25 * kMe2 * E / (Ep * Ep - Ep * E) : 1.0;
    q.xb0 = fmin(E, Ep - kMe2 / (4.0 * Ep));
    q.lnxa = log(q.xa);
    const double r = E / fmax(Ep - E, 1e-300);
    q.c9 = r * r / (2.0 + 2.0 * r);
    q.pre = 2.0 * kPi * (kAlpha * kAlpha) / (Ep * Ep);
    double* __restrict__ out = Kws + b.n_pairs_total + loc;        // plane 1
#if ACRO_FQ_LOGTAB
    const double2* lt = s_lt;
#else
    const double2* lt = nullptr;
#endif
    // One entry, branch-free (two of them are in flight per trip of the loop below): the closed form with the cut-off bracket
    // as a select.  Returns true when the combination needs icp_slow() instead; an empty range yields 0.
    auto entry = [&](int t, double& val) -> bool {
        const double T = sT[t], iT = sIT[t];
        const bool capped = v_cut * T < q.xb0;             // the range ends at the thermal cut-off
        const double xb = capped ? v_cut * T : q.xb0;
        const bool live = has && xb > q.xa;
        const double va = q.xa * iT, vb = capped ? v_cut : xb * iT, S = vb - va;
        const bool tail = S < kIcpTailFrom;
        const double ua = q.lnxa - sLnT[t];
        double x;
        const int k = icp_locate(ua, x) * kIcpRow;
        const double ea = fexp(-fmin(va, 700.0));
        double pa = sA[k + ACRO_ICP_DEG], pb = sB[k + ACRO_ICP_DEG];
#pragma unroll
        for (int j = ACRO_ICP_DEG - 1; j >= 0; --j) { pa = fma(pa, x, sA[k + j]); pb = fma(pb, x, sB[k + j]); }
        const double Lb = sCut[1];
        const double br0 = sCut[0] + va * (fma(2.0, ua, 1.0) * Lb - 2.0 * fma(va, sCut[2], sCut[3]));
        const double br1 = fma(-va, Lb, sCut[0]);
        const bool cut = tail && capped;
        const double phi0 = fma(pa, ea, cut ? -br0 : 0.0), phi1 = fma(pb, ea, cut ? -br1 : 0.0);
        val = live ? q.pre * sPre[t] * fma(q.c9, phi1, phi0) : 0.0;
        return live && (S < kIcpShort || va > kIcpVMax || (tail && !(capped && cut_tab)));
    };

    for (int t0 = 0; t0 < nt; t0 += kIcpChunk) {
        const int n = min(kIcpChunk, nt - t0);
        __syncthreads();
        if (threadIdx.x < n) {
            const int s = s0 + t0 + threadIdx.x;
            const double T = b.sys_T[s];
            sT[threadIdx.x] = T;
            sIT[threadIdx.x] = 1.0 / T;
            sLnT[threadIdx.x] = log(T);
            sPre[threadIdx.x] = T * T * (1.0 / kPi2);
            sK[threadIdx.x] = b.sys_k_off[s];
        }
        __syncthreads();
        if (!valid) continue;
        int t = 0;
#pragma unroll 1
        for (; t + 1 < n; t += 2) {
            double v0, v1;
            const bool s0f = entry(t, v0), s1f = entry(t + 1, v1);
            if (s0f) v0 = icp_slow(q, sT[t], v_cut, cut_tab, sA, sB, sM, sN, lt);
            if (s1f) v1 = icp_slow(q, sT[t + 1], v_cut, cut_tab, sA, sB, sM, sN, lt);
            out[sK[t]] = v0;
            out[sK[t + 1]] = v1;
        }
        if (t < n) {
            double v0;
            if (entry(t, v0)) v0 = icp_slow(q, sT[t], v_cut, cut_tab, sA, sB, sM, sN, lt);
            out[sK[t]] = v0;
        }
    }
}

}  // namespace acro
