// acro_kernels.cu -- hand-written sm_100a kernels of the ACROPOLIS hot path (FP64 throughout).
//
//   rates_kernel        thread per (system, E_i)        G[3,NE]: closed forms + rates.db bilinear interpolation
//   direct_rates_kernel warp per out-of-table (E,T)     the two 2-D rate integrals the reference does with dblquad
//   pair_closed_forms_kernel, ic_photon_closed_kernel, kernels_mma_kernel<KIND>, kernels_wien_kernel<Q,KIND>
//                                                       the production kernel builder (acro_icphoton.cuh, acro_tmma.cuh,
//                                                       acro_tshared.cuh): five K planes -- closed forms per pair; the
//                                                       e -> gamma integral from tabulated one-variable functions; the
//                                                       e -> e and gamma -> e integrals on a temperature-shared grid
//                                                       contracted with DMMA, Gauss-Laguerre on their Wien tail
//   kernels_kernel      thread per (system, E_i, E_j>=i) cross-check modes, only with -DACRO_CROSSCHECK (the test build):
//                                                       fixed-order Gauss-Legendre in log(eps_bar) per (E,E',T)
//                                                       (PLAIN_GL with the reference's F/G, PER_T reduced)
//   solve_pdi_kernel    warp per system                 back-substitution of the 3-species cascade equation,
//                                                       fused with the 17 sigma(E)*F(E) rate integrals
//   matrix_kernel       block (8 warps) per point       exact piecewise T-integral; 9x9 expm, decay matrices, Yf by warp 0
//   transfer_kernel     warp per matrix                 fast-parameter path: expm(f*mpdi+mdcy) and Yf only
//
// The sum over grid nodes of the shared-grid kernels is a dense FP64 contraction and runs on the FP64 tensor path
// (mma.sync m8n8k4 f64 -> DMMA; tcgen05 has no FP64 type); the bound is the FP64 pipe -- which DFMA and DMMA share on
// B200 -- for the quadrature kernels and HBM for the solve.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>
#include <algorithm>

#include "acro_internal.h"
#include "acro_physics.cuh"

namespace acro {

// ------------------------------------------------------------------------------------------------------------
// Gauss-Legendre tables in constant memory: orders 8, 16, 32, 64, 96, 128, 12, 10 packed back to back.
// ------------------------------------------------------------------------------------------------------------
// cross-section tables of sigma_rt (filled by upload_quadrature_tables)
__constant__ double c_sig[ACRO_NR][4];    // generic reactions: {log(N Q^p1), p2, p3, Q}
__constant__ double c_sig2[2][3];         // second power-law terms of reactions 13 and 16: {log(N Q^p1), p2, p3}

constexpr int kGLTotal = 8 + 16 + 32 + 64 + 96 + 128 + 12 + 10;
__constant__ double c_glx[kGLTotal];
__constant__ double c_glw[kGLTotal];

__host__ __device__ constexpr int gl_offset(int order) {
    return order == 8 ? 0 : order == 16 ? 8 : order == 32 ? 24 : order == 64 ? 56 : order == 96 ? 120 : order == 128 ? 216 : order == 12 ? 344 : 356;   // 356: order 10
}

}  // namespace acro
#include "acro_fastquad.cuh"
#include "acro_tshared.cuh"
#include "acro_tmma.cuh"
#include "acro_icphoton.cuh"
namespace acro {

static void gauss_laguerre(int n, double* x, double* w) {   // alpha = 0; Newton on L_n with the classic initial guesses
    std::vector<long double> xs(n);
    for (int i = 0; i < n; ++i) {
        long double z;
        if (i == 0) z = 3.0L / (1.0L + 2.4L * n);
        else if (i == 1) z = xs[0] + 15.0L / (1.0L + 2.5L * n);
        else { const long double ai = i - 1; z = xs[i - 1] + ((1.0L + 2.55L * ai) / (1.9L * ai)) * (xs[i - 1] - xs[i - 2]); }
        long double pp = 0, p2 = 0;
        for (int it = 0; it < 100; ++it) {
            long double p1 = 1, pm = 0;
            for (int j = 1; j <= n; ++j) { const long double p3 = pm; pm = p1; p1 = ((2 * j - 1 - z) * pm - (j - 1) * p3) / j; }
            p2 = pm;
            pp = n * (p1 - p2) / z;
            const long double z1 = z;
            z = z1 - p1 / pp;
            if (fabsl(z - z1) <= 1e-18L * fabsl(z)) break;
        }
        xs[i] = z;
        x[i] = (double)z;
        w[i] = (double)(-1.0L / (pp * n * p2));
    }
}

static void gauss_legendre(int n, double* x, double* w) {
    // Newton iteration on P_n (long double), roots symmetric
    for (int i = 0; i < (n + 1) / 2; ++i) {
        long double z = cosl(3.14159265358979323846264338327950288L * (i + 0.75L) / (n + 0.5L)), pp = 0, z1;
        do {
            long double p1 = 1, p2 = 0;
            for (int j = 0; j < n; ++j) {
                long double p3 = p2;
                p2 = p1;
                p1 = ((2 * j + 1) * z * p2 - j * p3) / (j + 1);
            }
            pp = n * (z * p1 - p2) / (z * z - 1);
            z1 = z;
            z  = z1 - p1 / pp;
        } while (fabsl(z - z1) > 1e-19L);
        x[i]         = (double)(-z);
        x[n - 1 - i] = (double)z;
        w[i] = w[n - 1 - i] = (double)(2 / ((1 - z * z) * pp * pp));
    }
}

cudaError_t upload_quadrature_tables() {
    std::vector<double> x(kGLTotal), w(kGLTotal);
    const int orders[8] = {8, 16, 32, 64, 96, 128, 12, 10};
    for (int o : orders) gauss_legendre(o, x.data() + gl_offset(o), w.data() + gl_offset(o));
    cudaError_t e = cudaMemcpyToSymbol(c_glx, x.data(), sizeof(double) * kGLTotal);
    if (e != cudaSuccess) return e;
    e = cudaMemcpyToSymbol(c_glw, w.data(), sizeof(double) * kGLTotal);
    if (e != cudaSuccess) return e;
    double ls[kLagTotal], lw[kLagTotal], les[kLagTotal];
    {
        const int orders_l[7] = {4, 6, 8, 12, 16, 24, 10};    // offsets 0, 4, 10, 18, 30, 46, 70 (thermal_integral)
        int off = 0;
        for (int o : orders_l) { gauss_laguerre(o, ls + off, lw + off); off += o; }
    }
    for (int i = 0; i < kLagTotal; ++i) les[i] = std::exp(-ls[i]);
    if ((e = cudaMemcpyToSymbol(c_lag_s, ls, sizeof(ls))) != cudaSuccess) return e;
    if ((e = cudaMemcpyToSymbol(c_lag_w, lw, sizeof(lw))) != cudaSuccess) return e;
    if ((e = cudaMemcpyToSymbol(c_lag_es, les, sizeof(les))) != cudaSuccess) return e;
    {   // log table of flog_tab (acro_fastquad.cuh): u_i = 1/c_i as a double, v_i = -log(u_i) - [i >= split] ln 2 in long double
        double2 lt[kLogTab];
        for (int i = 0; i < kLogTab; ++i) {
            const double u = 1.0 / (1.0 + (i + 0.5) / kLogTab);
            lt[i].x = u;
            lt[i].y = (double)(-logl((long double)u) - (i >= kLogTabSplit ? logl(2.0L) : 0.0L));
        }
        if ((e = cudaMemcpyToSymbol(g_logtab, lt, sizeof(lt))) != cudaSuccess) return e;
    }
    // Legendre tables of the 16-node panels (acro_tshared.cuh)
    double modal[kPanN][kPanN], rec[kPanN][2];
    const double* gx = x.data() + gl_offset(kPanN);
    const double* gw = w.data() + gl_offset(kPanN);
    for (int m = 0; m < kPanN; ++m) {
        long double pm = 1, pc = gx[m];
        for (int j = 0; j < kPanN; ++j) {
            const long double pj = (j == 0) ? 1.0L : (j == 1 ? (long double)gx[m] : 0.0L);
            long double val = pj;
            if (j >= 2) { val = ((2 * j - 1) * (long double)gx[m] * pc - (j - 1) * pm) / j; pm = pc; pc = val; }
            modal[j][m] = (double)((2 * j + 1) / 2.0L * gw[m] * val);
        }
    }
    for (int j = 0; j < kPanN; ++j) { rec[j][0] = (2.0 * j + 1.0) / (j + 1.0); rec[j][1] = (double)j / (j + 1.0); }
    {   // cross-section tables (nucl.py:217-347): {log(N Q^p1), p2, p3, Q}
        static const double gen[ACRO_NR][4] = {   // N, p1, p2, p3 (reactions with their own formula: N = 1, p = 0)
            {1, 0, 0, 0}, {9.8, 1.95, 1.65, 3.6}, {26.0, 2.6, 2.3, 4.9}, {8.88, 1.75, 1.65, 3.4}, {16.7, 1.95, 2.3, 4.25},
            {19.5, 3.5, 1.0, 4.5}, {20.7, 3.5, 1.0, 4.5}, {10.7, 10.2, 3.4, 13.6}, {21.7, 4.0, 3.0, 7.0}, {143.0, 2.3, 4.7, 7.0},
            {38.1, 3.0, 2.0, 5.0}, {1, 0, 0, 0}, {0.176, 1.51, 0.49, 2.0}, {122.0, 4.0, 3.0, 7.0}, {1, 0, 0, 0},
            {32.6, 10.0, 2.0, 12.0}, {133.0, 4.0, 3.0, 7.0}};
        static const double eth[ACRO_NR] = {2.224573, 6.257248, 8.481821, 5.493485, 7.718058, 19.813852, 20.577615, 23.846527,
                                            26.071100, 3.698892, 15.794685, 2.467032, 7.249962, 10.948850, 1.586627, 5.605794,
                                            9.304680};
        double sig[ACRO_NR][4], sig2[2][3];
        for (int r = 0; r < ACRO_NR; ++r) {
            sig[r][0] = std::log(gen[r][0]) + gen[r][1] * std::log(eth[r]);
            sig[r][1] = gen[r][2]; sig[r][2] = gen[r][3]; sig[r][3] = eth[r];
        }
        sig2[0][0] = std::log(1205.0) + 5.5 * std::log(eth[12]); sig2[0][1] = 5.0; sig2[0][2] = 10.5;
        sig2[1][0] = std::log(2.27e6) + 8.8335 * std::log(eth[15]); sig2[1][1] = 13.0; sig2[1][2] = 21.8335;
        if ((e = cudaMemcpyToSymbol(c_sig, sig, sizeof(sig))) != cudaSuccess) return e;
        if ((e = cudaMemcpyToSymbol(c_sig2, sig2, sizeof(sig2))) != cudaSuccess) return e;
    }
    if ((e = cudaMemcpyToSymbol(c_leg_modal, modal, sizeof(modal))) != cudaSuccess) return e;
    return cudaMemcpyToSymbol(c_leg_rec, rec, sizeof(rec));
}

// ------------------------------------------------------------------------------------------------------------
// small helpers
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ RateDb make_db(const DeviceTables& t) {
    RateDb db;
    db.tab = t.ratedb; db.nT = t.nT; db.nE = t.nE;
    db.Elog_min = t.Elog_min; db.Elog_max = t.Elog_max; db.Tlog_min = t.Tlog_min; db.Tlog_max = t.Tlog_max;
    return db;
}

// index of the system that owns flat entry `t` of a prefix array off[0..n) (off ascending, off[0]=0)
__device__ __forceinline__ int owner_of(const int64_t* __restrict__ off, int n, int64_t t) {
    int lo = 0, hi = n - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (off[mid] <= t) lo = mid; else hi = mid - 1;
    }
    return lo;
}


// ------------------------------------------------------------------------------------------------------------
// (i.a)+(v)  total rates G[3,NE] per system
//   photon:   photon-photon scattering + Compton + Bethe-Heitler + thermal pair creation (cascade.py:373-374)
//   e-/e+:    inverse Compton                                                           (cascade.py:497-498)
// ------------------------------------------------------------------------------------------------------------
struct RateEval {
    double g_ph, g_el;
    int direct;   // bit0: photon pair creation must come from the direct integral, bit1: inverse Compton
};

__device__ inline RateEval eval_rates(double E, double T, const DeviceTables& t, const acro_params_t& p) {
    const Densities d = densities(T, t.eta, t.Yp, t.YHe4);
    RateEval r;
    r.direct = 0;
    r.g_ph = rate_photon_photon(E, T) + rate_compton(E, d.ne) + rate_bethe_heitler(E, d.nNZ2);
    r.g_el = 0.0;
    const double E_log = log10(E), T_log = log10(T);
    const RateDb db = make_db(t);
    const bool indb = p.use_db && in_rate_db(db, E_log, T_log);
    if (!(E < kMe2 / (50.0 * T))) {   // pair-creation threshold (cascade.py:362)
        if (indb) r.g_ph += interp_rate_db(db, 0, E_log, T_log);
        else r.direct |= 1;
    }
    if (indb) r.g_el = interp_rate_db(db, 1, E_log, T_log);
    else r.direct |= 2;
    return r;
}

__global__ void __launch_bounds__(256) rates_kernel(acro_batch_t b, double* __restrict__ G, DeviceTables t, acro_params_t p,
                                                    DirectRateItem* __restrict__ items, int* __restrict__ item_count,
                                                    int* __restrict__ status_sys) {
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= b.n_sys_energy_total) return;
    const int s  = owner_of(b.sys_f_off, b.n_systems, tid);
    const int i  = (int)(tid - b.sys_f_off[s]);
    const int pt = b.sys_point[s];
    const int ne = b.pt_ne[pt];
    const double E = b.e_grid[b.pt_e_off[pt] + i];
    const double T = b.sys_T[s];
    const RateEval r = eval_rates(E, T, t, p);
    double* g = G + b.sys_f_off[s] * 3;
    g[i] = r.g_ph;
    g[ne + i] = r.g_el;
    g[2 * ne + i] = r.g_el;
    if (r.direct) {
        const int k = atomicAdd(item_count, 1);
        DirectRateItem it;
        it.g_index = b.sys_f_off[s] * 3 + i; it.ne = ne; it.kind = r.direct; it.E = E; it.T = T;
        items[k] = it;
        if (status_sys) status_sys[s] = 1;
    }
}

// ------------------------------------------------------------------------------------------------------------
// direct 2-D rate integrals (outside rates.db): one warp per item.
//
//  inverse Compton (cascade.py:469-485):  2 pi alpha^2/E^2 * int_0^ulim dx  x/(pi^2 (e^{x/T}-1)) int_x^{ymax(x)} dy F(y,E,x)
//     outer: GL-64 in log x on [log(1e-9 T), log ulim] (the integrand ~ x^2 as x->0);
//     inner: with w = 1 + Gamma q, y = E (1 - 1/w): GL-96 in s = log w on [log(1+Gamma q_min), log(1+Gamma)], dy = E/w ds.
//  photon pair creation (cascade.py:336-358): 1/(8E^2) int dx 1/(pi^2 (e^{x/T}-1)) int_{4me^2}^{4Ex} ds s sigma_pc(s)
//     outer: GL-128 in log x on [log(me^2/E), log(200 T)];
//     inner: u = log(s/4me^2) = v^2 (removes the sqrt threshold of sigma_pc), GL-32 in v.
//  Orders validated against the reference at eps=1e-9 (tools/proto/direct_rates_proto.py): <= 1e-7 relative.
// ------------------------------------------------------------------------------------------------------------
__device__ double direct_ic_lane(double E, double T, double Ephb_T_max, int lane) {
    const double ulim = fmin(E - kMe2 / (4.0 * E), Ephb_T_max * T);
    const double a = log(1e-9 * T), bq = log(ulim);
    const double h = 0.5 * (bq - a), m = 0.5 * (bq + a);
    const double* xo = c_glx + gl_offset(64); const double* wo = c_glw + gl_offset(64);
    const double* xi = c_glx + gl_offset(96); const double* wi = c_glw + gl_offset(96);
    double tot = 0.0;
    for (int k = lane; k < 64; k += 32) {
        const double x  = exp(m + h * xo[k]);
        const double Gm = 4.0 * x * E / kMe2;
        const double qmin = x / (Gm * (E - x));
        const double s0 = log1p(Gm * qmin), s1 = log1p(Gm);
        const double hs = 0.5 * (s1 - s0), ms = 0.5 * (s1 + s0);
        double inner = 0.0;
        for (int j = 0; j < 96; ++j) {
            const double s  = ms + hs * xi[j];
            const double em = expm1(s);                  // w - 1 to full relative accuracy (s can be ~1e-10)
            const double iw = frcp(1.0 + em);            // 1/w
            const double y  = E * em * iw;
            const double q  = y * frcp(Gm * (E - y));
            const double Gq = Gm * q;
            const double F  = 2.0 * q * flog(q) + (1.0 + 2.0 * q) * (1.0 - q) + Gq * Gq * (1.0 - q) * frcp(2.0 + 2.0 * Gq);
            inner += wi[j] * F * E * iw;
        }
        inner *= hs;
        tot += h * wo[k] * inner * x / (kPi2 * expm1(x / T)) * x;
    }
    return 2.0 * kPi * (kAlpha * kAlpha) * tot / (E * E);
}

__device__ double direct_pc_lane(double E, double T, double Ephb_T_max, int lane) {
    const double a = log(kMe2 / E), bq = log(Ephb_T_max * T);
    const double h = 0.5 * (bq - a), m = 0.5 * (bq + a);
    const double* xo = c_glx + gl_offset(128); const double* wo = c_glw + gl_offset(128);
    const double* xi = c_glx + gl_offset(32);  const double* wi = c_glw + gl_offset(32);
    double tot = 0.0;
    for (int k = lane; k < 128; k += 32) {
        const double x = exp(m + h * xo[k]);
        const double vmax = sqrt(log(E * x / kMe2));
        double inner = 0.0;
        for (int j = 0; j < 32; ++j) {
            const double v  = 0.5 * vmax * (1.0 + xi[j]);
            const double u  = v * v;
            const double b2 = -expm1(-u), bb = sqrt(b2);             // beta^2 = 1 - e^{-u} = 1 - 4 me^2/s
            const double s  = 4.0 * kMe2 * frcp(1.0 - b2);           // e^{u} = 1/(1 - beta^2)
            const double sig = 0.5 * kPi * (kRe * kRe) * (1.0 - b2) *
                               ((3.0 - b2 * b2) * flog((1.0 + bb) * frcp(1.0 - bb)) - 2.0 * bb * (2.0 - b2));
            inner += wi[j] * s * sig * s * 2.0 * v;
        }
        inner *= 0.5 * vmax;
        tot += h * wo[k] * inner / (kPi2 * expm1(x / T)) * x;
    }
    return tot / (8.0 * E * E);
}

__global__ void __launch_bounds__(128) direct_rates_kernel(double* __restrict__ G, const DirectRateItem* __restrict__ items,
                                                           const int* __restrict__ item_count, acro_params_t p) {
    const int n = *item_count;
    const int lane = threadIdx.x & 31;
    const int warps_total = (gridDim.x * blockDim.x) >> 5;
    for (int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < n; w += warps_total) {
        const DirectRateItem it = items[w];
        if (it.kind & 1) {
            const double v = warp_sum(direct_pc_lane(it.E, it.T, p.Ephb_T_max, lane));
            if (lane == 0) G[it.g_index] += v;
        }
        if (it.kind & 2) {
            const double v = warp_sum(direct_ic_lane(it.E, it.T, p.Ephb_T_max, lane));
            if (lane == 0) { G[it.g_index + it.ne] = v; G[it.g_index + 2 * (int64_t)it.ne] = v; }
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// (i.b)  kernel planes.  One thread per (system, i, j>=i).
//   plane 0  gamma->gamma : photon-photon scattering + Compton                          (cascade.py:439-442)
//   plane 1  e+-  ->gamma : inverse Compton, thermal integral                          (cascade.py:406-435)
//   plane 2  gamma->e-    : Compton (recoil electron) + Bethe-Heitler + gamma gamma_th  (cascade.py:644-647)
//   plane 3  gamma->e+    : Bethe-Heitler + gamma gamma_th                              (cascade.py:684-687)
//   plane 4  e+-  ->e+-   : inverse Compton, thermal integral                          (cascade.py:507-540)
// The thermal integrals run over log(eps_bar) with a Q-point Gauss-Legendre rule (Q = quad_order); the
// reference uses adaptive QAGS at epsrel=1e-3, which GL-64 reproduces to <=1.4e-10.
// ------------------------------------------------------------------------------------------------------------
#ifdef ACRO_CROSSCHECK
template <int Q>
__device__ __forceinline__ double integral_ic_photon(double E, double Ep, double T, double llim, double ulim) {
    const double a = log(llim), b = log(ulim);
    const double h = 0.5 * (b - a), m = 0.5 * (b + a);
    const double iT = 1.0 / T;
    double sum = 0.0;
#pragma unroll 2
    for (int k = 0; k < Q; ++k) {
        const double x = exp(fma(h, c_glx[gl_offset(Q) + k], m));
        const double f = ic_F(E, Ep, x) * (x * x) / (exp(x * iT) - 1.0);
        sum = fma(c_glw[gl_offset(Q) + k], f, sum);
    }
    return sum * h * (1.0 / kPi2);
}

template <int Q>
__device__ __forceinline__ double integral_ic_electron(double E, double Ep, double T, double llim, double ulim) {
    const double a = log(llim), b = log(ulim);
    const double h = 0.5 * (b - a), m = 0.5 * (b + a);
    const double iT = 1.0 / T;
    const double dE = Ep - E;
    double sum = 0.0;
#pragma unroll 2
    for (int k = 0; k < Q; ++k) {
        const double x = exp(fma(h, c_glx[gl_offset(Q) + k], m));
        const double f = ic_F(dE + x, Ep, x) * (x * x) / (exp(x * iT) - 1.0);
        sum = fma(c_glw[gl_offset(Q) + k], f, sum);
    }
    return sum * h * (1.0 / kPi2);
}

template <int Q>
__device__ __forceinline__ double integral_pc_electron(double E, double Ep, double T, double llim, double ulim) {
    const double a = log(llim), b = log(ulim);
    const double h = 0.5 * (b - a), m = 0.5 * (b + a);
    const double iT = 1.0 / T;
    double sum = 0.0;
#pragma unroll 2
    for (int k = 0; k < Q; ++k) {
        const double x = exp(fma(h, c_glx[gl_offset(Q) + k], m));
        const double f = pc_G(E, Ep, x) * x / (exp(x * iT) - 1.0);
        sum = fma(c_glw[gl_offset(Q) + k], f, sum);
    }
    return sum * h * (1.0 / kPi2);
}

struct KernelEntry { double gg, eg, gem, gep, ee; };

template <int Q>
__device__ __forceinline__ KernelEntry eval_kernels(double E, double Ep, double T, const Densities d, double Ephb_T_max) {
    KernelEntry k;
    const double xmax = Ephb_T_max * T;
    // gamma -> gamma
    k.gg = kernel_photon_photon(E, Ep, T) + kernel_compton_core(E, Ep, d.ne);
    // e -> gamma (inverse Compton), cascade.py:406-435
    k.eg = 0.0;
    const bool above = !(Ep < E + kMe);
    if (above) {
        const double llim = 0.25 * kMe2 * E / (Ep * Ep - Ep * E);
        const double ulim = fmin(fmin(E, Ep - kMe2 / (4.0 * Ep)), xmax);
        if (ulim > llim) k.eg = 2.0 * kPi * (kAlpha * kAlpha) * integral_ic_photon<Q>(E, Ep, T, llim, ulim) / (Ep * Ep);
    }
    // e -> e (inverse Compton), cascade.py:507-540
    k.ee = 0.0;
    if (E != Ep) {
        const double pf = 0.25 * kMe2 / Ep - E;
        const double qf = 0.25 * kMe2 * (Ep - E) / Ep;
        const double sq = sqrt((0.5 * pf) * (0.5 * pf) - qf);
        const double z1 = -0.5 * pf - sq, z2 = -0.5 * pf + sq;
        const double ulim = fmin(fmin(z2, Ep - kMe2 / (4.0 * Ep)), xmax);
        if (ulim > z1) k.ee = 2.0 * kPi * (kAlpha * kAlpha) * integral_ic_electron<Q>(E, Ep, T, z1, ulim) / (Ep * Ep);
    }
    // gamma -> e-/e+ : Bethe-Heitler (cascade.py:550-558) + gamma gamma_th -> e+ e- (cascade.py:562-598)
    double pair = 0.0;
    if (above) pair = d.nNZ2 * bh_dsdE_Z2(E, Ep);
    if (!(Ep < kMe2 / (50.0 * T))) {
        const double dE = Ep - E, rt = sqrt(E * E - kMe2), den = 4.0 * Ep * dE + kMe2;
        const double z1 = Ep * (kMe2 - 2.0 * dE * (rt - E)) / den;
        const double z2 = Ep * (kMe2 + 2.0 * dE * (rt + E)) / den;
        const double llim = fmax(kMe2 / Ep, z1);
        const double ulim = fmin(fmin(z2, Ep), xmax);
        if (ulim > llim)
            pair += 0.25 * kPi * (kAlpha * kAlpha) * kMe2 * integral_pc_electron<Q>(E, Ep, T, llim, ulim) / (Ep * Ep * Ep);
    }
    k.gep = pair;
    // Compton recoil electron: the photon formula with E -> Ep + me - E (cascade.py:621-639)
    k.gem = pair + kernel_compton_core(Ep + kMe - E, Ep, d.ne);
    return k;
}

// Production version: same limits and closed forms; thermal integrals either picked up from the temperature-shared
// producer kernel (acro_tshared.cuh) or evaluated here (Gauss-Laguerre / Gauss-Legendre, acro_fastquad.cuh).
struct SharedInfo {
    bool on;            // producer results are present in the K planes
    double x_bot, Tmin; // bottom of the point's shared grid, lowest temperature of the point
};

template <int Q>
__device__ __forceinline__ KernelEntry eval_kernels_fast(double E, double Ep, double T, const Densities d, double Ephb_T_max,
                                                         const SharedInfo si, const double* __restrict__ raw, int64_t np) {
    KernelEntry k;
    const double xmax = Ephb_T_max * T;
    PairConst c;
    c.E = E; c.Ep = Ep; c.dE = Ep - E;
#if ACRO_FQ_LOGTAB
    c.lt = g_logtab;
#endif
    k.gg = kernel_photon_photon(E, Ep, T) + kernel_compton_core(E, Ep, d.ne);
    k.eg = 0.0;
    k.ee = 0.0;
    const double pre_ic = 2.0 * kPi * (kAlpha * kAlpha) / (Ep * Ep);
    const PairLimits L = pair_limits(E, Ep);
    if (L.xa9 > 0.0) {
        const double ulim = fmin(L.xb9, xmax);
        if (ulim > L.xa9) {
            double I;
            if (si.on && shared_regime(L.xa9, kSteepX, T, si.x_bot, si.Tmin)) I = raw[1 * np];
            else {
                const double r = E / c.dE;
                c.c9 = r * r / (2.0 + 2.0 * r);
                I = thermal_integral<KIND_IC_PHOTON, Q>(c, T, L.xa9, ulim);
            }
            k.eg = pre_ic * I;
        }
    }
    if (L.xa10 > 0.0) {
        const double ulim = fmin(L.xb10, xmax);
        if (ulim > L.xa10) {
            double I;
            if (si.on && shared_regime(L.xa10, kSteepX, T, si.x_bot, si.Tmin)) I = raw[4 * np];
            else {
                c.c10 = 0.25 * kMe2 / Ep;
                c.i2Ep = 0.5 / Ep;
                I = thermal_integral<KIND_IC_ELECTRON, Q>(c, T, L.xa10, ulim);
            }
            k.ee = pre_ic * I;
        }
    }
    double pair = 0.0;
    if (L.xa9 > 0.0) pair = d.nNZ2 * bh_dsdE_Z2(E, Ep);     // same condition E' >= E + me (cascade.py:554)
    if (!(Ep < kMe2 / (50.0 * T))) {
        const double ulim = fmin(L.xb11, xmax);
        if (ulim > L.xa11) {
            double I;
            if (si.on && shared_regime(L.xa11, kSteepX, T, si.x_bot, si.Tmin)) I = raw[3 * np];
            else I = thermal_integral<KIND_PC_ELECTRON, Q>(c, T, L.xa11, ulim);
            pair += 0.25 * kPi * (kAlpha * kAlpha) * kMe2 * I / (Ep * Ep * Ep);
        }
    }
    k.gep = pair;
    k.gem = pair + kernel_compton_core(Ep + kMe - E, Ep, d.ne);
    return k;
}

#ifndef ACRO_KK_MINBLOCKS
#define ACRO_KK_MINBLOCKS 6
#endif
template <int Q, bool FAST>
__global__ void __launch_bounds__(128, ACRO_KK_MINBLOCKS) kernels_kernel(acro_batch_t b, double* __restrict__ Kws, DeviceTables t, acro_params_t p) {
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= b.n_pairs_total) return;
    const int s  = owner_of(b.sys_k_off, b.n_systems, tid);
    const int64_t loc = tid - b.sys_k_off[s];
    const int pt = b.sys_point[s];
    const int ne = b.pt_ne[pt];
    if (loc >= (int64_t)ne * (ne + 1) / 2) return;   // pad entries between systems (sector alignment)
    // invert loc = i*ne - i(i-1)/2 + (j-i)
    const double tn = 2.0 * ne + 1.0;
    int i = (int)floor(0.5 * (tn - sqrt(tn * tn - 8.0 * (double)loc)));
    if (i < 0) i = 0;
    if (i > ne - 1) i = ne - 1;
    while (i > 0 && tri_row_off(i, ne) > loc) --i;
    while (i < ne - 1 && tri_row_off(i + 1, ne) <= loc) ++i;
    const int j = i + (int)(loc - tri_row_off(i, ne));
    const double* eg = b.e_grid + b.pt_e_off[pt];
    const double E = eg[i], Ep = eg[j], T = b.sys_T[s];
    const Densities d = densities(T, t.eta, t.Yp, t.YHe4);
    const int64_t np = b.n_pairs_total;
    KernelEntry k;
    if (FAST) {
        SharedInfo si;
        si.on = false;   // per-(pair,T) evaluation only (ACRO_KERNELS_PER_T); the FAST mode runs kernels_mma_kernel instead
        const int s_first = b.pt_sys_off[pt], nt = b.pt_nt[pt];
        si.Tmin = b.sys_T[s_first];
        const SharedGrid g = shared_grid(si.Tmin, b.sys_T[s_first + nt - 1], eg[ne - 1], p.Ephb_T_max);
        si.x_bot = exp(g.y_bot);
        k = eval_kernels_fast<Q>(E, Ep, T, d, p.Ephb_T_max, si, Kws + tid, np);
    } else {
        k = eval_kernels<Q>(E, Ep, T, d, p.Ephb_T_max);
    }
    Kws[tid]          = k.gg;
    Kws[np + tid]     = k.eg;
    Kws[2 * np + tid] = k.gem;
    Kws[3 * np + tid] = k.gep;
    Kws[4 * np + tid] = k.ee;
}

#endif  // ACRO_CROSSCHECK

// Counts, for the roofline's algorithmic-work figure, the (system, i, j>=i) pairs and how many of them
// have a non-empty integration range for each of the three thermal integrals -- the same limit tests as
// eval_kernels, no quadrature.  out[0..3] = pairs, N_a9 (e->gamma), N_a10 (e->e), N_a11 (gamma->e).
__global__ void __launch_bounds__(128) count_work_kernel(acro_batch_t b, unsigned long long* __restrict__ out, acro_params_t p) {
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int c9 = 0, c10 = 0, c11 = 0, cp = 0;
    if (tid < b.n_pairs_total) {
        const int s  = owner_of(b.sys_k_off, b.n_systems, tid);
        const int64_t loc = tid - b.sys_k_off[s];
        const int pt = b.sys_point[s];
        const int ne = b.pt_ne[pt];
        const bool pad = loc >= (int64_t)ne * (ne + 1) / 2;     // pad entries between systems (sector alignment)
        const double tn = 2.0 * ne + 1.0;
        int i = (int)floor(0.5 * (tn - sqrt(tn * tn - 8.0 * (double)(pad ? 0 : loc))));
        i = max(0, min(i, ne - 1));
        while (i > 0 && tri_row_off(i, ne) > loc) --i;
        while (i < ne - 1 && tri_row_off(i + 1, ne) <= loc) ++i;
        const int j = i + (int)(loc - tri_row_off(i, ne));
        const double* eg = b.e_grid + b.pt_e_off[pt];
        const double E = eg[i], Ep = eg[j], T = b.sys_T[s], xmax = p.Ephb_T_max * T;
        cp = pad ? 0 : 1;
        if (!pad && !(Ep < E + kMe)) {
            const double llim = 0.25 * kMe2 * E / (Ep * Ep - Ep * E);
            c9 = fmin(fmin(E, Ep - kMe2 / (4.0 * Ep)), xmax) > llim;
        }
        if (!pad && E != Ep) {
            const double pf = 0.25 * kMe2 / Ep - E, qf = 0.25 * kMe2 * (Ep - E) / Ep;
            const double sq = sqrt((0.5 * pf) * (0.5 * pf) - qf);
            c10 = fmin(fmin(-0.5 * pf + sq, Ep - kMe2 / (4.0 * Ep)), xmax) > (-0.5 * pf - sq);
        }
        if (!pad && !(Ep < kMe2 / (50.0 * T))) {
            const double dE = Ep - E, rt = sqrt(E * E - kMe2), den = 4.0 * Ep * dE + kMe2;
            const double z1 = Ep * (kMe2 - 2.0 * dE * (rt - E)) / den, z2 = Ep * (kMe2 + 2.0 * dE * (rt + E)) / den;
            c11 = fmin(fmin(z2, Ep), xmax) > fmax(kMe2 / Ep, z1);
        }
    }
    const unsigned m9 = __ballot_sync(0xffffffffu, c9), m10 = __ballot_sync(0xffffffffu, c10);
    const unsigned m11 = __ballot_sync(0xffffffffu, c11), mp = __ballot_sync(0xffffffffu, cp);
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(out + 0, (unsigned long long)__popc(mp));
        atomicAdd(out + 1, (unsigned long long)__popc(m9));
        atomicAdd(out + 2, (unsigned long long)__popc(m10));
        atomicAdd(out + 3, (unsigned long long)__popc(m11));
    }
}

// dense copy of one system's planes (debug / parity checks)
__global__ void unpack_kernels_kernel(const double* __restrict__ Kws, int64_t np, int64_t k_off, int ne, double* __restrict__ out) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= 5 * ne * ne) return;
    const int k = idx / (ne * ne), r = idx % (ne * ne), i = r / ne, j = r % ne;
    out[idx] = (j >= i) ? Kws[(int64_t)k * np + k_off + tri_row_off(i, ne) + (j - i)] : 0.0;
}

// ------------------------------------------------------------------------------------------------------------
// (ii)+(iii)  cascade back-substitution fused with the photodisintegration-rate integrals.  One warp per system.
//
// Cascade equation on the log-energy grid, trapezoid weights (cascade.py:180-239): rows are solved from the top
// energy down; the sum over j>i is split across the lanes and combined with warp shuffles; the 3x3 system per row
// is solved with partial pivoting.  wF[X][j] holds weight_j * F_X(E_j) (for j = NE-1 the delta-source term
// S0_X/Gamma_X(E0) is folded in), so one fused multiply-add per kernel entry remains.
// Then Gamma_r = S0_gamma sigma_r(E0)/Gamma_gamma(E0) + int_{Eth}^{Emax} F_gamma(E) E sigma_r(E) dlogE with
// F_gamma the log-log interpolant of the grid values (nucl.py:398-467, utils.py:38-61), integrated exactly cell
// by cell (GL-8 per cell; the threshold cell with a quadratic substitution) instead of adaptive QAGS.
// ------------------------------------------------------------------------------------------------------------
// Cross-sections in mb for a RUN-TIME reaction index (one copy of the code: the 17-way inlined version made the solve
// kernel ~290 kB of SASS and instruction-cache bound, ncu: 8 warps stalled on no_instruction per issue).  Same formulas
// as cross_section_mb (acro_physics.cuh); power laws as exp(c + p2 log(E-Q) - p3 log E) with the fast exp/log.

__device__ __noinline__ double sigma_rt(int r, double E, double lnE) {   // r = 0..16
    const double Q = c_sig[r][3];
    if (!(E > Q)) return 0.0;
    const double lt = flog(E - Q);
    const double gen = fexp(fmax(c_sig[r][0] + c_sig[r][1] * lt - c_sig[r][2] * lnE, -700.0));
    switch (r) {
        case 0: {   // d+a>n+p
            const double rr = sqrt(Q * (E - Q)) / E;
            const double sq = sqrt(Q) - sqrt(0.037);
            return 18.75 * (rr * rr * rr + 0.007947 * (rr * rr) * (sq * sq / (E - Q + 0.037)));
        }
        case 10: {  // Li6+a>X
            const double g1 = (E - 19.0) / 3.5, g2 = (E - 30.0) / 3.0, g3 = (E - 43.0) / 5.0;
            return gen * (3.7 * exp(-0.5 * g1 * g1) + 2.75 * exp(-0.5 * g2 * g2) + 2.2 * exp(-0.5 * g3 * g3));   // exp: may underflow
        }
        case 11: {  // Li7+a>t+He4
            const double Ecm = E - Q, e2 = Ecm * Ecm;
            const double pp = 1.0 + 2.2875 * e2 - 1.1798 * (e2 * Ecm) + 2.5279 * (e2 * e2);
            return 0.105 * (2371.0 / (E * E)) * exp(-2.5954 / sqrt(Ecm) - 2.056 * Ecm) * pp;
        }
        case 12: {  // Li7+a>n+Li6
            const double z = (E - Q - 7.46) / 0.188;
            return gen + fexp(fmax(c_sig2[0][0] + c_sig2[0][1] * lt - c_sig2[0][2] * lnE, -700.0)) + 0.06 / (1.0 + z * z);
        }
        case 14: {  // Be7+a>He3+He4
            const double Ecm = E - Q, e2 = Ecm * Ecm;
            const double pp = 1.0 - 0.428 * e2 + 0.534 * (e2 * Ecm) - 0.115 * (e2 * e2);
            if (pp < 0.0) return 0.0;
            return 0.504 * (2371.0 / (E * E)) * exp(-5.1909 / sqrt(Ecm) - 0.548 * Ecm) * pp;
        }
        case 15:    // Be7+a>p+Li6
            return gen + fexp(fmax(c_sig2[1][0] + c_sig2[1][1] * lt - c_sig2[1][2] * lnE, -700.0));
        default:
            return gen;
    }
}

__device__ __forceinline__ void solve3(double B[3][3], double a[3], double x[3]) {
    // Gaussian elimination with partial pivoting (the reference calls LAPACK dgesv, cascade.py:221)
    int p[3] = {0, 1, 2};
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        int piv = c;
        double best = fabs(B[p[c]][c]);
#pragma unroll
        for (int r = c + 1; r < 3; ++r) {
            const double v = fabs(B[p[r]][c]);
            if (v > best) { best = v; piv = r; }
        }
        const int tmp = p[c]; p[c] = p[piv]; p[piv] = tmp;
#pragma unroll
        for (int r = c + 1; r < 3; ++r) {
            const double f = B[p[r]][c] / B[p[c]][c];
#pragma unroll
            for (int cc = c; cc < 3; ++cc) B[p[r]][cc] -= f * B[p[c]][cc];
            a[p[r]] -= f * a[p[c]];
        }
    }
#pragma unroll
    for (int c = 2; c >= 0; --c) {
        double v = a[p[c]];
#pragma unroll
        for (int cc = c + 1; cc < 3; ++cc) v -= B[p[c]][cc] * x[cc];
        x[c] = v / B[p[c]][c];
    }
}

__device__ __forceinline__ double sc_value(const acro_batch_t& b, int s, int pt, int ne, int X, int i) {
    if (b.sc_mode == ACRO_SC_NONE) return 0.0;
    if (b.sc_mode == ACRO_SC_SEPARABLE) return b.sc[(int64_t)s * 3 + X] * b.sc_E[(b.pt_e_off[pt] + i) * 3 + X];
    return b.sc[b.sys_f_off[s] * 3 + (int64_t)X * ne + i];
}

// The 17 cross-sections at the quadrature nodes of an energy cell depend on the point's energy grid only, not on the
// temperature: one table per point, shared by its NT systems (they were 29 % of solve_pdi_kernel's instructions, once per
// system).  Layout sig[(k * ACRO_NR + r) * n_energy_total + pt_e_off + c]: the lanes of solve_pdi_kernel walk over cells c.
// Same expressions for the node positions as part (A) of solve_pdi_kernel, which uses the table for its ordinary cells
// (one piece, not clipped by Emax) and evaluates the others itself.
__global__ void __launch_bounds__(128) pdi_sigma_kernel(acro_batch_t b, double* __restrict__ sig, acro_params_t p) {
    const int pt = blockIdx.y;
    const int ne = b.pt_ne[pt];
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c == ne - 1) {
        // the column of the last grid point is not a cell: its k = 0 slots carry sigma_r(E0) for the delta-source term of the
        // rates (nucl.py:441) -- once per point instead of 17 serial evaluations by one lane of every system's warp
        const double E0 = b.pt_e0[pt], lnE0 = log(E0);
        for (int r = 0; r < ACRO_NR; ++r) sig[(int64_t)r * b.n_energy_total + b.pt_e_off[pt] + c] = sigma_rt(r, E0, lnE0);
    }
    if (c >= ne - 1) return;
    const double* E = b.e_grid + b.pt_e_off[pt];
    const double y0 = log(E[c]), y1 = log(E[c + 1]);
    const double h = 0.5 * (y1 - y0), m = y0 + h;
    const int nq = p.pdi_cell_order;
    const double* qx = c_glx + gl_offset(nq);
    double* out = sig + b.pt_e_off[pt] + c;
    for (int k = 0; k < nq; ++k) {
        const double y = fma(h, qx[k], m);
        const double Ev = fexp(y);
        for (int r = 0; r < ACRO_NR; ++r) out[(int64_t)(k * ACRO_NR + r) * b.n_energy_total] = sigma_rt(r, Ev, y);
    }
}

#ifndef ACRO_SV_MINBLOCKS
#define ACRO_SV_MINBLOCKS 5
#endif
__global__ void __launch_bounds__(128, ACRO_SV_MINBLOCKS) solve_pdi_kernel(acro_batch_t b, const double* __restrict__ G, double* __restrict__ Fout,
                                                        double* __restrict__ pdi, const double* __restrict__ Kws,
                                                        const double* __restrict__ sig, acro_params_t p, int smem_stride) {
    extern __shared__ double smem[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int s = blockIdx.x * (blockDim.x >> 5) + wib;
    if (s >= b.n_systems) return;
    const int pt = b.sys_point[s];
    const int ne = b.pt_ne[pt];
    const double* E = b.e_grid + b.pt_e_off[pt];
    const double T = b.sys_T[s];
    const double* g = G + b.sys_f_off[s] * 3;
    // per-warp shared memory: lnF[ne] | wF[3][ne] during the back-substitution; afterwards the wF region is reused for
    // lnE[ne] and the 17x32 rate accumulators (stride = ne_max + max(3 ne_max, ne_max + 17*32))
    double* lnF = smem + (size_t)wib * smem_stride;   // [ne]  floored photon spectrum during the solve, its log afterwards
    double* wF  = lnF + ne;                           // [3][ne]
    double* lnE = wF;                                 // [ne]   (after the solve)
    const int64_t np = b.n_pairs_total;
    const double* K0 = Kws + b.sys_k_off[s];
    const double S0[3] = {b.s0[(int64_t)s * 3], b.s0[(int64_t)s * 3 + 1], b.s0[(int64_t)s * 3 + 2]};
    const int top = ne - 1;
    const double dy = log(E[top] / p.Emin) / (double)(ne - 1);
    double* fo = Fout ? Fout + b.sys_f_off[s] * 3 : nullptr;

    if (p.universal) {
        // flags.universal (cascade.py:774-812): F_gamma(E) = SN K0 (EX/E)^{3/2 | 2} / Gamma_gamma(E,T) below / above EX =
        // me^2/(80T), zero (floored) above 1.05 EC with EC = me^2/(22T); no cascade equation.  Leptons are not defined.
        const double EC = kMe2 / (22.0 * T), EX = kMe2 / (80.0 * T);
        const double K0u = b.pt_e0[pt] / ((EX * EX) * (2.0 + log(EC / EX)));
        const double SN = S0[0] + S0[1] + S0[2];
        for (int i = lane; i < ne; i += 32) {
            const double Ei = E[i];
            double F = 0.0;
            if (Ei < EX) F = SN * K0u * pow(EX / Ei, 1.5) / g[i];
            else if (Ei <= (1.0 + 5e-2) * EC) F = SN * K0u * pow(EX / Ei, 2.0) / g[i];
            F = fmax(F, p.approx_zero);
            lnF[i] = F;
            if (fo) { fo[i] = F; fo[ne + i] = p.approx_zero; fo[2 * ne + i] = p.approx_zero; }
        }
        __syncwarp();
    } else {
    // --- top row (cascade.py:194-198)
    {
        const int64_t r = tri_row_off(top, ne);
        const double kgg = K0[r], keg = K0[np + r], kgm = K0[2 * np + r], kgp = K0[3 * np + r], kee = K0[4 * np + r];
        const double G0 = g[top], G1 = g[ne + top], G2 = g[2 * ne + top];
        double F[3];
        F[0] = sc_value(b, s, pt, ne, 0, top) / G0 + (kgg * S0[0] / (G0 * G0) + keg * S0[1] / (G1 * G0) + keg * S0[2] / (G2 * G0));
        F[1] = sc_value(b, s, pt, ne, 1, top) / G1 + (kgm * S0[0] / (G0 * G1) + kee * S0[1] / (G1 * G1));
        F[2] = sc_value(b, s, pt, ne, 2, top) / G2 + (kgp * S0[0] / (G0 * G2) + kee * S0[2] / (G2 * G2));
        if (lane == 0) {
            const double hw = 0.5 * dy * E[top];
            wF[top]          = S0[0] / G0 + hw * F[0];
            wF[ne + top]     = S0[1] / G1 + hw * F[1];
            wF[2 * ne + top] = S0[2] / G2 + hw * F[2];
            lnF[top] = fmax(F[0], p.approx_zero);
            if (fo) {
                fo[top] = fmax(F[0], p.approx_zero);
                fo[ne + top] = fmax(F[1], p.approx_zero);
                fo[2 * ne + top] = fmax(F[2], p.approx_zero);
            }
        }
    }
    __syncwarp();

    // --- rows NE-2 .. 0 (cascade.py:201-224)
    for (int i = top - 1; i >= 0; --i) {
        const int64_t r = tri_row_off(i, ne);
        double a0 = 0.0, a1 = 0.0, a2 = 0.0;
        for (int j = i + 1 + lane; j < ne; j += 32) {
            const int64_t e = r + (j - i);
            const double f0 = wF[j], f1 = wF[ne + j], f2 = wF[2 * ne + j];
            const double kee = K0[4 * np + e];
            a0 = fma(K0[e], f0, fma(K0[np + e], f1 + f2, a0));
            a1 = fma(K0[2 * np + e], f0, fma(kee, f1, a1));
            a2 = fma(K0[3 * np + e], f0, fma(kee, f2, a2));
        }
        a0 = warp_sum(a0); a1 = warp_sum(a1); a2 = warp_sum(a2);
        const double Ei = E[i], hw = 0.5 * dy * Ei;
        const double kgg = K0[r], keg = K0[np + r], kgm = K0[2 * np + r], kgp = K0[3 * np + r], kee = K0[4 * np + r];
        double B[3][3] = {{g[i] - hw * kgg, -hw * keg, -hw * keg},
                          {-hw * kgm, g[ne + i] - hw * kee, 0.0},
                          {-hw * kgp, 0.0, g[2 * ne + i] - hw * kee}};
        double a[3] = {a0 + sc_value(b, s, pt, ne, 0, i), a1 + sc_value(b, s, pt, ne, 1, i), a2 + sc_value(b, s, pt, ne, 2, i)};
        double F[3];
        {
            // The matrix is an arrow (electrons and positrons couple only through the photons) with a dominant diagonal
            // (total rate minus half a cell of the diagonal kernel): eliminate the two lepton rows by their own pivots
            // and solve the 1x1 Schur complement -- 3 reciprocals instead of a pivoted 3x3 elimination with 6 IEEE
            // divisions, which was 21 % of this kernel.  The general solver stays for a matrix without that structure.
            const double d1 = B[1][1], d2 = B[2][2];
            bool fast = d1 > 1e-280 && d2 > 1e-280 && d1 < 1e280 && d2 < 1e280 && fabs(B[1][0]) <= d1 && fabs(B[2][0]) <= d2;
            if (fast) {
                const double i1 = frcp(d1), i2 = frcp(d2);
                const double S = B[0][0] - B[0][1] * (B[1][0] * i1 + B[2][0] * i2);
                fast = S > 0.25 * fabs(B[0][0]) && S < 1e300;
                if (fast) {
                    F[0] = (a[0] - B[0][1] * (a[1] * i1 + a[2] * i2)) * frcp(S);
                    F[1] = (a[1] - B[1][0] * F[0]) * i1;
                    F[2] = (a[2] - B[2][0] * F[0]) * i2;
                }
            }
            if (!fast) solve3(B, a, F);
        }
        if (lane == 0) {
            const double w = dy * Ei;
            wF[i] = w * F[0]; wF[ne + i] = w * F[1]; wF[2 * ne + i] = w * F[2];
            lnF[i] = fmax(F[0], p.approx_zero);     // the logarithm is taken after the loop, by all lanes: one log per row on
                                                     // lane 0 was ~300 cycles on the critical path of every row
            if (fo) {
                fo[i] = fmax(F[0], p.approx_zero);
                fo[ne + i] = fmax(F[1], p.approx_zero);
                fo[2 * ne + i] = fmax(F[2], p.approx_zero);
            }
        }
        __syncwarp();
    }
    }   // !universal
    if (!pdi) return;
    for (int i = lane; i < ne; i += 32) {     // wF is dead from here on
        lnF[i] = log(lnF[i]);
        lnE[i] = log(E[i]);
    }
    __syncwarp();

    // --- photodisintegration rates (nucl.py:398-467)
    const double E0 = b.pt_e0[pt];
    const double EC = kMe2 / (22.0 * T);
    const double Emax = p.universal ? fmin(E0, EC) : fmin(E0, p.E_EC_max * EC);    // nucl.py:399-402, :420-422
    const double lnEmax = log(Emax);
    const int nq = p.pdi_cell_order;   // 8 or 16
    const double* qx = c_glx + gl_offset(nq);
    const double* qw = c_glw + gl_offset(nq);
    double* accs = lnE + ne;             // [17][32] per-lane accumulators in shared memory (run-time reaction index)
    for (int r = 0; r < ACRO_NR; ++r) accs[r * 32 + lane] = 0.0;
    // (A) cells that lie entirely above a reaction's threshold: nodes shared by all reactions.  Lane = cell for the ordinary
    //     cells; the few cells that need pieces or their own cross-section values are done by the whole warp together.
    for (int c0 = 0; c0 < ne - 1; c0 += 32) {
        const int c = c0 + lane;
        bool coop = false;                                     // this lane's cell goes to the cooperative path below
        double kap = 0.0, bs = 0.0, start = 0.0, piece = 0.0;
        int nsub = 0;
        unsigned open = 0;
        if (c < ne - 1) {
            const double y0 = lnE[c], y1 = lnE[c + 1];
            const double hi = fmin(y1, lnEmax);
            if (hi > y0) {
                const double ms = (lnF[c + 1] - lnF[c]) / (y1 - y0);
                bs = lnF[c + 1] - ms * y1;
                const double Elo = E[c];
                const bool kink_cell = (Elo < kKink15 && kKink15 < E[c + 1]);   // reaction 15 is integrated separately there
                // integrand = exp(kap y + bs) * sigma(e^y), kap = ms + 1.  Next to a floored spectrum value (e.g. F_gamma(E0)
                // = 0 for pure e+e- injection) kap reaches ~1e4 and the cell's weight sits within 1/|kap| of one end.  The
                // cell is therefore cut, starting from its large end, into pieces over which the exponent changes by <= 2 (at
                // most 40 e-folds are kept: the rest is < e^-40 of the cell), each piece with the plain nq-point rule.
                // Ordinary cells (|kap| * width <= 2) are a single piece.
                kap = ms + 1.0;
                const double width = hi - y0;
                const double len = fmin(width, 40.0 / fmax(fabs(kap), 1e-300));
                nsub = max(1, (int)ceil(0.5 * fabs(kap) * len));
                piece = len / (double)nsub;
                start = (kap > 0.0) ? hi - len : y0;           // pieces cover [start, start + len]
                for (int r = 0; r < ACRO_NR; ++r)
                    if (kEth[r] <= Elo && !(r == 14 && kink_cell)) open |= 1u << r;
                if (open && sig && nsub == 1 && hi == y1) {
                    // ordinary cell: node positions are the ones of pdi_sigma_kernel, cross-sections from the point's table
                    const double h = 0.5 * (y1 - y0), m = y0 + h;
                    const double* sc = sig + b.pt_e_off[pt] + c;
#pragma unroll 1
                    for (int k = 0; k < nq; ++k) {
                        const double fe = qw[k] * h * exp(fma(kap, fma(h, qx[k], m), bs));
                        // the 17 table values of the node are requested together (L2-resident table, one coalesced line per
                        // reaction): with one load per trip of a rolled loop every FMA waited for its own load -- 27 % of the
                        // kernel's stall samples sat on this line (profiles/r02_ncu_solve_pdi_kernel_v1.txt)
                        double v[ACRO_NR];
#pragma unroll
                        for (int r = 0; r < ACRO_NR; ++r) v[r] = sc[(int64_t)(k * ACRO_NR + r) * b.n_energy_total];
#pragma unroll
                        for (int r = 0; r < ACRO_NR; ++r)
                            if (open & (1u << r)) accs[r * 32 + lane] = fma(fe, v[r], accs[r * 32 + lane]);
                    }
                } else if (open) {
                    coop = true;
                }
            }
        }
        // cells cut into pieces (up to 20 x nq nodes x 17 cross-sections: ~1 ms on one lane -- it WAS the solve stage of a
        // single e+e- point, 1.1 of 2.0 ms device latency) and cells clipped by Emax: (piece, node) items dealt to the lanes
        unsigned todo = __ballot_sync(0xffffffffu, coop);
        while (todo) {
            const int src = __ffs(todo) - 1;
            todo &= todo - 1;
            const double kap_s = __shfl_sync(0xffffffffu, kap, src), bs_s = __shfl_sync(0xffffffffu, bs, src);
            const double start_s = __shfl_sync(0xffffffffu, start, src), piece_s = __shfl_sync(0xffffffffu, piece, src);
            const int nsub_s = __shfl_sync(0xffffffffu, nsub, src);
            const unsigned open_s = __shfl_sync(0xffffffffu, open, src);
            const double h = 0.5 * piece_s;
            for (int it = lane; it < nsub_s * nq; it += 32) {
                const int q = it / nq, k = it - q * nq;
                const double m = start_s + piece_s * (double)q + h;
                const double y  = fma(h, qx[k], m);
                const double Ev = fexp(y);
                const double fe = qw[k] * h * exp(fma(kap_s, y, bs_s));
#pragma unroll 1
                for (int r = 0; r < ACRO_NR; ++r)
                    if (open_s & (1u << r)) accs[r * 32 + lane] = fma(fe, sigma_rt(r, Ev, y), accs[r * 32 + lane]);
            }
        }
    }
    // (B) the threshold cell of each reaction, [log Eth_r, min(cell end, log Emax)], with y = lo + (hi-lo) u^2, which
    //     regularises the (E-Q)^p onset of the cross-sections, and
    // (C) reaction 15 (Be7+a>He3+He4): its closing polynomial changes sign at E = kKink15, where the reference clamps the
    //     cross-section to zero (nucl.py:327-335) -- a kink inside one grid cell; that cell is integrated only up to the kink
    //     (the part above it is identically zero) so that no rule straddles it.
    // Lane r < 17 sets up segment r (B), lane 17 the kink segment (C); their 16-node rules are then dealt to all lanes as
    // (segment, node) items: two reactions per batch of 32 instead of 17 lanes each running its own cross-section formula
    // one after the other (that was 0.6 of this kernel's 6.6 ms per benchmark slice).
    double seg_lo = 0.0, seg_hi = 0.0, seg_ms = 0.0, seg_bs = 0.0;
    bool seg_on = false;
    if (lane < ACRO_NR) {
        const double Eth = kEth[lane];
        if (Emax > Eth && Eth > E[0]) {
            const double lo = log(Eth);
            // c = last grid index with E[c] < Eth (cells with E[c] >= Eth are covered by part A)
            int c = (int)((lo - lnE[0]) * (double)(ne - 1) / (lnE[ne - 1] - lnE[0]));
            c = max(0, min(c, ne - 2));
            while (c > 0 && !(E[c] < Eth)) --c;
            while (c < ne - 2 && E[c + 1] < Eth) ++c;
            const double y0 = lnE[c], y1 = lnE[c + 1];
            const double hi = fmin(y1, lnEmax);
            if (hi > lo) {
                seg_ms = (lnF[c + 1] - lnF[c]) / (y1 - y0);
                seg_bs = lnF[c + 1] - seg_ms * y1;
                seg_lo = lo; seg_hi = hi; seg_on = true;
            }
        }
    }
    if (lane == ACRO_NR && E[0] < kKink15 && kKink15 < E[top]) {
        int c = (int)((log(kKink15) - lnE[0]) * (double)(ne - 1) / (lnE[ne - 1] - lnE[0]));
        c = max(0, min(c, ne - 2));
        while (c > 0 && !(E[c] < kKink15)) --c;
        while (c < ne - 2 && E[c + 1] < kKink15) ++c;
        if (kEth[14] <= E[c]) {      // otherwise the threshold-cell rule (B) owns this cell
            const double y0 = lnE[c], y1 = lnE[c + 1];
            const double hi = fmin(fmin(y1, lnEmax), log(kKink15));
            if (hi > y0) {
                seg_ms = (lnF[c + 1] - lnF[c]) / (y1 - y0);
                seg_bs = lnF[c + 1] - seg_ms * y1;
                seg_lo = y0; seg_hi = hi; seg_on = true;
            }
        }
    }
    {
        const unsigned segs = __ballot_sync(0xffffffffu, seg_on);
        const int n_items = __popc(segs) * 16;
        const double* tx = c_glx + gl_offset(16);
        const double* tw = c_glw + gl_offset(16);
        for (int base = 0; base < n_items; base += 32) {
            const int it = base + lane;
            const bool live = it < n_items;
            const int src = live ? __fns(segs, 0, (it >> 4) + 1) : 0;
            const double lo = __shfl_sync(0xffffffffu, seg_lo, src), hi = __shfl_sync(0xffffffffu, seg_hi, src);
            const double ms = __shfl_sync(0xffffffffu, seg_ms, src), bs = __shfl_sync(0xffffffffu, seg_bs, src);
            if (live) {
                const int k = it & 15;
                const int r = src < ACRO_NR ? src : 14;
                double y, wgt;
                if (src < ACRO_NR) {
                    const double u = 0.5 * (1.0 + tx[k]);
                    y = fma(hi - lo, u * u, lo);
                    wgt = tw[k] * 0.5 * (hi - lo) * 2.0 * u;
                } else {
                    const double h = 0.5 * (hi - lo);
                    y = fma(h, tx[k], 0.5 * (hi + lo));
                    wgt = tw[k] * h;
                }
                const double Ev = exp(y);
                accs[r * 32 + lane] = fma(wgt * exp(fma(ms, y, bs)) * Ev, sigma_rt(r, Ev, y), accs[r * 32 + lane]);
            }
        }
    }
    const double rate_E0 = g[top];   // Gamma_gamma(E0, T): the grid's last point is E0 (cascade.py:745, nucl.py:436)
    const double lnE0 = log(E0);
    __syncwarp();
    if (lane < ACRO_NR) {
        // lane r adds up the 32 per-lane partial sums of reaction r (columns visited in a rotated order: no bank conflicts)
        const int r = lane;
        double v = 0.0;
#pragma unroll 8
        for (int j = 0; j < 32; ++j) v += accs[r * 32 + ((j + r) & 31)];
        const double sig0 = sig ? sig[(int64_t)r * b.n_energy_total + b.pt_e_off[pt] + top] : sigma_rt(r, E0, lnE0);
        double rate = S0[0] * kMbToIMeV2 * sig0 / rate_E0;
        if (Emax > kEth[r]) rate += kMbToIMeV2 * v;
        pdi[(int64_t)s * ACRO_NR + r] = fmax(p.approx_zero, rate);
    }
}

// ------------------------------------------------------------------------------------------------------------
// (iv)  T-integral of the rate matrix, 9x9 matrix exponential, pre/post decay matrices, final abundances.
// ------------------------------------------------------------------------------------------------------------
// stoichiometry of the 17 reactions and the 2 decays (nucl.py:49-67,116-120,144-171): initial nuclide and final
// multiplicities over (n,p,d,t,He3,He4,Li6,Li7,Be7)
__constant__ const signed char kReacInit[ACRO_NR] = {2, 3, 3, 4, 4, 5, 5, 5, 5, 6, 6, 7, 7, 7, 8, 8, 8};
__constant__ const signed char kReacFinal[ACRO_NR][ACRO_NY] = {
    {1, 1, 0, 0, 0, 0, 0, 0, 0},   //  1 d+a>n+p
    {1, 0, 1, 0, 0, 0, 0, 0, 0},   //  2 t+a>n+d
    {2, 1, 0, 0, 0, 0, 0, 0, 0},   //  3 t+a>n+p+n
    {0, 1, 1, 0, 0, 0, 0, 0, 0},   //  4 He3+a>p+d
    {1, 2, 0, 0, 0, 0, 0, 0, 0},   //  5 He3+a>n+p+p
    {0, 1, 0, 1, 0, 0, 0, 0, 0},   //  6 He4+a>p+t
    {1, 0, 0, 0, 1, 0, 0, 0, 0},   //  7 He4+a>n+He3
    {0, 0, 2, 0, 0, 0, 0, 0, 0},   //  8 He4+a>d+d
    {1, 1, 1, 0, 0, 0, 0, 0, 0},   //  9 He4+a>n+p+d
    {1, 1, 0, 0, 0, 1, 0, 0, 0},   // 10 Li6+a>n+p+He4
    {0, 0, 0, 0, 0, 0, 0, 0, 0},   // 11 Li6+a>X
    {0, 0, 0, 1, 0, 1, 0, 0, 0},   // 12 Li7+a>t+He4
    {1, 0, 0, 0, 0, 0, 1, 0, 0},   // 13 Li7+a>n+Li6
    {2, 1, 0, 0, 0, 1, 0, 0, 0},   // 14 Li7+a>n+n+p+He4
    {0, 0, 0, 0, 1, 1, 0, 0, 0},   // 15 Be7+a>He3+He4
    {0, 1, 0, 0, 0, 0, 1, 0, 0},   // 16 Be7+a>p+Li6
    {1, 2, 0, 0, 0, 1, 0, 0, 0}};  // 17 Be7+a>p+p+n+He4

// the entry-wise combinations of the Pade-13 numerator/denominator (explicit fma: no dependence on what the compiler contracts)
__device__ __forceinline__ double pade_w1(const double* b, double a2, double a4, double a6, int top) {
    return fma(b[top - 4], a2, fma(b[top - 2], a4, b[top] * a6));                      // b_top A6 + b_(top-2) A4 + b_(top-4) A2
}
__device__ __forceinline__ double pade_w2(const double* b, double v, double a2, double a4, double a6, int top, bool diag) {
    return fma(b[top - 4], a2, fma(b[top - 2], a4, fma(b[top], a6, v))) + (diag ? b[top - 6] : 0.0);   // v + ... + b_(top-6) I
}
__device__ __forceinline__ double pade_v2(const double* b, double v, double a2, double a4, double a6, int top, bool diag) {
    return v + (fma(b[top - 4], a2, fma(b[top - 2], a4, b[top] * a6)) + (diag ? b[top - 6] : 0.0));
}

__constant__ const double kPade13[14] = {64764752532480000., 32382376266240000., 7771770303897600., 1187353796428800., 129060195264000.,
                                         10559470521600., 670442572800., 33522128640., 1323241920., 40840800., 960960., 16380., 182., 1.};

// ---- 9x9 matrix exponential, Pade-13 scaling and squaring (Higham 2005, the algorithm behind scipy.linalg.expm used at
// models.py:130), by one warp on matrices in shared memory: lane e (+32, +64) owns entry e of every 9x9 product.  Rounds
// entry for entry like the one-thread form it replaced (rounds 1-2: 255 registers, 5.6 kB of local memory, 130 k dependent
// instructions of ONE lane per point -- 0.2 ms of a single point's 0.67 ms device latency).  matrix_kernel and transfer_kernel
// share it: a point's abundances are the same bits on both routes (tests: test_matrix_stage_equals_transfer_kernel).
__device__ __forceinline__ void mm9w(const double* A, const double* B, double* C, int lane) {
    for (int e = lane; e < 81; e += 32) {
        const int i = e / 9, j = e - 9 * i;
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < 9; ++k) s = fma(A[i * 9 + k], B[k * 9 + j], s);
        C[e] = s;
    }
    __syncwarp();
}

// ws: 7 x 81 doubles of shared memory; Ain, R: shared memory as well (R may not alias Ain)
__device__ void expm9_warp(const double* Ain, double* R, double* ws, int lane) {
    const double* b = kPade13;
    double *A = ws, *A2 = ws + 81, *A4 = ws + 162, *A6 = ws + 243, *U = ws + 324, *V = ws + 405, *W = ws + 486;
    double nrm = 0.0;
    for (int j = 0; j < 9; ++j) {        // every lane: the same loop as the serial form (81 loads; keeps the order of the sums)
        double c = 0.0;
        for (int i = 0; i < 9; ++i) c += fabs(Ain[i * 9 + j]);
        nrm = fmax(nrm, c);
    }
    int s = 0;
    if (nrm > 5.371920351148152) s = (int)fmax(0.0, ceil(log2(nrm / 5.371920351148152)));
    const double sc = ldexp(1.0, -s);
    for (int e = lane; e < 81; e += 32) A[e] = Ain[e] * sc;
    __syncwarp();
    mm9w(A, A, A2, lane); mm9w(A2, A2, A4, lane); mm9w(A4, A2, A6, lane);
    for (int e = lane; e < 81; e += 32) W[e] = pade_w1(b, A2[e], A4[e], A6[e], 13);
    __syncwarp();
    mm9w(A6, W, V, lane);
    for (int e = lane; e < 81; e += 32) W[e] = pade_w2(b, V[e], A2[e], A4[e], A6[e], 7, e % 10 == 0);
    __syncwarp();
    mm9w(A, W, U, lane);
    for (int e = lane; e < 81; e += 32) W[e] = pade_w1(b, A2[e], A4[e], A6[e], 12);
    __syncwarp();
    mm9w(A6, W, V, lane);
    for (int e = lane; e < 81; e += 32) V[e] = pade_v2(b, V[e], A2[e], A4[e], A6[e], 6, e % 10 == 0);
    __syncwarp();
    for (int e = lane; e < 81; e += 32) { W[e] = V[e] - U[e]; R[e] = V[e] + U[e]; }
    __syncwarp();
    // LU with partial pivoting: lanes 0..8 own column k of W, lanes 9..17 column k of R (18 independent column updates per step)
    for (int c = 0; c < 9; ++c) {
        int piv = c;
        double best = fabs(W[c * 9 + c]);
        for (int r = c + 1; r < 9; ++r) { const double v = fabs(W[r * 9 + c]); if (v > best) { best = v; piv = r; } }
        const double ip = 1.0 / W[piv * 9 + c];
        double f[9];
#pragma unroll
        for (int r = 0; r < 9; ++r) f[r] = 0.0;
        // multipliers from the un-swapped matrix: row piv plays row c, row c plays row piv
#pragma unroll
        for (int r = 0; r < 9; ++r)
            if (r > c) f[r] = W[(r == piv ? c : r) * 9 + c] * ip;
        __syncwarp();
        if (lane < 18) {
            double* M = lane < 9 ? W : R;
            const int k = lane < 9 ? lane : lane - 9;
            if (piv != c) { const double t = M[c * 9 + k]; M[c * 9 + k] = M[piv * 9 + k]; M[piv * 9 + k] = t; }
            if (lane >= 9 || k >= c) {
                const double top = M[c * 9 + k];
#pragma unroll
                for (int r = 0; r < 9; ++r)
                    if (r > c && f[r] != 0.0) M[r * 9 + k] = fma(-f[r], top, M[r * 9 + k]);
            }
        }
        __syncwarp();
    }
    if (lane < 9) {
        const int k = lane;
        for (int c = 8; c >= 0; --c) {
            const double ip = 1.0 / W[c * 9 + c];
            double v = R[c * 9 + k];
            for (int cc = c + 1; cc < 9; ++cc) v = fma(-W[c * 9 + cc], R[cc * 9 + k], v);
            R[c * 9 + k] = v * ip;
        }
    }
    __syncwarp();
    for (int q = 0; q < s; ++q) {
        mm9w(R, R, W, lane);
        for (int e = lane; e < 81; e += 32) R[e] = W[e];
        __syncwarp();
    }
}

// Yf = postd . expm(A) . pred . Y0 for the three abundance columns (models.py:77-89,133-157), by one warp; A and the scratch ws
// (9 x 81 doubles) in shared memory.  Lane 0 returns the flag.  expm_out (global, optional): the plain exponential of A.
__device__ void apply_transfer_warp(const double* A, double t_at_tmax, const double* Y0, double* Yf /*global [9][3]*/, int* nonfinite,
                                    double* ws, int lane, double* expm_out = nullptr) {
    double* B = ws + 7 * 81;
    double* R = ws + 8 * 81;
    if (expm_out) {            // the plain exponential, for callers that ask for it (models.py:130)
        expm9_warp(A, R, ws, lane);
        for (int e = lane; e < 81; e += 32) expm_out[e] = R[e];
        __syncwarp();
    }
    // For the abundances the free-neutron decay is folded out of the matrix first.  Neutrons (row/column 0) are never a
    // target: their column is pure decay into protons and the proton column is empty; the pre-decay below empties the
    // neutron entry and the post-decay adds whatever neutrons remain to the protons.  So Y_n + Y_p and every other nuclide
    // do not depend on the neutron lifetime at all, and  exp(A) -> exp(B)  with  B = A, column n := 0, row p += row n,
    // row n := 0  gives the SAME transfer -- exactly, for any rate -- while ||B||_1 no longer contains rate x time of the
    // neutron (1.45e9 at tau = 7.5e8 s, ||B||_1 = 2.3e3).  Scaling and squaring doubles the rounding error of the Pade
    // solve with every squaring, 29 of them for ||A||_1 = 1.45e9: nucleon number drifted by 7.7e-8 (SciPy: 3e-9 on the same
    // matrix); with the fold the abundances agree with a 60-digit evaluation to 6e-14 (tools/proto/expm_fold_study.py).
    bool foldable = (A[0 * 9 + 0] + A[1 * 9 + 0] == 0.0);
    for (int i = 2; i < 9; ++i) foldable = foldable && A[i * 9 + 0] == 0.0;
    for (int i = 0; i < 9; ++i) foldable = foldable && A[i * 9 + 1] == 0.0;
    for (int e = lane; e < 81; e += 32) {
        const int i = e / 9, k = e - 9 * i;
        double v = A[e];
        if (foldable) {
            if (k == 0 || i == 0) v = 0.0;
            else if (i == 1) v = A[9 + k] + A[k];
        }
        B[e] = v;
    }
    __syncwarp();
    expm9_warp(B, R, ws, lane);
    const double ef = exp(-t_at_tmax / kTauT);
    bool bad = false;
    if (lane < 3) {
        const int c = lane;
        double y[9], z[9];
        for (int i = 0; i < 9; ++i) y[i] = Y0[i * 3 + c];
        y[1] += y[0]; y[0] = 0.0;
        y[4] += (1.0 - ef) * y[3]; y[3] = ef * y[3];
        for (int i = 0; i < 9; ++i) {
            double v = 0.0;
            for (int k = 0; k < 9; ++k) v = fma(R[i * 9 + k], y[k], v);
            z[i] = v;
        }
        z[1] += z[0]; z[0] = 0.0;
        z[4] += z[3]; z[3] = 0.0;
        z[7] += z[8]; z[8] = 0.0;
        for (int i = 0; i < 9; ++i) { Yf[i * 3 + c] = z[i]; bad |= !isfinite(z[i]); }
    }
    if (__any_sync(0xffffffffu, bad) && lane == 0) *nonfinite = 1;
}

// int_{x1}^{x2} 10^{alpha x + beta} ln10 dx, stable for any alpha (evaluated from the larger end)
__device__ __forceinline__ double pow10_segment(double alpha, double beta, double x1, double x2) {
    const double ln10 = 2.302585092994046;
    const double dx = x2 - x1;
    const double z = alpha * ln10 * dx;
    if (z >= 0.0) {
        const double top = exp((alpha * x2 + beta) * ln10);
        const double phi = (z > 1e-8) ? -expm1(-z) / z : 1.0 - 0.5 * z;
        return ln10 * top * dx * phi;
    }
    const double bot = exp((alpha * x1 + beta) * ln10);
    const double phi = (z < -1e-8) ? expm1(z) / z : 1.0 + 0.5 * z;
    return ln10 * bot * dx * phi;
}

// One block of eight warps per point (a single point would otherwise spend 0.7 ms in ONE warp walking its ~2000 cosmo-table
// cells: half of the device latency of a run_disintegration() call).
// Integrates Gamma_r(T) T / |dT/dt| dlogT exactly: both factors are piecewise power laws (log10-log10 linear on the NT-point
// rate grid, utils.py:38-61, and on the cosmo table, input.py:205-226), so each sub-cell has the closed form of
// pow10_segment; no quadrature (the reference uses QAGS, nucl.py:613-614).
__global__ void __launch_bounds__(256) matrix_kernel(int n_points, const int32_t* __restrict__ pt_nt,
                                                     const int32_t* __restrict__ pt_sys_off, const double* __restrict__ sys_T,
                                                     const double* __restrict__ pdi, const double* __restrict__ T_lo,
                                                     const double* __restrict__ pt_t_at_tmax, const int* __restrict__ status_sys,
                                                     const double* __restrict__ pt_e0, double* __restrict__ mpdi_out,
                                                     double* __restrict__ mdcy_out, double* __restrict__ Yf_out,
                                                     int32_t* __restrict__ status_out, DeviceTables t, acro_params_t p) {
    extern __shared__ double sm[];
    __shared__ double s_part[8][ACRO_NR + 1];
    __shared__ int s_st[8];
    const int pt = blockIdx.x, tid = threadIdx.x, nth = blockDim.x, lane = tid & 31, warp = tid >> 5;
    if (pt >= n_points) return;
    const int nt = pt_nt[pt], s0 = pt_sys_off[pt];
    double* X = sm;             // [nt] log10 T
    double* Y = sm + nt;        // [17][nt] log10 rate
    for (int k = tid; k < nt; k += nth) {
        X[k] = log10(sys_T[s0 + k]);
        for (int r = 0; r < ACRO_NR; ++r) Y[r * nt + k] = log10(pdi[(int64_t)(s0 + k) * ACRO_NR + r]);
    }
    __syncthreads();
    int st = 0;
    if (status_sys)
        for (int k = tid; k < nt; k += nth) st |= status_sys[s0 + k] ? ACRO_ST_DIRECT_RATE : 0;
    const double x_hi = X[nt - 1];
    double x_lo = X[0];
    if (T_lo) x_lo = fmin(fmax(log10(T_lo[pt]), X[0]), x_hi);
    const double* LT = t.cosmo_lT;      // descending
    const double* LD = t.cosmo_ldTdt;
    const int nc = t.n_cosmo;
    if (!(LT[nc - 1] <= x_lo && x_hi <= LT[0])) st |= ACRO_ST_T_OUTSIDE;
    // cosmo cell c spans [LT[c+1], LT[c]]; first cell containing x_hi and last cell containing x_lo
    int lo = 0, hi = nc - 2;
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (LT[mid + 1] < x_hi) hi = mid; else lo = mid + 1; }
    const int c_first = lo;
    lo = 0; hi = nc - 2;
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (LT[mid + 1] <= x_lo) hi = mid; else lo = mid + 1; }
    const int c_last = lo;
    double acc[ACRO_NR + 1];
#pragma unroll
    for (int r = 0; r <= ACRO_NR; ++r) acc[r] = 0.0;
    const double inv_dx = (double)(nt - 1) / (X[nt - 1] - X[0]);
    for (int c = c_first + tid; c <= c_last; c += nth) {
        const double ca = fmax(LT[c + 1], x_lo), cb = fmin(LT[c], x_hi);
        if (!(cb > ca)) continue;
        const double mc = (LD[c + 1] - LD[c]) / (LT[c + 1] - LT[c]);
        const double bc = LD[c] - mc * LT[c];
        int k0 = (int)((ca - X[0]) * inv_dx);
        k0 = max(0, min(k0, nt - 2));
        while (k0 > 0 && X[k0] > ca) --k0;
        while (k0 < nt - 2 && X[k0 + 1] <= ca) ++k0;
        for (int k = k0; k < nt - 1 && X[k] < cb; ++k) {
            const double x1 = fmax(ca, X[k]), x2 = fmin(cb, X[k + 1]);
            if (!(x2 > x1)) continue;
            acc[ACRO_NR] += pow10_segment(1.0 - mc, -bc, x1, x2);
            const double idx = 1.0 / (X[k + 1] - X[k]);
#pragma unroll
            for (int r = 0; r < ACRO_NR; ++r) {
                const double mr = (Y[r * nt + k + 1] - Y[r * nt + k]) * idx;
                const double br = Y[r * nt + k + 1] - mr * X[k + 1];
                acc[r] += pow10_segment(mr + 1.0 - mc, br - bc, x1, x2);
            }
        }
    }
#pragma unroll
    for (int r = 0; r <= ACRO_NR; ++r) acc[r] = warp_sum(acc[r]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) st |= __shfl_xor_sync(0xffffffffu, st, o);
    if (nth > 32) {          // combine the warps in a fixed order (the result does not depend on scheduling)
        if (lane == 0) {
#pragma unroll
            for (int r = 0; r <= ACRO_NR; ++r) s_part[warp][r] = acc[r];
            s_st[warp] = st;
        }
        __syncthreads();
        if (tid == 0)
            for (int w = 1; w < (nth >> 5); ++w) {
#pragma unroll
                for (int r = 0; r <= ACRO_NR; ++r) acc[r] += s_part[w][r];
                st |= s_st[w];
            }
    }
    if (warp != 0) return;
    // warp 0: lane 0 holds the 18 integrals; rate matrices by lane 0 (a few hundred operations), the exponential and the
    // abundances by the whole warp on matrices in shared memory
    __shared__ double s_mx[12 * 81];
    double* A = s_mx;
    double* Mp = s_mx + 81;
    double* Md = s_mx + 162;
    if (lane == 0) {
        for (int i = 0; i < 81; ++i) { Mp[i] = 0.0; Md[i] = 0.0; }
        // mpdi[i][j] = sum_r pref_r(i,j) I_r with I_r = int_{Tmin}^{Tmax} Gamma_r/|dT/dt| dT > 0: the reference integrates
        // pref*Gamma/(dT/dt) from log Tmax DOWN to log Tmin with dT/dt < 0, which is the same number (nucl.py:563-570,613)
        for (int r = 0; r < ACRO_NR; ++r) {
            const int j = kReacInit[r];
            bool in_final = false;
            for (int i = 0; i < 9; ++i)
                if (kReacFinal[r][i] != 0) { Mp[i * 9 + j] += kReacFinal[r][i] * acc[r]; if (i == j) in_final = true; }
            if (!in_final) Mp[j * 9 + j] += -acc[r];
        }
        const double In = (kHbar / kTauN) * acc[ACRO_NR], It = (kHbar / kTauT) * acc[ACRO_NR];
        Md[0 * 9 + 0] = -In; Md[1 * 9 + 0] = In;    // n -> p
        Md[3 * 9 + 3] = -It; Md[4 * 9 + 3] = It;    // t -> He3
    }
    __syncwarp();
    for (int i = lane; i < 81; i += 32) {
        A[i] = Mp[i] + Md[i];
        if (mpdi_out) mpdi_out[(int64_t)pt * 81 + i] = Mp[i];
        if (mdcy_out) mdcy_out[(int64_t)pt * 81 + i] = Md[i];
    }
    __syncwarp();
    int bad = 0;
    if (Yf_out) apply_transfer_warp(A, pt_t_at_tmax[pt], t.Y0, Yf_out + (int64_t)pt * 27, &bad, s_mx + 3 * 81, lane);
    if (lane != 0) return;
    if (bad) st |= ACRO_ST_NONFINITE;
    if (pt_e0 && (double)(int)pt_e0[pt] > 1e3) st |= ACRO_ST_E0_ABOVE_GEV;
    if (status_out) status_out[pt] = st;
}

// Fast-parameter path: one warp per job (a rescaling of a buffered matrix pair), four jobs per block.
__global__ void __launch_bounds__(128) transfer_kernel(int n, const double* __restrict__ mpdi, const double* __restrict__ mdcy,
                                                       const int32_t* __restrict__ matrix_index,
                                                       const double* __restrict__ factor, const double* __restrict__ t_at_tmax,
                                                       double* __restrict__ expm_out, double* __restrict__ Yf_out,
                                                       int32_t* __restrict__ status, DeviceTables t) {
    __shared__ double s_mx[4][10 * 81];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int i = blockIdx.x * 4 + warp;
    if (i >= n) return;
    double* A = s_mx[warp];
    const double f = factor ? factor[i] : 1.0;
    const int64_t m = matrix_index ? matrix_index[i] : i;       // many rescalings of one (mpdi, mdcy) pair share its storage
    for (int k = lane; k < 81; k += 32) A[k] = f * mpdi[m * 81 + k] + mdcy[m * 81 + k];
    __syncwarp();
    int bad = 0;
    apply_transfer_warp(A, t_at_tmax[i], t.Y0, Yf_out + (int64_t)i * 27, &bad, A + 81, lane, expm_out ? expm_out + (int64_t)i * 81 : nullptr);
    if (lane == 0 && status) status[i] = bad ? ACRO_ST_NONFINITE : 0;
}

// total rates for arbitrary (E,T) pairs -> out[n][3]
__global__ void total_rates_kernel(int n, const double* __restrict__ E, const double* __restrict__ T, double* __restrict__ out,
                                   DeviceTables t, acro_params_t p, DirectRateItem* __restrict__ items, int* __restrict__ item_count) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const RateEval r = eval_rates(E[i], T[i], t, p);
    out[3 * (int64_t)i] = r.g_ph; out[3 * (int64_t)i + 1] = r.g_el; out[3 * (int64_t)i + 2] = r.g_el;
    if (r.direct) {
        const int k = atomicAdd(item_count, 1);
        DirectRateItem it;
        it.g_index = 3 * (int64_t)i; it.ne = 1; it.kind = r.direct; it.E = E[i]; it.T = T[i];
        items[k] = it;
    }
}

// the two tabulated rates as direct 2-D integrals for arbitrary (E,T) pairs -> out[n][3] = {pair creation, IC, IC}: what a
// rates.db entry holds before log10 (tools/create_db.py:34-40 of the reference; the table generator tools/make_rates_db.py)
__global__ void direct_rate_items_kernel(int n, const double* __restrict__ E, const double* __restrict__ T, double* __restrict__ out,
                                         DirectRateItem* __restrict__ items, int* __restrict__ item_count) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    out[3 * (int64_t)i] = 0.0; out[3 * (int64_t)i + 1] = 0.0; out[3 * (int64_t)i + 2] = 0.0;
    DirectRateItem it;
    it.g_index = 3 * (int64_t)i; it.ne = 1; it.E = E[i]; it.T = T[i];
    it.kind = 2 | (!(E[i] < kMe2 / (50.0 * T[i])) ? 1 : 0);        // pair-creation threshold (cascade.py:340)
    items[atomicAdd(item_count, 1)] = it;
}

// DFMA throughput micro-benchmark: 8 independent chains per thread
__global__ void __launch_bounds__(256) dfma_kernel(double* out, int iters, double a, double b) {
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; ++i) {
        x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
        x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

// ------------------------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------------------------
static int sm_count() {
    int n = 0, dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    return n > 0 ? n : 148;
}

cudaError_t launch_rates(const acro_batch_t& b, const acro_out_t& o, const DeviceTables& t, const acro_params_t& p,
                         DirectRateItem* items, int* item_count, int* status_sys, cudaStream_t st, LaunchCounters& lc) {
    if (b.n_sys_energy_total == 0) return cudaSuccess;
    const int64_t blocks = (b.n_sys_energy_total + 255) / 256;
    rates_kernel<<<(unsigned)blocks, 256, 0, st>>>(b, o.G, t, p, items, item_count, status_sys);
    lc.launches++;
    return cudaGetLastError();
}

cudaError_t launch_direct_rates(const DeviceTables& t, const acro_params_t& p, double* G, const DirectRateItem* items,
                                const int* item_count, int64_t max_items, cudaStream_t st, LaunchCounters& lc) {
    (void)t;
    if (max_items == 0) return cudaSuccess;
    int64_t blocks = (max_items + 3) / 4;                       // 4 warps per block, grid-stride over the list
    const int64_t cap = (int64_t)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    direct_rates_kernel<<<(unsigned)blocks, 128, 0, st>>>(G, items, item_count, p);
    lc.launches++;
    return cudaGetLastError();
}

cudaError_t launch_kernels(const acro_batch_t& b, const DeviceTables& t, const acro_params_t& p, double* Kws, double* closed,
                           cudaStream_t st, LaunchCounters& lc) {
    lc.split_valid = false;
    if (b.n_pairs_total == 0) return cudaSuccess;
    if (p.kernel_mode == ACRO_KERNELS_FAST) {   // the production kernel builder: all five planes
        const int64_t tri_max = (int64_t)b.ne_max * (b.ne_max + 1) / 2;
        const size_t smem = sizeof(double) * (3 * (size_t)kMaxPan * kPanN + (size_t)kMaxPan * kPanN * kTT + (kTileScalars + 1) * kTT +
                                              2 * kLogTab + kPanN * kPanN + (size_t)kMmWarps * MM_NCONST * 32);
        const int64_t passes_max = (tri_max + kMmPairs - 1) / kMmPairs;
        // passes (32 pairs per warp) per block and Bose-table fill: up to kMmPassesPerBlock for scan batches (measured, 512
        // threads, ms per 1000-point slice with three integrals: 1: 28.5, 2: 24.8, 4: 22.9, 8: 19.8, 12: 19.6; with two: 8: 13.6, 12: 13.5,
        // 16: 13.4, 24: 13.1, 32: 12.8, 48: 12.7 -- a block drains once, when its slowest warp is done), fewer while that would leave less
        // than two blocks per SM (single points and short batches: latency, not throughput)
        const int ppb = (int)std::max<int64_t>(1, std::min<int64_t>(kMmPassesPerBlock, (int64_t)b.n_points * passes_max / (2 * (int64_t)sm_count())));
        dim3 grid((unsigned)((passes_max + ppb - 1) / ppb), (unsigned)b.n_points);
        // function attributes are per device (each primary context has its own CUfunction): set before every launch
        cudaError_t e = cudaSuccess;
        // function attributes are per device (each primary context has its own CUfunction): set before every launch
        if ((e = cudaFuncSetAttribute(kernels_mma_kernel<KIND_IC_ELECTRON>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
        if ((e = cudaFuncSetAttribute(kernels_mma_kernel<KIND_PC_ELECTRON>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
        dim3 gp((unsigned)((tri_max + 127) / 128), (unsigned)b.n_points);
        pair_closed_forms_kernel<<<gp, 128, 0, st>>>(b, closed);
        lc.launches++;
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
        if (lc.split[0]) cudaEventRecord(lc.split[0], st);
        // plane 1 (e -> gamma) in closed form, complete; the shared-grid and Wien kernels do the other two integrals
        ic_photon_closed_kernel<<<dim3((unsigned)((tri_max + kIcpThreads - 1) / kIcpThreads), (unsigned)b.n_points), kIcpThreads, 0, st>>>(b, Kws, p);
        lc.launches++;
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
        if (lc.split[1]) cudaEventRecord(lc.split[1], st);
        kernels_mma_kernel<KIND_IC_ELECTRON><<<grid, kMmThreads, smem, st>>>(b, Kws, closed, t, p, ppb);
        kernels_mma_kernel<KIND_PC_ELECTRON><<<grid, kMmThreads, smem, st>>>(b, Kws, closed, t, p, ppb);
        lc.launches += 2;
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
        if (lc.split[2]) lc.split_valid = cudaEventRecord(lc.split[2], st) == cudaSuccess;
        kernels_wien_kernel<64, KIND_IC_ELECTRON><<<gp, 128, 0, st>>>(b, Kws, closed, t, p);
        kernels_wien_kernel<64, KIND_PC_ELECTRON><<<gp, 128, 0, st>>>(b, Kws, closed, t, p);
        lc.launches += 2;
        return cudaGetLastError();
    }
#ifdef ACRO_CROSSCHECK
    const unsigned g = (unsigned)((b.n_pairs_total + 127) / 128);
    if (p.kernel_mode == ACRO_KERNELS_PLAIN_GL) {
        switch (p.quad_order) {
            case 16: kernels_kernel<16, false><<<g, 128, 0, st>>>(b, Kws, t, p); break;
            case 32: kernels_kernel<32, false><<<g, 128, 0, st>>>(b, Kws, t, p); break;
            case 96: kernels_kernel<96, false><<<g, 128, 0, st>>>(b, Kws, t, p); break;
            default: kernels_kernel<64, false><<<g, 128, 0, st>>>(b, Kws, t, p); break;
        }
    } else {
        switch (p.quad_order) {
            case 96: kernels_kernel<96, true><<<g, 128, 0, st>>>(b, Kws, t, p); break;
            default: kernels_kernel<64, true><<<g, 128, 0, st>>>(b, Kws, t, p); break;
        }
    }
    lc.launches++;
    return cudaGetLastError();
#else
    return cudaErrorNotSupported;   // cross-check modes exist only in the test build (-DACRO_CROSSCHECK)
#endif
}

cudaError_t launch_count_work(const acro_batch_t& b, const acro_params_t& p, unsigned long long* out, cudaStream_t st,
                              LaunchCounters& lc) {
    if (b.n_pairs_total == 0) return cudaSuccess;
    const int64_t blocks = (b.n_pairs_total + 127) / 128;
    count_work_kernel<<<(unsigned)blocks, 128, 0, st>>>(b, out, p);
    lc.launches++;
    return cudaGetLastError();
}

cudaError_t launch_solve_pdi(const acro_batch_t& b, const acro_out_t& o, const DeviceTables& t, const acro_params_t& p,
                             const double* Kws, double* sig, cudaStream_t st, LaunchCounters& lc) {
    (void)t;
    if (b.n_systems == 0) return cudaSuccess;
    const int stride = b.ne_max + max(3 * b.ne_max, b.ne_max + ACRO_NR * 32);   // doubles per warp, see the kernel
    int wpb = 4;
    while (wpb > 1 && (size_t)wpb * stride * sizeof(double) > 200 * 1024) wpb >>= 1;
    const size_t smem = (size_t)wpb * stride * sizeof(double);
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(solve_pdi_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    const int blocks = (b.n_systems + wpb - 1) / wpb;
    if (sig && o.pdi && b.ne_max > 1) {
        dim3 gs((unsigned)((b.ne_max + 127) / 128), (unsigned)b.n_points);
        pdi_sigma_kernel<<<gs, 128, 0, st>>>(b, sig, p);
        lc.launches++;
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return e;
    }
    solve_pdi_kernel<<<blocks, wpb * 32, smem, st>>>(b, o.G, o.F, o.pdi, Kws, (o.pdi ? sig : nullptr), p, stride);
    lc.launches++;
    return cudaGetLastError();
}

cudaError_t launch_matrix(int n_points, const int32_t* pt_nt, const int32_t* pt_sys_off, const double* sys_T,
                          const double* pdi, const double* T_lo, const double* pt_t_at_tmax, const int* status_sys,
                          const double* pt_e0, int nt_max, double* mpdi, double* mdcy, double* Yf, int32_t* status,
                          const DeviceTables& t, const acro_params_t& p, cudaStream_t st, LaunchCounters& lc) {
    if (n_points == 0) return cudaSuccess;
    const size_t smem = (size_t)(ACRO_NR + 1) * nt_max * sizeof(double);
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(matrix_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    // eight warps share a point's ~2000 cosmo-table cells, whatever the batch size (one warp per point for large batches was
    // 0.83 against 0.43 ms per 1000-point slice once the exponential no longer ran in one lane); the partition of the cells and
    // with it the rounding of the 18 integrals is then the same for a single point and for a scan
    const int threads = 256;
    matrix_kernel<<<n_points, threads, smem, st>>>(n_points, pt_nt, pt_sys_off, sys_T, pdi, T_lo, pt_t_at_tmax, status_sys, pt_e0,
                                                   mpdi, mdcy, Yf, status, t, p);
    lc.launches++;
    return cudaGetLastError();
}

cudaError_t launch_transfer(int n, const double* mpdi, const double* mdcy, const int32_t* matrix_index, const double* factor, const double* t_at_tmax,
                            double* expm_out, double* Yf, int32_t* status, const DeviceTables& t, cudaStream_t st,
                            LaunchCounters& lc) {
    if (n == 0) return cudaSuccess;
    transfer_kernel<<<(n + 3) / 4, 128, 0, st>>>(n, mpdi, mdcy, matrix_index, factor, t_at_tmax, expm_out, Yf, status, t);
    lc.launches++;
    return cudaGetLastError();
}

cudaError_t launch_total_rates(int n, const double* E, const double* T, double* out, const DeviceTables& t,
                               const acro_params_t& p, DirectRateItem* items, int* item_count, cudaStream_t st,
                               LaunchCounters& lc) {
    if (n == 0) return cudaSuccess;
    total_rates_kernel<<<(n + 127) / 128, 128, 0, st>>>(n, E, T, out, t, p, items, item_count);
    lc.launches++;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    return launch_direct_rates(t, p, out, items, item_count, n, st, lc);
}

cudaError_t launch_direct_rate_table(int n, const double* E, const double* T, double* out, const DeviceTables& t,
                                     const acro_params_t& p, DirectRateItem* items, int* item_count, cudaStream_t st,
                                     LaunchCounters& lc) {
    if (n == 0) return cudaSuccess;
    direct_rate_items_kernel<<<(n + 127) / 128, 128, 0, st>>>(n, E, T, out, items, item_count);
    lc.launches++;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    return launch_direct_rates(t, p, out, items, item_count, n, st, lc);
}

cudaError_t launch_unpack_kernels(const double* Kws, int64_t n_pairs_total, int64_t k_off, int ne, double* K_dense,
                                  cudaStream_t st, LaunchCounters& lc) {
    const int n = 5 * ne * ne;
    unpack_kernels_kernel<<<(n + 255) / 256, 256, 0, st>>>(Kws, n_pairs_total, k_off, ne, K_dense);
    lc.launches++;
    return cudaGetLastError();
}

cudaError_t run_dfma_benchmark(double* tflops, LaunchCounters& lc) {
    const int blocks = sm_count() * 8, threads = 256, iters = 1 << 16;
    double* d = nullptr;
    cudaError_t e = cudaMalloc(&d, sizeof(double) * blocks * threads);
    if (e != cudaSuccess) return e;
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(a);
        dfma_kernel<<<blocks, threads>>>(d, iters, 0.999999, 1e-9);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms = 0;
        cudaEventElapsedTime(&ms, a, b);
        if (rep > 0 && ms < best) best = ms;
        lc.launches++;
    }
    e = cudaGetLastError();
    cudaEventDestroy(a); cudaEventDestroy(b);
    cudaFree(d);
    *tflops = 2.0 * 8.0 * (double)iters * blocks * threads / (best * 1e-3) / 1e12;
    return e;
}

}  // namespace acro
