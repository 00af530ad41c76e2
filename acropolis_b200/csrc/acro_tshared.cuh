// acro_tshared.cuh -- the production kernel builder (kernel_mode = ACRO_KERNELS_FAST): temperature-shared evaluation of
// the three thermal kernel integrals plus the per-temperature Wien-tail kernel (kernels_wien_kernel).  This file holds the
// shared grid, the integration limits, the partial-panel machinery, the Wien kernel and kernels_shared_kernel, the
// one-thread-per-pair form of the shared kernel (build-time alternative, -DACRO_TS_LANES4=0); the form that ships,
// kernels_shared4_kernel with four lanes per pair, is in acro_tshared4.cuh and uses the same grid, limits and tables.
//
// In   K(E,E',T) ~ int f(E,E',x) * x^p / (e^{x/T} - 1) dlog x   the factor f (the functions F, G of cascade.py:29-80) does
// not depend on T and the Bose factor does not depend on (E,E').  For all temperatures of one model point whose lower
// limit satisfies x_a <= wien_from * T (acro_params_t.wien_from, e..e^2) the integral is therefore evaluated on ONE panel
// grid in y = log x that is shared by every pair and every temperature of the point:
//
//   * panels of 12 Gauss-Legendre nodes: 1 e-fold wide from y_top = log(min(50 Tmax, E0)) down to log(Tmin) - 1 (where
//     some temperature still sees the Planck roll-off), 2 e-folds wide below, down to log(Tmin) - 25 (what lies below
//     contributes < 2e-11; what lies above 50 Tmax < 3e-19);
//   * per temperature tile the Bose factors at the grid nodes are staged in SHARED MEMORY once per thread block and serve
//     kTilesPerBlock tiles of 256 pairs -- the "thermal-photon grid" of BASELINE.json:north_star;
//   * each thread owns one (E,E') pair: f is evaluated once per node and accumulated into kTT temperature sums with one
//     FMA each; the panel that contains the lower limit is integrated on its own nodes against the Legendre expansion of
//     the Bose factor (exact for what the panel rule itself resolves), and so is the panel that contains the upper limit
//     when that is the kinematic one (x_b0 < 200 T); the thermal cut-off 200 T is applied as zeros in the Bose table.
//
// The shared kernel writes all five planes (closed forms included); kernels_wien_kernel then overwrites / adds the
// (pair, T) combinations beyond wien_from * T with its own Gauss-Laguerre rule (acro_fastquad.cuh).  Both evaluate the
// regime predicates with the same expressions on the same numbers.
//
// Tuning constants (each is a -D macro so that tools/gpu_ab_variants.py can time a variant library; values = measured best
// on the 840-point benchmark slice):
//   (measured on kernels_shared_kernel; acro_tshared4.cuh inherits TT, PANN, HC, MAXPAN, TOPX and TILES)
//   ACRO_TS_TT      40   temperatures per tile = accumulators per thread.  One 256-thread block per SM (248 registers) beats
//                        20 per tile with two blocks by 18 %: f is evaluated once instead of twice for NT = 39/40.
//   ACRO_TS_PANN    12   nodes per panel (16: same answers to 5e-10, +14 % time; 10: 1.9e-8 and -15 %; 8: 1e-6, rejected)
//   ACRO_TS_HC      2.0  width of the coarse panels [e-folds]
//   ACRO_TS_MAXPAN  28   panels per point (12 fine + 12 coarse for a 2-decade T range; 17 + 11 for 4 decades)
//   ACRO_TS_TOPX    50   top of the grid in units of Tmax (200 = the reference's cut-off: same results, +9 % time)
//   ACRO_TS_FISSION 2    node values evaluated together in the panel loop (1: +7 % time; 3 and more: the loop body outgrows
//                        the 6 KB L0 instruction cache and loses again)
//   ACRO_TS_TILES   4    tiles of 256 pairs per block and Bose-table fill (1: +20 %; 8 and more: tail of unequal blocks)
//   ACRO_TS_THREADS 256  threads per block (320 at 168 registers: spills, +13 %)
#pragma once
#include "acro_fastquad.cuh"

namespace acro {

#ifndef ACRO_TS_TT
#define ACRO_TS_TT 40
#endif
#ifndef ACRO_TS_PANN
#define ACRO_TS_PANN 12
#endif
#ifndef ACRO_TS_HC
#define ACRO_TS_HC 2.0
#endif
#ifndef ACRO_TS_MAXPAN
#define ACRO_TS_MAXPAN 28
#endif
#ifndef ACRO_TS_TOPX
#define ACRO_TS_TOPX 50.0
#endif
#ifndef ACRO_TS_FISSION
#define ACRO_TS_FISSION 2
#endif
#ifndef ACRO_TS_TILES
#define ACRO_TS_TILES 4
#endif
#ifndef ACRO_TS_THREADS
#define ACRO_TS_THREADS 256
#endif
constexpr int kTT = ACRO_TS_TT;
constexpr int kPanN = ACRO_TS_PANN;
constexpr double kHC = ACRO_TS_HC;
constexpr int kMaxPan = ACRO_TS_MAXPAN;
constexpr double kTopX = ACRO_TS_TOPX;
constexpr int kSharedThreads = ACRO_TS_THREADS;
constexpr int kTilesPerBlock = ACRO_TS_TILES;
constexpr int kTileScalars = 8;   // per-temperature scalars of a tile kept in shared memory (kernels_shared_kernel)

__constant__ double c_leg_modal[kPanN][kPanN];   // (2j+1)/2 * w_m * P_j(x_m): Legendre coefficients from nodal values
__constant__ double c_leg_rec[kPanN][2];         // recurrence P_{j+1} = a_j t P_j - b_j P_{j-1}: a_j = (2j+1)/(j+1), b_j = j/(j+1)

struct SharedGrid {
    double y_top;     // top of the grid
    int n_fine, n_pan;
    double y_bot;     // bottom of the grid
};

// the grid of one model point (uniform over the block and identical in both kernels)
__device__ __forceinline__ SharedGrid shared_grid(double Tmin, double Tmax, double E_top, double Ephb_T_max) {
    SharedGrid g;
    // thermal photons beyond kTopX * Tmax carry < e^{-(kTopX - wien_from)} of any integral served here: no panels there
    g.y_top = log(fmin(fmin(Ephb_T_max, kTopX) * Tmax, E_top));
    const double y_split = log(Tmin) - 1.0, y_want = log(Tmin) - 25.0;
    int nf = (int)ceil(g.y_top - y_split);
    nf = max(1, min(nf, kMaxPan - 1));
    int nc = (int)ceil((g.y_top - (double)nf - y_want) / kHC);
    nc = max(1, min(nc, kMaxPan - nf));
    g.n_fine = nf;
    g.n_pan = nf + nc;
    g.y_bot = g.y_top - (double)nf - kHC * (double)nc;
    return g;
}

// panel pn counted from the top: lower edge and width
__device__ __forceinline__ void panel_geom(const SharedGrid& g, int pn, double& lo, double& h) {
    if (pn < g.n_fine) { h = 1.0; lo = g.y_top - (double)(pn + 1); }
    else { h = kHC; lo = g.y_top - (double)g.n_fine - kHC * (double)(pn - g.n_fine + 1); }
}

// (pair, T) combinations served by the shared grid; evaluated identically by producer and consumer
// (wien_from = acro_params_t.wien_from, e <= wien_from <= e^2: above it the combination goes to kernels_wien_kernel)
__device__ __forceinline__ bool shared_regime(double x_a, double wien_from, double T, double x_grid_bot, double T_min) {
    return (x_a <= wien_from * T) && (x_a >= x_grid_bot || x_grid_bot <= 2e-11 * T_min);
}

struct PairLimits {           // T-independent limits of the three integrals of one pair; x_a < 0 marks "no range"
    double xa9, xb9, xa10, xb10, xa11, xb11;
};

__device__ __forceinline__ PairLimits pair_limits(double E, double Ep) {
    PairLimits L;
    L.xa9 = L.xa10 = L.xa11 = -1.0;
    L.xb9 = L.xb10 = L.xb11 = 0.0;
    const double dE = Ep - E, ub = Ep - kMe2 / (4.0 * Ep);
    if (!(Ep < E + kMe)) { L.xa9 = 0.25 * kMe2 * E / (Ep * Ep - Ep * E); L.xb9 = fmin(E, ub); }
    if (E != Ep) {
        const double pf = 0.25 * kMe2 / Ep - E, qf = 0.25 * kMe2 * dE / Ep;
        const double sq = sqrt((0.5 * pf) * (0.5 * pf) - qf);
        L.xa10 = -0.5 * pf - sq; L.xb10 = fmin(-0.5 * pf + sq, ub);
    }
    {
        const double rt = sqrt(E * E - kMe2), den = 4.0 * Ep * dE + kMe2;
        const double z1 = Ep * (kMe2 - 2.0 * dE * (rt - E)) / den, z2 = Ep * (kMe2 + 2.0 * dE * (rt + E)) / den;
        L.xa11 = fmax(kMe2 / Ep, z1); L.xb11 = fmin(z2, Ep);
    }
    return L;
}

// Legendre moments of f over [lo_y, hi_y] inside the panel [p_lo, p_lo+h] (accumulated into mu):
//   mu_j += int_{lo_y}^{hi_y} f(y) P_j(t(y)) dy      with the panel coordinate t in [-1, 1]
template <int KIND>
__device__ __forceinline__ void add_moments(const PairConst& c, double x_a, double lna, double lo_y, double hi_y, double p_lo,
                                            double h, double mu[kPanN]) {
    const double hp = 0.5 * (hi_y - lo_y), mp = 0.5 * (hi_y + lo_y);
    const double ic = 2.0 / h, mid = p_lo + 0.5 * h;
#pragma unroll 1
    for (int s = 0; s < kPanN; ++s) {
        const double y = fma(hp, c_glx[gl_offset(kPanN) + s], mp);
        const double x = fexp(y);
        double f;
        if (KIND == KIND_IC_PHOTON) f = integrand_core<KIND_IC_PHOTON>(c, x, x_a, y - lna, true);
        else if (KIND == KIND_IC_ELECTRON) f = integrand_core<KIND_IC_ELECTRON>(c, x, x_a, 0.0, false);
        else f = integrand_core<KIND_PC_ELECTRON>(c, x, x_a, 0.0, false) * frcp(x);
        f *= c_glw[gl_offset(kPanN) + s] * hp;
        const double tt = (y - mid) * ic;
        double pm = 1.0, pc = tt;
        mu[0] += f;
        mu[1] = fma(f, tt, mu[1]);
#pragma unroll
        for (int j = 1; j < kPanN - 1; ++j) {
            const double pnx = fma(c_leg_rec[j][0] * tt, pc, -c_leg_rec[j][1] * pm);
            pm = pc; pc = pnx;
            mu[j + 1] = fma(f, pc, mu[j + 1]);
        }
    }
}

// int_{lo_y}^{hi_y} f(y) B(y) dy ~= sum_m nu[m] B(y_m), B = the degree-15 interpolant through the panel's nodes.
// `y_layer` (or a value >= hi_y + 1 for "none"): position of a kinematic upper limit where the integrand has a boundary
// layer much narrower than a panel (the electron kernel: q -> 1 within a few per cent of x_b0, peak at 1 - x/x_b0 ~ 0.03);
// the interval is then cut at y_layer - 1/4 and y_layer - 1/16 so that the pieces are graded towards it.
template <int KIND>
__device__ __forceinline__ void partial_panel(const PairConst& c, double x_a, double lna, double lo_y, double hi_y, double p_lo,
                                              double h, double y_layer, double nu[kPanN]) {
    double mu[kPanN];
#pragma unroll
    for (int j = 0; j < kPanN; ++j) mu[j] = 0.0;
    double from = lo_y;
    const double c1 = y_layer - 0.25, c2 = y_layer - 0.0625;
    if (c1 > from && c1 < hi_y) { add_moments<KIND>(c, x_a, lna, from, c1, p_lo, h, mu); from = c1; }
    if (c2 > from && c2 < hi_y) { add_moments<KIND>(c, x_a, lna, from, c2, p_lo, h, mu); from = c2; }
    add_moments<KIND>(c, x_a, lna, from, hi_y, p_lo, h, mu);
#pragma unroll
    for (int m = 0; m < kPanN; ++m) {
        double v = 0.0;
#pragma unroll
        for (int j = 0; j < kPanN; ++j) v = fma(c_leg_modal[j][m], mu[j], v);
        nu[m] = v;
    }
}

__device__ __forceinline__ int panel_of(const SharedGrid& g, double y) {   // panel (from the top) that contains y
    int pn;
    if (y >= g.y_top - (double)g.n_fine) pn = (int)floor(g.y_top - y);
    else pn = g.n_fine + (int)floor((g.y_top - (double)g.n_fine - y) / kHC);
    return max(0, min(pn, g.n_pan - 1));
}

// One integral of one pair on the shared grid for the kTT temperatures of the current tile.
//   sx/six/sy: node tables x, 1/x, y;  sB: raw Bose values [node][kTT]
// Range: [x_a, x_b0] clipped to the grid; the thermal cut-off 200 T is in the table (zeros above it).
template <int KIND>
__device__ __forceinline__ void shared_integral(const PairConst& c, double x_a, double x_b0, const SharedGrid& g,
                                                const double* __restrict__ sx, const double* __restrict__ six,
                                                const double* __restrict__ sy, const double* __restrict__ sB, double* __restrict__ myW,
                                                bool active, double acc[kTT]) {
#pragma unroll
    for (int t = 0; t < kTT; ++t) acc[t] = 0.0;
    double a = active ? flog(fmax(x_a, 1e-300)) : g.y_top;
    const double lna = a;
    double bb = active ? flog(x_b0) : g.y_top;
    int p_first = -1, p_last = 0;     // lowest / highest panel touched (from the top: p_first >= p_last)
    bool part_lo = false, part_hi = false;
    if (active) {
        if (a <= g.y_bot) { a = g.y_bot; p_first = g.n_pan - 1; }
        else { p_first = panel_of(g, a); part_lo = true; }
        if (bb >= g.y_top) { bb = g.y_top; p_last = 0; }
        else { p_last = panel_of(g, bb); part_hi = true; }
        if (bb <= a) { p_first = -1; }                          // nothing inside the grid
    }
    const bool act = p_first >= 0;
    int p_lo = p_first;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) p_lo = max(p_lo, __shfl_xor_sync(0xffffffffu, p_lo, o));
    for (int pn = p_lo; pn >= 0; --pn) {
        double lo, h;
        panel_geom(g, pn, lo, h);
        const bool inside = act && pn <= p_first && pn >= p_last;
        // panels integrated through Legendre moments on their own nodes: the ones cut by a limit, and -- for the electron
        // kernel -- the full panel just below a kinematic upper limit (its boundary layer can reach into it)
        const bool layer = (KIND == KIND_IC_ELECTRON) && part_hi;
        const bool is_part = inside && ((part_lo && pn == p_first) || (part_hi && pn == p_last) || (layer && pn == p_last + 1));
        const bool full = inside && !is_part;
        const int base = pn * kPanN;
        if (!__any_sync(0xffffffffu, inside)) continue;
        const bool any_part = __any_sync(0xffffffffu, is_part), any_full = __any_sync(0xffffffffu, full);
        if (any_part) {
            double nu[kPanN];
#pragma unroll
            for (int m = 0; m < kPanN; ++m) nu[m] = 0.0;
            if (is_part) {
                const double lo_y = (part_lo && pn == p_first) ? a : lo;
                const double hi_y = (part_hi && pn == p_last) ? bb : lo + h;
                partial_panel<KIND>(c, x_a, lna, lo_y, hi_y, lo, h, layer ? bb : hi_y + 2.0, nu);
            }
            // through the thread's shared-memory column, so that ONE accumulation block below serves the lanes that are
            // partial in this panel and the lanes that are full (a warp usually has both near the lower limits)
#pragma unroll
            for (int m = 0; m < kPanN; ++m) myW[m * kSharedThreads] = nu[m];
        }
        // Node values first, accumulation second, kGroup nodes at a time: the evaluations are independent dependency
        // chains that the scheduler interleaves (two warps per scheduler cannot hide one chain behind another warp's);
        // more than two outgrow the L0 instruction cache.  Branch-free: lanes without this panel carry weight 0 and
        // finite arguments.
        const double wscale = full ? 0.5 * h : 0.0;
        constexpr int kGroup = ACRO_TS_FISSION;
        static_assert(kPanN % kGroup == 0, "node group must divide the panel");
#pragma unroll 1
        for (int m0 = 0; m0 < kPanN; m0 += kGroup) {
            double fv[kGroup];
            if (any_full) {
#pragma unroll
                for (int mm = 0; mm < kGroup; ++mm) {
                    const int m = m0 + mm;
                    const int k = base + m;
                    const double x = sx[k];
                    double f;
                    if (KIND == KIND_IC_PHOTON) {
                        const double q = x_a * six[k];
                        const double omq = 1.0 - q;
                        f = fma(2.0 * q, lna - sy[k], fma(fma(2.0, q, 1.0), omq, c.c9 * omq));
                    } else if (KIND == KIND_IC_ELECTRON) {
                        f = integrand_core<KIND_IC_ELECTRON>(c, x, x_a, 0.0, false);
                    } else {   // the table holds x^2/(e^{x/T}-1); the pair-creation integrand carries x^1
                        f = integrand_core<KIND_PC_ELECTRON>(c, x, x_a, 0.0, false) * six[k];
                    }
                    fv[mm] = full ? f * (c_glw[gl_offset(kPanN) + m] * wscale) : 0.0;
                }
            } else {
#pragma unroll
                for (int mm = 0; mm < kGroup; ++mm) fv[mm] = 0.0;
            }
            if (any_part) {
#pragma unroll
                for (int mm = 0; mm < kGroup; ++mm) fv[mm] += myW[(m0 + mm) * kSharedThreads];
            }
#pragma unroll
            for (int mm = 0; mm < kGroup; ++mm) {
                const double* __restrict__ Bk = sB + (size_t)(base + m0 + mm) * kTT;
                const double f = fv[mm];
#pragma unroll
                for (int t = 0; t < kTT; ++t) acc[t] = fma(f, Bk[t], acc[t]);
            }
        }
    }
}

// The kernel builder of the production path: grid (pair tiles, points), kSharedThreads threads, dynamic shared memory.
// One thread owns one (E,E') pair of one model point and produces the five kernel planes for ALL temperatures of the
// point: closed forms (pair-only parts computed once) and the temperature-shared integrals.  The (pair, T) combinations
// whose lower limit sits on the Wien tail are completed by kernels_wien_kernel (below).
__global__ void __launch_bounds__(kSharedThreads, 1) kernels_shared_kernel(acro_batch_t b, double* __restrict__ Kws, DeviceTables tb,
                                                                          acro_params_t p) {
    extern __shared__ double sm[];
    const int pt = blockIdx.y;
    const int ne = b.pt_ne[pt], nt = b.pt_nt[pt], s0 = b.pt_sys_off[pt];
    const int64_t tri = (int64_t)ne * (ne + 1) / 2;
    const int tiles = (int)((tri + kSharedThreads - 1) / kSharedThreads);
    const int tile0 = blockIdx.x * kTilesPerBlock, tile_end = min(tiles, tile0 + kTilesPerBlock);
    if (tile0 >= tiles) return;
    const double* eg = b.e_grid + b.pt_e_off[pt];
    const double Tmin = b.sys_T[s0], Tmax = b.sys_T[s0 + nt - 1];
    const SharedGrid g = shared_grid(Tmin, Tmax, eg[ne - 1], p.Ephb_T_max);
    const int nn = g.n_pan * kPanN;
    double* sx = sm;
    double* six = sx + kMaxPan * kPanN;
    double* sy = six + kMaxPan * kPanN;
    double* sB = sy + kMaxPan * kPanN;                 // [nn][kTT]
    double* sT = sB + (size_t)kMaxPan * kPanN * kTT;   // per-temperature scalars of the tile: T, ne, nNZ2, photon-photon prefactor
    int64_t* sK = reinterpret_cast<int64_t*>(sT + kTileScalars * kTT);   // plane offsets of the tile's systems
#if ACRO_FQ_LOGTAB
    double2* s_lt = reinterpret_cast<double2*>(sT + (kTileScalars + 1) * kTT);
    for (int k = threadIdx.x; k < kLogTab; k += blockDim.x) s_lt[k] = g_logtab[k];
    double* sA = reinterpret_cast<double*>(s_lt + kLogTab);   // [kTT][kSharedThreads] accumulators on their way out
#else
    double* sA = sT + (kTileScalars + 1) * kTT;
#endif
    for (int k = threadIdx.x; k < nn; k += blockDim.x) {
        double lo, h;
        panel_geom(g, k / kPanN, lo, h);
        const double y = lo + 0.5 * h * (1.0 + c_glx[gl_offset(kPanN) + (k % kPanN)]);
        const double x = fexp(y);
        sy[k] = y; sx[k] = x; six[k] = frcp(x);
    }
    for (int t0 = 0; t0 < nt; t0 += kTT) {
        __syncthreads();
        if (threadIdx.x < kTT) {
            const int t = threadIdx.x;
            const int s = s0 + min(t0 + t, nt - 1);
            const double T = b.sys_T[s];
            const Densities d = densities(T, tb.eta, tb.Yp, tb.YHe4);
            const double T3 = T * T * T;
            sT[t] = T; sT[kTT + t] = d.ne; sT[2 * kTT + t] = d.nNZ2; sT[3 * kTT + t] = T3 * T3;
            sT[4 * kTT + t] = p.Ephb_T_max * T;       // thermal cut-off of the photon background
            sT[5 * kTT + t] = p.wien_from * T;           // lower limits above this go to the Wien-tail kernel
            sT[6 * kTT + t] = kMe2 / (50.0 * T);      // photon-photon pair creation switches on here (cascade.py:583)
            sT[7 * kTT + t] = -T / kMe2;
            sK[t] = b.sys_k_off[s];
        }
        __syncthreads();
        // Bose table of this temperature tile: raw x^2/(pi^2 (e^{x/T}-1)), zero above the cut-off 200 T
        for (int idx = threadIdx.x; idx < nn * kTT; idx += blockDim.x) {
            const int k = idx / kTT, t = idx % kTT;
            double v = 0.0;
            if (t0 + t < nt) {
                const double T = sT[t];
                const double x = sx[k];
                if (x <= sT[4 * kTT + t]) v = x * x * bose_inv(x / T) * (1.0 / kPi2);
            }
            sB[idx] = v;
        }
        __syncthreads();
        const int ntile = min(kTT, nt - t0);
        // the block walks over kTilesPerBlock tiles of 256 pairs: the node tables and the Bose table above are per point and
        // temperature tile, not per pair (filling them once per 256 pairs was 17 % of the instructions)
        for (int tile = tile0; tile < tile_end; ++tile) {
        // this thread's pair
        const int64_t loc = (int64_t)tile * kSharedThreads + threadIdx.x;
        const bool active = loc < tri;
        int i = 0, j = 0;
        if (active) {
            const double tn = 2.0 * ne + 1.0;
            i = (int)floor(0.5 * (tn - sqrt(tn * tn - 8.0 * (double)loc)));
            i = max(0, min(i, ne - 1));
            while (i > 0 && tri_row_off(i, ne) > loc) --i;
            while (i < ne - 1 && tri_row_off(i + 1, ne) <= loc) ++i;
            j = i + (int)(loc - tri_row_off(i, ne));
        }
        const double E = eg[i], Ep = eg[j];
        PairConst c;
        c.E = E; c.Ep = Ep; c.dE = Ep - E;
        {
            const double r = E / fmax(c.dE, 1e-300);
            c.c9 = r * r / (2.0 + 2.0 * r);
            c.c10 = 0.25 * kMe2 / Ep;
            c.i2Ep = 0.5 / Ep;
        }
    #if ACRO_FQ_LOGTAB
        c.lt = s_lt;
    #endif
        const PairLimits L = pair_limits(E, Ep);
        const double x_bot = exp(g.y_bot);
        const int64_t np = b.n_pairs_total;
        // pair-only parts of the closed forms (cascade.py:383-402, 550-558, 621-639)
        const double a2 = kAlpha * kAlpha;
        const double pre_ic = 2.0 * kPi * a2 / (Ep * Ep);
        const double pre_pc = 0.25 * kPi * a2 * kMe2 / (Ep * Ep * Ep);
        const double rr = E / Ep, gg_shape = 1.0 - rr + rr * rr;
        const double pp_pair = 1112.0 / (10125.0 * kPi) * (a2 * a2) / ((kMe2 * kMe2) * (kMe2 * kMe2)) * 8.0 * (kPi2 * kPi2) / 63.0 *
                               (Ep * Ep) * (gg_shape * gg_shape);
        const double compton_g = kernel_compton_core(E, Ep, 1.0);               // per unit electron density
        const double compton_e = kernel_compton_core(Ep + kMe - E, Ep, 1.0);
        const double bh_pair = (L.xa9 > 0.0) ? bh_dsdE_Z2(E, Ep) : 0.0;         // same condition E' >= E + me

        // the lowest grid node is below every lower limit that matters (shared_regime(), second clause)
        const bool deep = x_bot <= 2e-11 * Tmin;
        const bool has9 = active && L.xa9 > 0.0, has10 = active && L.xa10 > 0.0, has11 = active && L.xb11 > L.xa11;
        const bool reg9 = has9 && (deep || L.xa9 >= x_bot), reg10 = has10 && (deep || L.xa10 >= x_bot),
                   reg11 = has11 && (deep || L.xa11 >= x_bot);
        double acc[kTT];
        // a pair takes part in the shared integrals of this tile only if its lower limit is off the Wien tail for the
        // tile's highest temperature
        const double xa_max = sT[5 * kTT + ntile - 1];
        // Epilogues: acc[] must only ever be indexed with compile-time constants (a run-time index puts the accumulators
        // into local memory for the whole kernel), and a 40-fold unrolled epilogue costs ~20 KB of code against a 6 KB L0
        // instruction cache -- so the accumulators go through a thread-private shared-memory column and the store loops
        // stay rolled.
        double* __restrict__ myA = sA + threadIdx.x;
        // ---- e -> gamma (plane 1)
        {
            const bool act = has9 && L.xb9 > L.xa9 && L.xa9 <= xa_max;
            shared_integral<KIND_IC_PHOTON>(c, L.xa9, L.xb9, g, sx, six, sy, sB, myA, act, acc);
#pragma unroll
            for (int t = 0; t < kTT; ++t) myA[t * kSharedThreads] = acc[t];
            if (active) {
                double* __restrict__ out = Kws + 1 * np + loc;
                for (int t = 0; t < ntile; ++t) {
                    const bool on = reg9 && fmin(L.xb9, sT[4 * kTT + t]) > L.xa9 && L.xa9 <= sT[5 * kTT + t];
                    out[sK[t]] = on ? pre_ic * myA[t * kSharedThreads] : 0.0;
                }
            }
        }
        // ---- e -> e (plane 4)
        {
            const bool act = has10 && L.xb10 > L.xa10 && L.xa10 <= xa_max;
            shared_integral<KIND_IC_ELECTRON>(c, L.xa10, L.xb10, g, sx, six, sy, sB, myA, act, acc);
#pragma unroll
            for (int t = 0; t < kTT; ++t) myA[t * kSharedThreads] = acc[t];
            if (active) {
                double* __restrict__ out = Kws + 4 * np + loc;
                for (int t = 0; t < ntile; ++t) {
                    const bool on = reg10 && fmin(L.xb10, sT[4 * kTT + t]) > L.xa10 && L.xa10 <= sT[5 * kTT + t];
                    out[sK[t]] = on ? pre_ic * myA[t * kSharedThreads] : 0.0;
                }
            }
        }
        // ---- gamma -> e-/e+ (planes 2, 3) and gamma -> gamma (plane 0)
        {
            const bool act = has11 && L.xa11 <= xa_max && !(Ep < sT[6 * kTT + ntile - 1]);
            shared_integral<KIND_PC_ELECTRON>(c, L.xa11, L.xb11, g, sx, six, sy, sB, myA, act, acc);
#pragma unroll
            for (int t = 0; t < kTT; ++t) myA[t * kSharedThreads] = acc[t];
            if (active) {
                double* __restrict__ out = Kws + loc;
                for (int t = 0; t < ntile; ++t) {
                    const double dne = sT[kTT + t];
                    const bool on = reg11 && !(Ep < sT[6 * kTT + t]) && fmin(L.xb11, sT[4 * kTT + t]) > L.xa11 &&
                                    L.xa11 <= sT[5 * kTT + t];
                    const double pair = fma(sT[2 * kTT + t], bh_pair, on ? pre_pc * myA[t * kSharedThreads] : 0.0);
                    const double arg = Ep * sT[7 * kTT + t];
                    const double damp = arg > -700.0 ? fexp(arg) : 0.0;
                    const int64_t o = sK[t];
                    out[3 * np + o] = pair;
                    out[2 * np + o] = fma(dne, compton_e, pair);
                    out[o] = fma(pp_pair * sT[3 * kTT + t], damp, dne * compton_g);
                }
            }
        }
        }   // tiles
    }
}

// Second kernel of the production path: the (pair, T) combinations that the shared grid does not serve (lower limit on
// the Wien tail) get their own Gauss-Laguerre / Gauss-Legendre rule (acro_fastquad.cuh), written over / added to what
// kernels_shared_kernel left in the planes.
//
// Work distribution: a thread owns one (E,E') pair of a point; for each of the three integrals the temperatures it has to
// do form a contiguous index range (temperatures ascend, all predicates are monotone).  The ranges differ a lot between
// the 32 pairs of a warp (ncu: 15.6 of 32 lanes active when every lane simply loops over its own range), so the warp
// pools its items: exclusive prefix sum of the range lengths, then 32 items at a time, each lane fetching the constants
// of the owning pair with shuffles.
#ifndef ACRO_WT_MINBLOCKS
#define ACRO_WT_MINBLOCKS 8
#endif

// first index t in [0, nt) with pred(T[t]) true, for a predicate that is false...false true...true along the grid
template <class Pred>
__device__ __forceinline__ int first_true(const double* __restrict__ T, int nt, Pred pred) {
    int lo = 0, hi = nt;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (pred(T[mid])) hi = mid; else lo = mid + 1;
    }
    return lo;
}

__device__ __forceinline__ double shfl_d(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }

constexpr int kWienChunk = 64;   // temperatures pooled per pass

// Items are enumerated TEMPERATURE-MAJOR: for one temperature the lanes that have work form (nearly) a contiguous set of
// neighbouring pairs with similar x_a/T, hence the same Gauss-Laguerre order and coalesced stores; pooling consecutive
// temperatures fills the lanes that would idle.  (Pair-major pooling was measured 25 % SLOWER than no pooling: every
// batch of 32 items then spans the whole x_a/T range of a pair and runs at the largest node count.)
//   smask[t]  = ballot of the lanes whose range contains temperature c_lo + t,  spref[t] = exclusive prefix of popc
template <int KIND, int Q>
__device__ __forceinline__ void wien_pooled(const acro_batch_t& b, double* __restrict__ Kws, const acro_params_t& p, int s0, int nt,
                                            int t_lo, int n_mine, double x_a, double x_b0, double c0, double c1, double c2,
                                            double c3, double pre, long long loc, int lane, unsigned* __restrict__ smask,
                                            int* __restrict__ spref, const double2* __restrict__ lt) {
    const int64_t np = b.n_pairs_total;
    const int t_hi = t_lo + n_mine;
    int w_lo = n_mine > 0 ? t_lo : nt, w_hi = n_mine > 0 ? t_hi : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        w_lo = min(w_lo, __shfl_xor_sync(0xffffffffu, w_lo, o));
        w_hi = max(w_hi, __shfl_xor_sync(0xffffffffu, w_hi, o));
    }
    for (int c_lo = w_lo; c_lo < w_hi; c_lo += kWienChunk) {
        const int nc = min(kWienChunk, w_hi - c_lo);
        // masks and their prefix (lane tt handles temperature c_lo+tt and c_lo+32+tt)
        __syncwarp();
        for (int tt = 0; tt < nc; ++tt) {
            const int t = c_lo + tt;
            const unsigned m = __ballot_sync(0xffffffffu, t >= t_lo && t < t_hi);
            if (lane == 0) smask[tt] = m;
        }
        __syncwarp();
        int cnt0 = (lane < nc) ? __popc(smask[lane]) : 0;
        int cnt1 = (32 + lane < nc) ? __popc(smask[32 + lane]) : 0;
        int inc0 = cnt0, inc1 = cnt1;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v0 = __shfl_up_sync(0xffffffffu, inc0, o), v1 = __shfl_up_sync(0xffffffffu, inc1, o);
            if (lane >= o) { inc0 += v0; inc1 += v1; }
        }
        const int tot0 = __shfl_sync(0xffffffffu, inc0, 31);
        const int total = tot0 + __shfl_sync(0xffffffffu, inc1, 31);
        spref[lane] = inc0 - cnt0;
        spref[32 + lane] = tot0 + inc1 - cnt1;
        __syncwarp();
        for (int base = 0; base < total; base += 32) {
            const int q = base + lane;
            const bool live = q < total;
            int tt = 0, own = 0;
            if (live) {
                int lo = 0, hi = nc - 1;          // last temperature slot whose exclusive prefix is <= q and that is non-empty
                while (lo < hi) {
                    const int mid = (lo + hi + 1) >> 1;
                    if (spref[mid] <= q) lo = mid; else hi = mid - 1;
                }
                tt = lo;
                own = __fns(smask[tt], 0, q - spref[tt] + 1);   // the (q - prefix + 1)-th set bit
            }
            const double o_xa = shfl_d(x_a, own), o_xb0 = shfl_d(x_b0, own), o_pre = shfl_d(pre, own);
            const double o_c0 = shfl_d(c0, own), o_c1 = shfl_d(c1, own), o_c2 = shfl_d(c2, own), o_c3 = shfl_d(c3, own);
            const long long o_loc = __shfl_sync(0xffffffffu, loc, own);
            if (live) {
                const int t = c_lo + tt;
                const double T = b.sys_T[s0 + t];
                const double ulim = fmin(o_xb0, p.Ephb_T_max * T);
                PairConst c;
#if ACRO_FQ_LOGTAB
                c.lt = lt;
#else
                (void)lt;
#endif
                if (KIND == KIND_IC_PHOTON) { c.c9 = o_c0; }
                else if (KIND == KIND_IC_ELECTRON) { c.E = o_c0; c.dE = o_c1; c.c10 = o_c2; c.i2Ep = o_c3; }
                else { c.E = o_c0; c.Ep = o_c1; }
                const int64_t o = b.sys_k_off[s0 + t] + o_loc;
                // the pair-creation integral is ADDED to what the shared kernel left in the gamma -> e planes (closed forms):
                // request those two values before the integral, not after it (the read-modify-write was 4 % of the samples)
                double k3 = 0.0, k2 = 0.0;
                if (KIND == KIND_PC_ELECTRON) { k3 = Kws[3 * np + o]; k2 = Kws[2 * np + o]; }
                const double v = o_pre * thermal_integral<KIND, Q>(c, T, o_xa, ulim);
                if (KIND == KIND_IC_PHOTON) Kws[1 * np + o] = v;
                else if (KIND == KIND_IC_ELECTRON) Kws[4 * np + o] = v;
                else { Kws[3 * np + o] = k3 + v; Kws[2 * np + o] = k2 + v; }
            }
        }
    }
}

template <int Q>
__global__ void __launch_bounds__(128, ACRO_WT_MINBLOCKS) kernels_wien_kernel(acro_batch_t b, double* __restrict__ Kws, acro_params_t p) {
    const int pt = blockIdx.y;
    const int ne = b.pt_ne[pt], nt = b.pt_nt[pt], s0 = b.pt_sys_off[pt];
    const int64_t tri = (int64_t)ne * (ne + 1) / 2;
    __shared__ unsigned smask_all[4 * kWienChunk];
    __shared__ int spref_all[4 * kWienChunk];
    const int lane = threadIdx.x & 31;
    unsigned* smask = smask_all + (threadIdx.x >> 5) * kWienChunk;
    int* spref = spref_all + (threadIdx.x >> 5) * kWienChunk;
    const int64_t loc = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
#if ACRO_FQ_LOGTAB
    __shared__ double2 s_lt[kLogTab];
    for (int k = threadIdx.x; k < kLogTab; k += blockDim.x) s_lt[k] = g_logtab[k];
    __syncthreads();
    const double2* lt = s_lt;
#else
    const double2* lt = nullptr;
#endif
    if (loc - lane >= tri) return;                     // whole warp beyond the point's pairs
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
    const double* Tg = b.sys_T + s0;
    const double Tmin = Tg[0], Tmax = Tg[nt - 1];
    const PairLimits L = pair_limits(E, Ep);
    const SharedGrid g = shared_grid(Tmin, Tmax, eg[ne - 1], p.Ephb_T_max);
    const double x_bot = exp(g.y_bot);
    const bool grid_full = (x_bot <= 2e-11 * Tmin);    // same expression as in shared_regime
    const double a2 = kAlpha * kAlpha;
    const double pre_ic = 2.0 * kPi * a2 / (Ep * Ep), pre_pc = 0.25 * kPi * a2 * kMe2 / (Ep * Ep * Ep);
    const double xf = p.Ephb_T_max;
    // range of temperature indices that this pair has to do for one integral: [first non-empty range, first T that the
    // shared grid serves) -- the predicates are the expressions used by kernels_shared_kernel, evaluated on the same T
    auto range = [&](double x_a, double x_b0, bool has, int t_extra, int& t_lo, int& n) {
        t_lo = 0; n = 0;
        if (!valid || !has || !(x_b0 > x_a)) return;
        const int t_ne = first_true(Tg, nt, [&](double T) { return fmin(x_b0, xf * T) > x_a; });
        const bool served = (x_a >= x_bot) || grid_full;
        const int t_cut = served ? first_true(Tg, nt, [&](double T) { return x_a <= p.wien_from * T; }) : nt;
        t_lo = max(t_ne, t_extra);
        n = max(0, t_cut - t_lo);
    };
    int t_lo, n;
    {   // e -> gamma
        range(L.xa9, L.xb9, L.xa9 > 0.0, 0, t_lo, n);
        const double r = E / fmax(Ep - E, 1e-300);
        wien_pooled<KIND_IC_PHOTON, Q>(b, Kws, p, s0, nt, t_lo, n, L.xa9, L.xb9, r * r / (2.0 + 2.0 * r), 0.0, 0.0, 0.0, pre_ic, loc, lane, smask, spref, lt);
    }
    {   // e -> e
        range(L.xa10, L.xb10, L.xa10 > 0.0, 0, t_lo, n);
        wien_pooled<KIND_IC_ELECTRON, Q>(b, Kws, p, s0, nt, t_lo, n, L.xa10, L.xb10, E, Ep - E, 0.25 * kMe2 / Ep, 0.5 / Ep, pre_ic, loc, lane, smask, spref, lt);
    }
    {   // gamma -> e (pair creation): needs E' >= me^2/(50 T)
        const int t_pc = first_true(Tg, nt, [&](double T) { return !(Ep < kMe2 / (50.0 * T)); });
        range(L.xa11, L.xb11, true, t_pc, t_lo, n);
        wien_pooled<KIND_PC_ELECTRON, Q>(b, Kws, p, s0, nt, t_lo, n, L.xa11, L.xb11, E, Ep, 0.0, 0.0, pre_pc, loc, lane, smask, spref, lt);
    }
}

}  // namespace acro
