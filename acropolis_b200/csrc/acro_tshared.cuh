// acro_tshared.cuh -- shared pieces of the production kernel builder (kernel_mode = ACRO_KERNELS_FAST): the temperature-shared
// panel grid, the T-independent integration limits of a pair, and the per-temperature Wien-tail kernel (kernels_wien_kernel).
// The kernel that evaluates the integrals on the shared grid, kernels_mma_kernel, is in acro_tmma.cuh.
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
//     up to 48 passes of 512 pairs -- the "thermal-photon grid" of BASELINE.json:north_star;
//   * f is evaluated once per (pair, node) and contracted with the Bose table over the nodes for all temperatures of the
//     tile; the panel that contains the lower limit is integrated on its own nodes against the Legendre expansion of the Bose
//     factor (exact for what the panel rule itself resolves), and so is the panel that contains the upper limit when that is
//     the kinematic one (x_b0 < 200 T); the thermal cut-off 200 T is applied as zeros in the Bose table.
//
// kernels_mma_kernel (one launch for the e -> e plane, one for the gamma -> e planes and the closed-form gamma -> gamma
// plane) writes every entry of its planes; kernels_wien_kernel then overwrites the (pair, T) combinations beyond
// wien_from * T with its own Gauss-Laguerre rule (acro_fastquad.cuh).  Both evaluate the regime predicates with the same
// expressions on the same numbers.  The e -> gamma plane comes from ic_photon_closed_kernel (acro_icphoton.cuh).
//
// Tuning constants (each is a -D macro so that tools/gpu_ab_variants.py can time a variant library; values = measured best
// on the 840-point benchmark slice in round 1):
//   ACRO_TS_TT      40   temperatures per tile (five DMMA N-tiles)
//   ACRO_TS_PANN    12   nodes per panel (16: same answers to 5e-10, +14 % time; 10: 1.9e-8 and -15 %; 8: 1e-6, rejected)
//   ACRO_TS_HC      2.0  width of the coarse panels [e-folds]
//   ACRO_TS_MAXPAN  28   panels per point (12 fine + 12 coarse for a 2-decade T range; 17 + 11 for 4 decades)
//   ACRO_TS_TOPX    50   top of the grid in units of Tmax (200 = the reference's cut-off: same results, +9 % time)
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
constexpr int kTT = ACRO_TS_TT;
constexpr int kPanN = ACRO_TS_PANN;
constexpr double kHC = ACRO_TS_HC;
constexpr int kMaxPan = ACRO_TS_MAXPAN;
constexpr double kTopX = ACRO_TS_TOPX;
constexpr int kTileScalars = 8;   // per-temperature scalars of a tile kept in shared memory (kernels_mma_kernel)

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

__device__ __forceinline__ int panel_of(const SharedGrid& g, double y) {   // panel (from the top) that contains y
    int pn;
    if (y >= g.y_top - (double)g.n_fine) pn = (int)floor(g.y_top - y);
    else pn = g.n_fine + (int)floor((g.y_top - (double)g.n_fine - y) / kHC);
    return max(0, min(pn, g.n_pan - 1));
}

// Second kernel of the production path: the (pair, T) combinations that the shared grid does not serve (lower limit on
// the Wien tail) get their own Gauss-Laguerre / Gauss-Legendre rule (acro_fastquad.cuh), written over / added to what
// kernels_mma_kernel left in the planes.
//
// Work distribution: a thread owns one (E,E') pair of a point; for the launch's integral the temperatures it has to
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
//   smask[t]  = ballot of the lanes whose range contains temperature c_lo + t;  slist = the chunk's items in that order
template <int KIND, int Q>
__device__ __forceinline__ void wien_pooled(const acro_batch_t& b, double* __restrict__ Kws, const acro_params_t& p, int s0, int nt,
                                            int t_lo, int n_mine, double x_a, double x_b0, double c0, double c1, double c2,
                                            double c3, double pre, long long loc, int lane, unsigned* __restrict__ smask,
                                            unsigned short* __restrict__ slist,
                                            const double2* __restrict__ lt, const DeviceTables& tb) {
    const int64_t np = b.n_pairs_total;
    const int t_hi = t_lo + n_mine;
    int w_lo = n_mine > 0 ? t_lo : nt, w_hi = n_mine > 0 ? t_hi : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        w_lo = min(w_lo, __shfl_xor_sync(0xffffffffu, w_lo, o));
        w_hi = max(w_hi, __shfl_xor_sync(0xffffffffu, w_hi, o));
    }
    // Two passes over the pooled items: first the ones that take the Gauss-Laguerre rule (span (x_b - x_a)/T >= 32), then the
    // short-span ones (Gauss-Legendre, 32 nodes).  Mixed in one batch of 32 the few short-span items kept the whole warp
    // waiting through their 32 nodes: 18 % of the warp instructions of this kernel ran with 4-6 active lanes
    // (profiles/r02_ncu_kernels_wien_kernel_v1.txt).
#ifndef ACRO_WT_TWOPHASE
#define ACRO_WT_TWOPHASE 1
#endif
    // "Short" for this scheduling = what takes the 32-node rule: the range ends at the KINEMATIC limit within 32 T of x_a (a span
    // cut by the thermal limit 200 T costs 8 nodes, see thermal_integral).  Rare (0.14 % of the pair-creation items of the
    // benchmark scan, none in the other two integrals) and monotone in T: a lane whose last temperature is long-span has no
    // short item, and a warp without such a lane makes ONE pass with plain range masks (the classification only schedules
    // the items; thermal_integral picks the rule from the same numbers itself).
    auto is_short_at = [&](double T) { return !(x_b0 > p.Ephb_T_max * T) && !((x_b0 - x_a) >= kSteepSpan * T); };
    const bool two_phase = ACRO_WT_TWOPHASE && __any_sync(0xffffffffu, n_mine > 0 && is_short_at(b.sys_T[s0 + max(t_hi - 1, 0)]));
#pragma unroll 1
    for (int phase = 0; phase < (two_phase ? 2 : 1); ++phase) {
    for (int c_lo = w_lo; c_lo < w_hi; c_lo += kWienChunk) {
        const int nc = min(kWienChunk, w_hi - c_lo);
        // masks and their prefix (lane tt handles temperature c_lo+tt and c_lo+32+tt)
        __syncwarp();
        if (!two_phase) {
#pragma unroll 4
            for (int tt = 0; tt < nc; ++tt) {
                const int t = c_lo + tt;
                const unsigned m = __ballot_sync(0xffffffffu, t >= t_lo && t < t_hi);
                if (lane == 0) smask[tt] = m;
            }
        } else {
            for (int tt = 0; tt < nc; ++tt) {
                const int t = c_lo + tt;
                const bool is_short = is_short_at(b.sys_T[s0 + t]);
                const unsigned m = __ballot_sync(0xffffffffu, t >= t_lo && t < t_hi && is_short == (phase == 1));
                if (lane == 0) smask[tt] = m;
            }
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
        // item list of the chunk, temperature-major: (slot << 5) | owning lane.  Lane l expands the masks of slots l and 32 + l;
        // a batch then reads its 32 items with one load each (a binary search over the prefix plus __fns per item was 25 %
        // of this kernel's instructions)
        {
            int pos = inc0 - cnt0;
            unsigned m = (lane < nc) ? smask[lane] : 0u;
            while (m) { slist[pos++] = (unsigned short)((lane << 5) | (__ffs(m) - 1)); m &= m - 1; }
            pos = tot0 + inc1 - cnt1;
            m = (32 + lane < nc) ? smask[32 + lane] : 0u;
            while (m) { slist[pos++] = (unsigned short)(((32 + lane) << 5) | (__ffs(m) - 1)); m &= m - 1; }
        }
        __syncwarp();
        for (int base = 0; base < total; base += 32) {
            const int q = base + lane;
            const bool live = q < total;
            const int item = live ? slist[q] : 0;
            const int tt = item >> 5, own = item & 31;
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
                const double v = o_pre * thermal_integral<KIND, Q, true>(c, T, o_xa, ulim, o_xb0);
                if (KIND == KIND_IC_PHOTON) Kws[1 * np + o] = v;
                else if (KIND == KIND_IC_ELECTRON) Kws[4 * np + o] = v;
                else {
                    // the gamma -> e planes also carry closed forms (Bethe-Heitler o_c2, Compton recoil electron o_c3, per unit
                    // density; pair_closed_forms_kernel): written complete here, with the expressions of kernels_mma_kernel,
                    // so nothing is read back from the planes
                    const Densities d = densities(T, tb.eta, tb.Yp, tb.YHe4);
                    const double pair = fma(d.nNZ2, o_c2, v);
                    Kws[3 * np + o] = pair;
                    Kws[2 * np + o] = fma(d.ne, o_c3, pair);
                }
            }
        }
    }
    }
}

#ifndef ACRO_WT_SPLIT
#define ACRO_WT_SPLIT 1          // 1: one launch per integral kind (KSEL = 0, 1, 2) instead of one launch for all three (KSEL = -1)
#endif
template <int Q, int KSEL>
__global__ void __launch_bounds__(128, ACRO_WT_MINBLOCKS) kernels_wien_kernel(acro_batch_t b, double* __restrict__ Kws,
                                                                               const double* __restrict__ closed, DeviceTables tb,
                                                                               acro_params_t p) {
    const int pt = blockIdx.y;
    const int ne = b.pt_ne[pt], nt = b.pt_nt[pt], s0 = b.pt_sys_off[pt];
    const int64_t tri = (int64_t)ne * (ne + 1) / 2;
    __shared__ unsigned smask_all[4 * kWienChunk];
    __shared__ unsigned short slist_all[4 * kWienChunk * 32];
    const int lane = threadIdx.x & 31;
    unsigned* smask = smask_all + (threadIdx.x >> 5) * kWienChunk;
    unsigned short* slist = slist_all + (threadIdx.x >> 5) * kWienChunk * 32;
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
    // shared grid serves) -- the predicates are the expressions used by kernels_mma_kernel, evaluated on the same T
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
    if (KSEL < 0 || KSEL == KIND_IC_PHOTON) {   // e -> gamma
        range(L.xa9, L.xb9, L.xa9 > 0.0, 0, t_lo, n);
        const double r = E / fmax(Ep - E, 1e-300);
        wien_pooled<KIND_IC_PHOTON, Q>(b, Kws, p, s0, nt, t_lo, n, L.xa9, L.xb9, r * r / (2.0 + 2.0 * r), 0.0, 0.0, 0.0, pre_ic, loc, lane, smask, slist, lt, tb);
    }
    if (KSEL < 0 || KSEL == KIND_IC_ELECTRON) {   // e -> e
        range(L.xa10, L.xb10, L.xa10 > 0.0, 0, t_lo, n);
        wien_pooled<KIND_IC_ELECTRON, Q>(b, Kws, p, s0, nt, t_lo, n, L.xa10, L.xb10, E, Ep - E, 0.25 * kMe2 / Ep, 0.5 / Ep, pre_ic, loc, lane, smask, slist, lt, tb);
    }
    if (KSEL < 0 || KSEL == KIND_PC_ELECTRON) {   // gamma -> e (pair creation): needs E' >= me^2/(50 T)
        const int t_pc = first_true(Tg, nt, [&](double T) { return !(Ep < kMe2 / (50.0 * T)); });
        range(L.xa11, L.xb11, true, t_pc, t_lo, n);
        const int64_t oc = b.pt_pair_off[pt] + locc;
        wien_pooled<KIND_PC_ELECTRON, Q>(b, Kws, p, s0, nt, t_lo, n, L.xa11, L.xb11, E, Ep, closed[2 * b.n_pairs_points + oc],
                                         closed[b.n_pairs_points + oc], pre_pc, loc, lane, smask, slist, lt, tb);
    }
}

}  // namespace acro
