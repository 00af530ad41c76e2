// acro_tshared4.cuh -- four lanes per (E,E') pair: the temperature-shared kernel builder with its register load split in two
// dimensions (same integrals, same grid, same tables as kernels_shared_kernel in acro_tshared.cuh; ACRO_TS_LANES4 selects it).
//
// kernels_shared_kernel gives a pair to ONE thread: 12 node values per panel and 40 temperature sums.  The 40 accumulators
// (80 registers) allow one block of 8 warps per SM, and the dependent FP64 chains of the node values are exposed (ncu:
// FP64 pipe 40 % busy, stall_wait 42 %).  Here a warp takes 32 pairs at a time in two phases:
//   A. lane = pair: index decode, integration limits and the pair-only closed forms (Compton shapes, Bethe-Heitler, photon-
//      photon prefactor), once per pair, into the warp's shared-memory constant block;
//   B. four sub-passes of 8 pairs, the four lanes 4p..4p+3 sharing pair p:
//        node values   lane q evaluates nodes 3q..3q+2 of the panel (three independent chains) and publishes the weights
//                      in the warp's slot w[node][pair];
//        accumulation  lane q owns temperatures 10q..10q+9 of the tile: per node one weight (broadcast) and ten FMAs
//                      against the Bose table;
//        partial panels: each lane takes 3 of the 12 nodes of every sub-interval for the Legendre moments, the moments are
//                      summed over the four lanes with two xor-shuffles, and each lane folds 3 of the 12 node weights.
// Nothing is evaluated twice, the accumulators shrink to 20 registers, and a 512-thread block (16 warps, 128 registers)
// shares one Bose table.  The integral kind is a RUN-TIME argument (one copy of the panel walk instead of three): with 16
// warps in different phases the hot code has to fit the 32 KB instruction cache of the SM.
#pragma once
#include "acro_tshared.cuh"

namespace acro {

constexpr int kL4Threads = 512;
constexpr int kL4Warps = kL4Threads / 32;
constexpr int kL4Pairs = kL4Warps * 32;     // pairs per macro-pass of a block (32 per warp)
#ifndef ACRO_L4_LANES
#define ACRO_L4_LANES 4                     // lanes that share one pair: 4 or 2
#endif
constexpr int kLP = ACRO_L4_LANES;
constexpr int kLPshift = kLP == 4 ? 2 : 1;
constexpr int kL4PW = 32 / kLP;             // pairs per sub-pass of a warp
constexpr int kL4T = kTT / kLP;             // temperatures per lane
constexpr int kL4N = kPanN / kLP;           // nodes per lane
constexpr int kL4G = 3;                     // node values evaluated together (independent chains)
static_assert(kLP == 2 || kLP == 4, "two or four lanes per pair");
static_assert(kTT % (2 * kLP) == 0 && kPanN % kLP == 0 && kL4N % kL4G == 0,
              "tile and panel must split over the lanes (LDS.128 needs an even share)");
#ifndef ACRO_L4_FMA_UNROLL
#define ACRO_L4_FMA_UNROLL 4              // nodes per trip of the accumulation loop
#endif
#ifndef ACRO_L4_PASSES
#define ACRO_L4_PASSES 4                    // macro-passes (512 pairs each) per block and Bose-table fill (1: 33.5 ms, 2: 31.5, 4: 30.9, 8: 31.1)
#endif
constexpr int kL4PassesPerBlock = ACRO_L4_PASSES;
enum { L4_E = 0, L4_EP, L4_XA9, L4_XB9, L4_XA10, L4_XB10, L4_XA11, L4_XB11, L4_CG, L4_CE, L4_BH, L4_PP, L4_C9, L4_C10, L4_I2EP,
       L4_ACTIVE, L4_TON, L4_NCONST };     // L4_TON: the three first-served temperature indices, packed

__device__ __forceinline__ double node_core(int kind, const PairConst& c, double x, double ix, double y, double x_a, double lna) {
    if (kind == KIND_IC_PHOTON) {
        const double q = x_a * ix;
        const double omq = 1.0 - q;
        return fma(2.0 * q, lna - y, fma(fma(2.0, q, 1.0), omq, c.c9 * omq));
    } else if (kind == KIND_IC_ELECTRON) {
        return integrand_core<KIND_IC_ELECTRON>(c, x, x_a, 0.0, false);
    } else {   // the table holds x^2/(e^{x/T}-1); the pair-creation integrand carries x^1
        return integrand_core<KIND_PC_ELECTRON>(c, x, x_a, 0.0, false) * ix;
    }
}

// this lane's share (nodes q, q+4, q+8 of the sub-interval rule) of the Legendre moments of f over [lo_y, hi_y]
__device__ __forceinline__ void add_moments4(int kind, const PairConst& c, double x_a, double lna, double lo_y, double hi_y,
                                             double p_lo, double h, int q, double mu[kPanN]) {
    const double hp = 0.5 * (hi_y - lo_y), mp = 0.5 * (hi_y + lo_y);
    const double ic = 2.0 / h, mid = p_lo + 0.5 * h;
#pragma unroll 1
    for (int ss = 0; ss < kL4N; ++ss) {
        const int s = q + kLP * ss;
        const double y = fma(hp, c_glx[gl_offset(kPanN) + s], mp);
        const double x = fexp(y);
        double f = node_core(kind, c, x, frcp(x), y, x_a, lna);
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

// Partial panel of one pair, executed by its four lanes together (all lanes of the warp call it; `on` marks the pairs that
// really have a partial panel here, the others publish zeros).  Publishes the 12 node weights of the pair in wp[node * kL4PW].
__device__ __forceinline__ void partial_panel4(int kind, const PairConst& c, double x_a, double lna, double lo_y, double hi_y,
                                               double p_lo, double h, double y_layer, int q, bool on,
                                               const double* __restrict__ s_modal, double* __restrict__ wp) {
    double mu[kPanN];
#pragma unroll
    for (int j = 0; j < kPanN; ++j) mu[j] = 0.0;
    double from = lo_y;
#pragma unroll 1
    for (int piece = 0; piece < 3; ++piece) {           // graded cuts at y_layer - 1/4 and y_layer - 1/16, then the rest
        const double cut = piece == 0 ? y_layer - 0.25 : (piece == 1 ? y_layer - 0.0625 : hi_y);
        const bool take = on && (piece == 2 || (cut > from && cut < hi_y));
        if (take) { add_moments4(kind, c, x_a, lna, from, cut, p_lo, h, q, mu); from = cut; }
    }
    __syncwarp();
#pragma unroll
    for (int j = 0; j < kPanN; ++j) {
        mu[j] += __shfl_xor_sync(0xffffffffu, mu[j], 1);
        if (kLP == 4) mu[j] += __shfl_xor_sync(0xffffffffu, mu[j], 2);
    }
#pragma unroll
    for (int mm = 0; mm < kL4N; ++mm) {
        const int m = kL4N * q + mm;
        double v = 0.0;
#pragma unroll
        for (int j = 0; j < kPanN; ++j) v = fma(s_modal[j * kPanN + m], mu[j], v);
        wp[m * kL4PW] = on ? v : 0.0;
    }
}

// One integral of the warp's current 8 pairs on the shared grid for this lane's kL4T temperatures.
//   w: the warp's weight slot [kPanN][8];  sBq: Bose table shifted to this lane's temperatures
__device__ __forceinline__ void shared_integral4(int kind, const PairConst& c, double x_a, double x_b0, const SharedGrid& g,
                                                 const double* __restrict__ sx, const double* __restrict__ six,
                                                 const double* __restrict__ sy, const double* __restrict__ sBq,
                                                 const double* __restrict__ s_modal, double* __restrict__ w, int q, int pw,
                                                 bool active, double acc[kL4T]) {
#pragma unroll
    for (int t = 0; t < kL4T; ++t) acc[t] = 0.0;
    double a = active ? flog(fmax(x_a, 1e-300)) : g.y_top;
    const double lna = a;
    double bb = active ? flog(x_b0) : g.y_top;
    int p_first = -1, p_last = 0;
    bool part_lo = false, part_hi = false;
    if (active) {
        if (a <= g.y_bot) { a = g.y_bot; p_first = g.n_pan - 1; }
        else { p_first = panel_of(g, a); part_lo = true; }
        if (bb >= g.y_top) { bb = g.y_top; p_last = 0; }
        else { p_last = panel_of(g, bb); part_hi = true; }
        if (bb <= a) { p_first = -1; }
    }
    const bool act = p_first >= 0;
    int p_lo = p_first;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) p_lo = max(p_lo, __shfl_xor_sync(0xffffffffu, p_lo, o));
    double* __restrict__ wp = w + pw;          // this pair's column
    for (int pn = p_lo; pn >= 0; --pn) {
        double lo, h;
        panel_geom(g, pn, lo, h);
        const bool inside = act && pn <= p_first && pn >= p_last;
        const bool layer = (kind == KIND_IC_ELECTRON) && part_hi;
        const bool is_part = inside && ((part_lo && pn == p_first) || (part_hi && pn == p_last) || (layer && pn == p_last + 1));
        const bool full = inside && !is_part;
        const int base = pn * kPanN;
        if (!__any_sync(0xffffffffu, inside)) continue;
        const bool any_part = __any_sync(0xffffffffu, is_part), any_full = __any_sync(0xffffffffu, full);
        __syncwarp();                           // the slot is free: every lane has finished the previous accumulation
        if (any_part) {
            const double lo_y = (part_lo && pn == p_first) ? a : lo;
            const double hi_y = (part_hi && pn == p_last) ? bb : lo + h;
            partial_panel4(kind, c, x_a, lna, is_part ? lo_y : lo, is_part ? hi_y : lo + h, lo, h, layer ? bb : hi_y + 2.0, q,
                           is_part, s_modal, wp);
        }
        if (any_full) {
            const double wscale = full ? 0.5 * h : 0.0;
#pragma unroll 1
            for (int m0 = kL4N * q; m0 < kL4N * (q + 1); m0 += kL4G) {
                double fv[kL4G];
#pragma unroll
                for (int mm = 0; mm < kL4G; ++mm) {
                    const int m = m0 + mm;
                    const int k = base + m;
                    fv[mm] = node_core(kind, c, sx[k], six[k], sy[k], x_a, lna) * (c_glw[gl_offset(kPanN) + m] * wscale);
                }
#pragma unroll
                for (int mm = 0; mm < kL4G; ++mm)
                    if (full || !any_part) wp[(m0 + mm) * kL4PW] = full ? fv[mm] : 0.0;   // partial pairs keep their weights
            }
        }
        __syncwarp();                           // weights of all pairs are published
ACRO_UNROLL(ACRO_L4_FMA_UNROLL)
        for (int m = 0; m < kPanN; ++m) {
            const double v = wp[m * kL4PW];
            const double* __restrict__ Bk = sBq + (size_t)(base + m) * kTT;
#pragma unroll
            for (int t = 0; t < kL4T; ++t) acc[t] = fma(v, Bk[t], acc[t]);
        }
    }
}

// Pair-only closed forms (Compton shapes per unit electron density, Bethe-Heitler dsigma/dE for Z^2 = 1; cascade.py:383-402,
// 621-639) of every (E,E') pair of every point, thread per (point, pair).  They are ~1000 instructions that
// kernels_shared4_kernel would otherwise execute once per 32 pairs in the middle of its hot loops (instruction cache!); here
// they are parked in the point's OWN kernel-plane slots of its LAST temperature (planes 0, 1, 2): the main kernel reads them
// at the start of a pass and overwrites them with the final values only in the epilogue of the last temperature tile.
__global__ void __launch_bounds__(128) pair_closed_forms_kernel(acro_batch_t b, double* __restrict__ Kws) {
    const int pt = blockIdx.y;
    const int ne = b.pt_ne[pt];
    const int64_t tri = (int64_t)ne * (ne + 1) / 2;
    const int64_t loc = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (loc >= tri) return;
    const double tn = 2.0 * ne + 1.0;
    int i = (int)floor(0.5 * (tn - sqrt(tn * tn - 8.0 * (double)loc)));
    i = max(0, min(i, ne - 1));
    while (i > 0 && tri_row_off(i, ne) > loc) --i;
    while (i < ne - 1 && tri_row_off(i + 1, ne) <= loc) ++i;
    const int j = i + (int)(loc - tri_row_off(i, ne));
    const double* eg = b.e_grid + b.pt_e_off[pt];
    const double E = eg[i], Ep = eg[j];
    const int64_t o = b.sys_k_off[b.pt_sys_off[pt] + b.pt_nt[pt] - 1] + loc, np = b.n_pairs_total;
    Kws[o] = kernel_compton_core(E, Ep, 1.0);
    Kws[np + o] = kernel_compton_core(Ep + kMe - E, Ep, 1.0);
    Kws[2 * np + o] = !(Ep < E + kMe) ? bh_dsdE_Z2(E, Ep) : 0.0;        // the condition of pair_limits(): xa9 > 0
}

__global__ void __launch_bounds__(kL4Threads, 1) kernels_shared4_kernel(acro_batch_t b, double* __restrict__ Kws, DeviceTables tb,
                                                                        acro_params_t p) {
    extern __shared__ double sm[];
    const int pt = blockIdx.y;
    const int ne = b.pt_ne[pt], nt = b.pt_nt[pt], s0 = b.pt_sys_off[pt];
    const int64_t tri = (int64_t)ne * (ne + 1) / 2;
    const int passes = (int)((tri + kL4Pairs - 1) / kL4Pairs);
    const int pass0 = blockIdx.x * kL4PassesPerBlock, pass_end = min(passes, pass0 + kL4PassesPerBlock);
    if (pass0 >= passes) return;
    const double* eg = b.e_grid + b.pt_e_off[pt];
    const double Tmin = b.sys_T[s0], Tmax = b.sys_T[s0 + nt - 1];
    const SharedGrid g = shared_grid(Tmin, Tmax, eg[ne - 1], p.Ephb_T_max);
    const int nn = g.n_pan * kPanN;
    double* sx = sm;
    double* six = sx + kMaxPan * kPanN;
    double* sy = six + kMaxPan * kPanN;
    double* sB = sy + kMaxPan * kPanN;                 // [nn][kTT]
    double* sT = sB + (size_t)kMaxPan * kPanN * kTT;   // per-temperature scalars of the tile
    int64_t* sK = reinterpret_cast<int64_t*>(sT + kTileScalars * kTT);
    double2* s_lt = reinterpret_cast<double2*>(sT + (kTileScalars + 1) * kTT);
    double* s_modal = reinterpret_cast<double*>(s_lt + kLogTab);          // [kPanN][kPanN] copy of c_leg_modal
    double* s_w = s_modal + kPanN * kPanN;                                 // [warps][kPanN][8]
    double* s_c = s_w + (size_t)kL4Warps * kPanN * kL4PW;                      // [warps][L4_NCONST][32]
    for (int k = threadIdx.x; k < kLogTab; k += blockDim.x) s_lt[k] = g_logtab[k];
    for (int k = threadIdx.x; k < kPanN * kPanN; k += blockDim.x) s_modal[k] = c_leg_modal[k / kPanN][k % kPanN];
    for (int k = threadIdx.x; k < nn; k += blockDim.x) {
        double lo, h;
        panel_geom(g, k / kPanN, lo, h);
        const double y = lo + 0.5 * h * (1.0 + c_glx[gl_offset(kPanN) + (k % kPanN)]);
        const double x = fexp(y);
        sy[k] = y; sx[k] = x; six[k] = frcp(x);
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q = lane & (kLP - 1), pw = lane >> kLPshift;             // quarter of the pair's work, pair within the sub-pass
    double* w = s_w + (size_t)warp * kPanN * kL4PW;
    double* cw = s_c + (size_t)warp * L4_NCONST * 32;
    const double x_bot = exp(g.y_bot);
    const bool deep = x_bot <= 2e-11 * Tmin;
    const int64_t np = b.n_pairs_total;
    const int64_t k_park = b.sys_k_off[s0 + nt - 1];     // where pair_closed_forms_kernel parked the closed forms
    const int tq = q * kL4T;                            // first temperature of this lane
    const double a2 = kAlpha * kAlpha;

    for (int t0 = 0; t0 < nt; t0 += kTT) {
        __syncthreads();
        if (threadIdx.x < kTT) {
            const int t = threadIdx.x;
            const int s = s0 + min(t0 + t, nt - 1);
            const double T = b.sys_T[s];
            const Densities d = densities(T, tb.eta, tb.Yp, tb.YHe4);
            const double T3 = T * T * T;
            sT[t] = T; sT[kTT + t] = d.ne; sT[2 * kTT + t] = d.nNZ2; sT[3 * kTT + t] = T3 * T3;
            sT[4 * kTT + t] = p.Ephb_T_max * T;
            sT[5 * kTT + t] = p.wien_from * T;
            sT[6 * kTT + t] = kMe2 / (50.0 * T);
            sT[7 * kTT + t] = -T / kMe2;
            sK[t] = b.sys_k_off[s];
        }
        __syncthreads();
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
        const double xa_max = sT[5 * kTT + ntile - 1];
        const double pc_min = sT[6 * kTT + ntile - 1];
        const double* __restrict__ sBq = sB + tq;
        for (int pass = pass0; pass < pass_end; ++pass) {
            const int64_t loc0 = (int64_t)pass * kL4Pairs + warp * 32;
            if (loc0 >= tri) continue;              // warp-uniform: nothing left for this warp
            // ---- phase A: lane = pair
            __syncwarp();
            {
                const int64_t loc = loc0 + lane;
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
                const PairLimits L = pair_limits(E, Ep);
                const double r = E / fmax(Ep - E, 1e-300);
                const double rr = E / Ep, gg_shape = 1.0 - rr + rr * rr;
                cw[L4_E * 32 + lane] = E; cw[L4_EP * 32 + lane] = Ep;
                cw[L4_XA9 * 32 + lane] = L.xa9; cw[L4_XB9 * 32 + lane] = L.xb9;
                cw[L4_XA10 * 32 + lane] = L.xa10; cw[L4_XB10 * 32 + lane] = L.xb10;
                cw[L4_XA11 * 32 + lane] = L.xa11; cw[L4_XB11 * 32 + lane] = L.xb11;
                {
                    const int64_t o = k_park + (active ? loc : 0);
                    cw[L4_CG * 32 + lane] = Kws[o];
                    cw[L4_CE * 32 + lane] = Kws[np + o];
                    cw[L4_BH * 32 + lane] = Kws[2 * np + o];
                }
                cw[L4_PP * 32 + lane] = 1112.0 / (10125.0 * kPi) * (a2 * a2) / ((kMe2 * kMe2) * (kMe2 * kMe2)) * 8.0 * (kPi2 * kPi2) /
                                        63.0 * (Ep * Ep) * (gg_shape * gg_shape);
                cw[L4_C9 * 32 + lane] = r * r / (2.0 + 2.0 * r);
                cw[L4_C10 * 32 + lane] = 0.25 * kMe2 / Ep;
                cw[L4_I2EP * 32 + lane] = 0.5 / Ep;
                cw[L4_ACTIVE * 32 + lane] = active ? 1.0 : 0.0;
                // First temperature of the tile that the shared grid serves, per integral: the regime predicates (the
                // expressions kernels_wien_kernel evaluates on the same numbers) are monotone in T, which ascends with t
                // -- the Wien kernel's own range search relies on the same fact -- so a bisection replaces 30 predicate
                // evaluations per lane in the store loops (they were 9 % of the kernel's stall samples).
                int ton[3];
#pragma unroll
                for (int kind = 0; kind < 3; ++kind) {
                    const double xa = kind == 0 ? L.xa9 : (kind == 1 ? L.xa10 : L.xa11);
                    const double xb = kind == 0 ? L.xb9 : (kind == 1 ? L.xb10 : L.xb11);
                    const bool has = active && (kind == KIND_PC_ELECTRON ? xb > xa : xa > 0.0);
                    const bool reg = has && (deep || xa >= x_bot);
                    int lo = 0, hi = ntile;           // first t in [0, ntile] with the predicate true
                    while (lo < hi) {
                        const int mid = (lo + hi) >> 1;
                        const bool on = fmin(xb, sT[4 * kTT + mid]) > xa && xa <= sT[5 * kTT + mid] &&
                                        (kind != KIND_PC_ELECTRON || !(Ep < sT[6 * kTT + mid]));
                        if (on) hi = mid; else lo = mid + 1;
                    }
                    ton[kind] = reg ? lo : kTT;
                }
                cw[L4_TON * 32 + lane] = (double)(ton[0] | (ton[1] << 8) | (ton[2] << 16));
            }
            __syncwarp();
            // ---- phase B: four lanes per pair, 8 pairs at a time
#pragma unroll 1
            for (int sub = 0; sub < kLP; ++sub) {
                const int pp = sub * kL4PW + pw;
                const int64_t loc = loc0 + pp;
                const bool active = cw[L4_ACTIVE * 32 + pp] != 0.0;
                const double E = cw[L4_E * 32 + pp], Ep = cw[L4_EP * 32 + pp];
                PairConst c;
                c.E = E; c.Ep = Ep; c.dE = Ep - E;
                c.c9 = cw[L4_C9 * 32 + pp]; c.c10 = cw[L4_C10 * 32 + pp]; c.i2Ep = cw[L4_I2EP * 32 + pp];
#if ACRO_FQ_LOGTAB
                c.lt = s_lt;
#endif
                const double pre_ic = 2.0 * kPi * a2 / (Ep * Ep);
                double acc[kL4T];
#pragma unroll 1
                for (int kind = 0; kind < 3; ++kind) {
                    const double xa = cw[(L4_XA9 + 2 * kind) * 32 + pp], xb = cw[(L4_XB9 + 2 * kind) * 32 + pp];
                    const bool has = active && (kind == KIND_PC_ELECTRON ? xb > xa : xa > 0.0);
                    const bool act = has && xb > xa && xa <= xa_max && (kind != KIND_PC_ELECTRON || !(Ep < pc_min));
                    const int t_on = ((int)cw[L4_TON * 32 + pp] >> (8 * kind)) & 255;
                    shared_integral4(kind, c, xa, xb, g, sx, six, sy, sBq, s_modal, w, q, pw, act, acc);
                    if (!active) continue;
                    if (kind != KIND_PC_ELECTRON) {
                        // ---- e -> gamma (plane 1), e -> e (plane 4)
                        double* __restrict__ out = Kws + (kind == 0 ? 1 : 4) * np + loc;
#pragma unroll
                        for (int tt = 0; tt < kL4T; ++tt) {
                            const int t = tq + tt;
                            if (t < ntile) {
                                out[sK[t]] = t >= t_on ? pre_ic * acc[tt] : 0.0;
                            }
                        }
                    } else {
                        // ---- gamma -> e-/e+ (planes 2, 3) and gamma -> gamma (plane 0)
                        const double pre_pc = 0.25 * kPi * a2 * kMe2 / (Ep * Ep * Ep);
                        const double pp_pair = cw[L4_PP * 32 + pp], compton_g = cw[L4_CG * 32 + pp];
                        const double compton_e = cw[L4_CE * 32 + pp], bh_pair = cw[L4_BH * 32 + pp];
                        double* __restrict__ out = Kws + loc;
#pragma unroll
                        for (int tt = 0; tt < kL4T; ++tt) {
                            const int t = tq + tt;
                            if (t < ntile) {
                                const double pair = fma(sT[2 * kTT + t], bh_pair, t >= t_on ? pre_pc * acc[tt] : 0.0);
                                const int64_t o = sK[t];
                                out[3 * np + o] = pair;
                                out[2 * np + o] = fma(sT[kTT + t], compton_e, pair);
                            }
                        }
                        // plane 0 needs no accumulator: a rolled loop (one copy of exp instead of ten)
#pragma unroll 1
                        for (int t = tq; t < min(ntile, tq + kL4T); ++t) {
                            const double arg = Ep * sT[7 * kTT + t];
                            const double damp = arg > -700.0 ? fexp(arg) : 0.0;
                            out[sK[t]] = fma(pp_pair * sT[3 * kTT + t], damp, sT[kTT + t] * compton_g);
                        }
                    }
                }
            }
        }   // passes
    }
}

}  // namespace acro
