// acro_tmma.cuh -- the production kernel builder for the e -> e and gamma -> e thermal kernel integrals: temperature-shared
// evaluation with the temperature accumulation on the FP64 tensor-core path (DMMA, mma.sync m8n8k4 f64).  (The e -> gamma
// integral needs no quadrature: acro_icphoton.cuh.)
//
// Panel grid, limits, Bose table and partial-panel machinery: acro_tshared.cuh.  For one model point
//
//        I[pair][T] = sum_node  W[pair][node] * B[node][T]
//
// with W = weight x f(E,E',x_node) (zero outside the pair's range; Legendre-folded weights in the panels cut by a limit) and
// B = the Bose table of the temperature tile in shared memory: a dense FP64 contraction of shape (pairs x 12) x (12 x 40) per
// panel.  On B200 DMMA and DFMA share one pipe (tools/microbench/fp64_pipes.cu: 37.0 vs 34.1 TFLOP/s alone, 33-34 mixed), so
// the tensor path adds no flops -- but one DMMA.8x8x4 replaces 8 DFMA issue slots and, with the fragment layout below, the
// five 16-byte Bose loads + one weight load per ten FMAs of round 1's FMA form become one 8-byte load per DMMA; the weights
// never pass through shared memory at all:
//
//   lane = 4 g + q  (g = 0..7, q = 0..3) works for pair g of the warp's current 8 pairs (one M-tile; two M-tiles per warp were
//                                    measured: spills, slower)
//   A fragment (8 pairs x 4 nodes):  lane holds W[pair g][node q + 4 ks]  -> the lane EVALUATES nodes q, q+4, q+8 of a panel
//                                    for its pair and the values are the A fragments of the three k-steps as they are;
//   B fragment (4 nodes x 8 T):      lane loads B[node q + 4 ks][T = 8 nt + g] from the shared-memory table;
//   C fragment (8 pairs x 8 T):      lane accumulates I[pair g][T = 8 nt + 2q + {0,1}]: 5 N-tiles = 10 doubles.
//
// Stores: the 8 pairs of a sub-pass are neighbours in the packed triangle: 64 contiguous bytes per temperature (the K planes
// are padded so that every system starts on a 32-byte boundary: sector-aligned, no partial-sector read-modify-write in DRAM).
#pragma once
#include "acro_tshared.cuh"

namespace acro {

#ifndef ACRO_MM_THREADS
#define ACRO_MM_THREADS 512
#endif
constexpr int kMmThreads = ACRO_MM_THREADS;
constexpr int kMmWarps = kMmThreads / 32;
constexpr int kMmPairs = kMmWarps * 32;       // pairs per macro-pass of a block (32 per warp: phase A has lane = pair)
#ifndef ACRO_MM_PASSES
#define ACRO_MM_PASSES 48                     // macro-passes (32 pairs per warp each) per block and Bose-table fill
#endif
constexpr int kMmPassesPerBlock = ACRO_MM_PASSES;
constexpr int kMmNT = kTT / 8;                // N-tiles (8 temperatures each)
constexpr int kMmKS = kPanN / 4;              // k-steps (4 nodes each)
static_assert(kTT % 8 == 0 && kPanN % 4 == 0, "tile and panel must split into DMMA 8x8x4 shapes");

// per-pair constants of a warp's current 32 pairs in shared memory, [field][32]: energies, the limits of THIS launch's integral
// and their logs, the photon kernel's c9, and MM_FLAGS = first served temperature of the tile | active << 8
enum { MM_E = 0, MM_EP, MM_XA, MM_XB, MM_LA, MM_LB, MM_C9, MM_FLAGS, MM_NCONST };

#ifndef ACRO_MM_ASM_VOLATILE
#define ACRO_MM_ASM_VOLATILE 1
#endif
#if ACRO_MM_ASM_VOLATILE
#define ACRO_MM_ASM asm volatile
#else
#define ACRO_MM_ASM asm
#endif
#ifndef ACRO_MM_SWIZZLE
#define ACRO_MM_SWIZZLE 1                     // Bose-table columns rotated by 4 within their 8-group for nodes 2, 3 (mod 4)
#endif
__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
    ACRO_MM_ASM("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// integrand without the Bose factor at a grid node (x, 1/x, y = log x); see integrand_core (acro_fastquad.cuh)
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
__device__ __forceinline__ void add_moments_q(int kind, const PairConst& c, double x_a, double lna, double lo_y, double hi_y,
                                              double p_lo, double h, int q, double mu[kPanN]) {
    const double hp = 0.5 * (hi_y - lo_y), mp = 0.5 * (hi_y + lo_y);
    const double ic = 2.0 / h, mid = p_lo + 0.5 * h;
#ifndef ACRO_MM_MOM_UNROLL
#define ACRO_MM_MOM_UNROLL 1          // sub-interval nodes evaluated together (independent chains) in the partial panels
#endif
ACRO_UNROLL(ACRO_MM_MOM_UNROLL)
    for (int ss = 0; ss < kMmKS; ++ss) {
        const int s = q + 4 * ss;
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
// really have a partial panel here).  Returns this lane's three node weights (nodes q, q+4, q+8): the A fragments.
__device__ __forceinline__ void partial_panel_q(int kind, const PairConst& c, double x_a, double lna, double lo_y, double hi_y,
                                                double p_lo, double h, double y_layer, int q, bool on,
                                                const double* __restrict__ s_modal, double w[kMmKS]) {
    double mu[kPanN];
#pragma unroll
    for (int j = 0; j < kPanN; ++j) mu[j] = 0.0;
    double from = lo_y;
#pragma unroll 1
    for (int piece = 0; piece < 3; ++piece) {           // graded cuts at y_layer - 1/4 and y_layer - 1/16, then the rest
        const double cut = piece == 0 ? y_layer - 0.25 : (piece == 1 ? y_layer - 0.0625 : hi_y);
        const bool take = on && (piece == 2 || (cut > from && cut < hi_y));
        if (take) { add_moments_q(kind, c, x_a, lna, from, cut, p_lo, h, q, mu); from = cut; }
    }
    __syncwarp();
#pragma unroll
    for (int j = 0; j < kPanN; ++j) {
        mu[j] += __shfl_xor_sync(0xffffffffu, mu[j], 1);
        mu[j] += __shfl_xor_sync(0xffffffffu, mu[j], 2);
    }
#pragma unroll
    for (int mm = 0; mm < kMmKS; ++mm) {
        const int m = q + 4 * mm;
        double v = 0.0;
#pragma unroll
        for (int j = 0; j < kPanN; ++j) v = fma(s_modal[j * kPanN + m], mu[j], v);
        w[mm] = on ? v : 0.0;
    }
}

// Pair-only closed forms (Compton shapes per unit electron density, Bethe-Heitler dsigma/dE for Z^2 = 1; cascade.py:383-402,
// 550-558, 621-639) of every (E,E') pair of every point, thread per (point, pair), into the workspace's closed-form block
// closed[c * n_pairs_points + pt_pair_off[pt] + loc]: ~1000 instructions that neither kernels_mma_kernel (instruction cache)
// nor kernels_wien_kernel should carry in their loops.
__global__ void __launch_bounds__(128) pair_closed_forms_kernel(acro_batch_t b, double* __restrict__ closed) {
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
    const int64_t o = b.pt_pair_off[pt] + loc, npp = b.n_pairs_points;
    closed[o] = kernel_compton_core(E, Ep, 1.0);
    closed[npp + o] = kernel_compton_core(Ep + kMe - E, Ep, 1.0);
    closed[2 * npp + o] = !(Ep < E + kMe) ? bh_dsdE_Z2(E, Ep) : 0.0;        // the condition of pair_limits(): xa9 > 0
}

// One launch per integral kind (KIND = KIND_IC_PHOTON -> plane 1, KIND_IC_ELECTRON -> plane 4, KIND_PC_ELECTRON -> planes 3, 2
// and the closed-form plane 0).  Three specialised kernels instead of one that loops over the kinds: the integrand is a
// compile-time choice (no dispatch in the node loops, a third of the code per kernel), the pair constants shrink to 8 fields,
// and three launches together were 10 % faster than the fused kernel although each fills its own Bose table (21.8 against
// 24.4 ms per 1000-point slice; with 2 passes per block, the fused kernel's optimum, the split form LOSES).  Production
// launches KIND_IC_ELECTRON and KIND_PC_ELECTRON; the KIND_IC_PHOTON instantiation is not launched any more.
template <int KIND>
__global__ void __launch_bounds__(kMmThreads, 1) kernels_mma_kernel(acro_batch_t b, double* __restrict__ Kws,
                                                                    const double* __restrict__ closed, DeviceTables tb,
                                                                    acro_params_t p, int passes_per_block) {
    extern __shared__ double sm[];
    const int pt = blockIdx.y;
    const int ne = b.pt_ne[pt], nt = b.pt_nt[pt], s0 = b.pt_sys_off[pt];
    const int64_t tri = (int64_t)ne * (ne + 1) / 2;
    const int passes = (int)((tri + kMmPairs - 1) / kMmPairs);
    const int pass0 = blockIdx.x * passes_per_block, pass_end = min(passes, pass0 + passes_per_block);
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
    double* s_c = s_modal + kPanN * kPanN;                                 // [warps][MM_NCONST][32]
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
    const int q = lane & 3, gq = lane >> 2;                 // node / temperature-pair index, pair-row index of the fragments
    double* cw = s_c + (size_t)warp * MM_NCONST * 32;
    const double x_bot = exp(g.y_bot);
    const bool deep = x_bot <= 2e-11 * Tmin;
    const int64_t np = b.n_pairs_total;
    const int64_t c_off = b.pt_pair_off[pt], npp = b.n_pairs_points;
    const double a2 = kAlpha * kAlpha;
    constexpr bool kEl = KIND == KIND_IC_ELECTRON;

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
            // B-fragment loads of a half-warp touch nodes q = 0..3 (row stride kTT = 40 doubles = 8 mod 16 banks-of-8-bytes)
            // and temperatures g .. g+3: rows q and q+2 would collide; rotating the columns of rows 2, 3 (mod 4) by 4 within
            // their group of 8 makes the 16 addresses of a half-warp fall into 16 different 8-byte banks
            sB[ACRO_MM_SWIZZLE ? k * kTT + ((t & ~7) | ((t + ((k & 2) << 1)) & 7)) : idx] = v;
        }
        __syncthreads();
        const int ntile = min(kTT, nt - t0);
        const double xa_max = sT[5 * kTT + ntile - 1];
        const double pc_min = sT[6 * kTT + ntile - 1];
        const double* __restrict__ sBq = sB + (ACRO_MM_SWIZZLE ? ((gq + ((q & 2) << 1)) & 7) : gq);   // B fragment: temperature 8 nt + g
        for (int pass = pass0; pass < pass_end; ++pass) {
            const int64_t loc0 = (int64_t)pass * kMmPairs + warp * 32;
            if (loc0 >= tri) continue;              // warp-uniform: nothing left for this warp
            // ---- phase A: lane = pair: decode, this integral's limits and their logs, first served temperature
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
                const double xa = KIND == KIND_IC_PHOTON ? L.xa9 : (kEl ? L.xa10 : L.xa11);
                const double xb = KIND == KIND_IC_PHOTON ? L.xb9 : (kEl ? L.xb10 : L.xb11);
                cw[MM_E * 32 + lane] = E; cw[MM_EP * 32 + lane] = Ep;
                cw[MM_XA * 32 + lane] = xa; cw[MM_XB * 32 + lane] = xb;
                cw[MM_LA * 32 + lane] = flog(fmax(xa, 1e-300));      // same calls as the other forms of the shared kernel
                cw[MM_LB * 32 + lane] = flog(fmax(xb, 1e-300));
                if (KIND == KIND_IC_PHOTON) {
                    const double r = E / fmax(Ep - E, 1e-300);
                    cw[MM_C9 * 32 + lane] = r * r / (2.0 + 2.0 * r);
                }
                // First temperature of the tile that the shared grid serves: the regime predicates (the expressions
                // kernels_wien_kernel evaluates on the same numbers) are monotone in T, which ascends with t.
                const bool has = active && (KIND == KIND_PC_ELECTRON ? xb > xa : xa > 0.0);
                const bool reg = has && (deep || xa >= x_bot);
                int lo = 0, hi = ntile;
                while (lo < hi) {
                    const int mid = (lo + hi) >> 1;
                    const bool on = fmin(xb, sT[4 * kTT + mid]) > xa && xa <= sT[5 * kTT + mid] &&
                                    (KIND != KIND_PC_ELECTRON || !(Ep < sT[6 * kTT + mid]));
                    if (on) hi = mid; else lo = mid + 1;
                }
                cw[MM_FLAGS * 32 + lane] = (double)((reg ? lo : kTT) | ((active ? 1 : 0) << 8));
            }
            __syncwarp();
            // ---- phase B: four sub-passes of 8 pairs; lane (g, q) works for pair g of the sub-pass
#pragma unroll 1
            for (int sub = 0; sub < 4; ++sub) {
                const int pp = sub * 8 + gq;
                const int64_t loc = loc0 + pp;
                const int fl = (int)cw[MM_FLAGS * 32 + pp];
                const bool active = (fl >> 8) & 1;
                const int ton = fl & 255;
                const bool st_ok = loc < tri;
                const double E = cw[MM_E * 32 + pp], Ep = cw[MM_EP * 32 + pp];
                const double xa = cw[MM_XA * 32 + pp], xb = cw[MM_XB * 32 + pp];
                const double lna = cw[MM_LA * 32 + pp], lnb = cw[MM_LB * 32 + pp];
                double acc[kMmNT][2];
#pragma unroll
                for (int n = 0; n < kMmNT; ++n) { acc[n][0] = 0.0; acc[n][1] = 0.0; }
                // walk parameters: first (lowest) and last (highest) panel touched, which of them are cut by a limit
                int p_first = -1, p_last = 0;
                bool part_lo = false, part_hi = false;
                {
                    const bool has = active && (KIND == KIND_PC_ELECTRON ? xb > xa : xa > 0.0);
                    const bool act = has && xb > xa && xa <= xa_max && (KIND != KIND_PC_ELECTRON || !(Ep < pc_min));
                    if (act) {
                        if (lna <= g.y_bot) p_first = g.n_pan - 1;
                        else { p_first = panel_of(g, lna); part_lo = true; }
                        if (lnb >= g.y_top) p_last = 0;
                        else { p_last = panel_of(g, lnb); part_hi = true; }
                        if (fmin(lnb, g.y_top) <= fmax(lna, g.y_bot)) p_first = -1;
                    }
                }
                PairConst c;
                c.E = E; c.Ep = Ep; c.dE = Ep - E;
                c.c9 = KIND == KIND_IC_PHOTON ? cw[MM_C9 * 32 + pp] : 0.0;
                {
                    const double iEp = frcp(Ep);
                    c.c10 = 0.25 * kMe2 * iEp; c.i2Ep = 0.5 * iEp;
                }
#if ACRO_FQ_LOGTAB
                c.lt = s_lt;
#endif
#pragma unroll 1
                for (int pn = __reduce_max_sync(0xffffffffu, p_first); pn >= 0; --pn) {
                    double lo, h;
                    panel_geom(g, pn, lo, h);
                    const bool in = p_first >= 0 && pn <= p_first && pn >= p_last;
                    if (!__any_sync(0xffffffffu, in)) continue;
                    const bool part = in && ((part_lo && pn == p_first) || (part_hi && pn == p_last) || (kEl && part_hi && pn == p_last + 1));
                    const bool full = in && !part;
                    const bool any_part = __any_sync(0xffffffffu, part);
                    const bool any_full = __any_sync(0xffffffffu, full);
                    const int base = pn * kPanN;
                    double w[kMmKS];          // this lane's A fragments: weights of nodes q, q+4, q+8 for its pair
#pragma unroll
                    for (int mm = 0; mm < kMmKS; ++mm) w[mm] = 0.0;
                    if (any_part) {
                        const double a_s = fmax(lna, g.y_bot), b_s = fmin(lnb, g.y_top);
                        const bool layer = kEl && part_hi;
                        const double lo_y = (part_lo && pn == p_first) ? a_s : lo;
                        const double hi_y = (part_hi && pn == p_last) ? b_s : lo + h;
                        partial_panel_q(KIND, c, xa, lna, part ? lo_y : lo, part ? hi_y : lo + h, lo, h,
                                        layer ? b_s : hi_y + 2.0, q, part, s_modal, w);
                    }
                    if (any_full) {
                        const double wscale = 0.5 * h;
#pragma unroll
                        for (int mm = 0; mm < kMmKS; ++mm) {
                            const int m = q + 4 * mm;
                            const int k = base + m;
                            const double fv = node_core(KIND, c, sx[k], six[k], sy[k], xa, lna) * (c_glw[gl_offset(kPanN) + m] * wscale);
                            if (full) w[mm] = fv;
                        }
                    }
                    // ---- the contraction over the panel's 12 nodes: 3 k-steps x 5 N-tiles of DMMA.8x8x4
#pragma unroll
                    for (int ks = 0; ks < kMmKS; ++ks) {
                        const double* __restrict__ Bk = sBq + (size_t)(base + q + 4 * ks) * kTT;
#pragma unroll
                        for (int n = 0; n < kMmNT; ++n) dmma884(acc[n][0], acc[n][1], w[ks], Bk[8 * n]);
                    }
                }
                // ---------------- epilogue: this lane holds I[pair g][t = 8 n + 2 q + e]
                if (KIND != KIND_PC_ELECTRON) {
                    // ---- e -> gamma (plane 1), e -> e (plane 4)
                    const double pre = 2.0 * kPi * a2 / (Ep * Ep);
                    double* __restrict__ out = Kws + (KIND == KIND_IC_PHOTON ? 1 : 4) * np + loc;
#pragma unroll
                    for (int n = 0; n < kMmNT; ++n)
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            const int t = 8 * n + 2 * q + e;
                            if (st_ok && t < ntile) out[sK[t]] = (active && t >= ton) ? pre * acc[n][e] : 0.0;
                        }
                } else {
                    // ---- gamma -> e-/e+ (planes 2, 3) and gamma -> gamma (plane 0); the pair's closed forms (Compton shapes,
                    // Bethe-Heitler) come from the workspace block of pair_closed_forms_kernel
                    const double pre = 0.25 * kPi * a2 * kMe2 / (Ep * Ep * Ep);
                    const int64_t oc = c_off + (st_ok ? loc : 0);
                    const double cg = closed[oc], ce = closed[npp + oc], bh = closed[2 * npp + oc];
                    double* __restrict__ out = Kws + loc;
#pragma unroll
                    for (int n = 0; n < kMmNT; ++n)
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            const int t = 8 * n + 2 * q + e;
                            if (st_ok && t < ntile) {
                                const double pr = fma(sT[2 * kTT + t], bh, (active && t >= ton) ? pre * acc[n][e] : 0.0);
                                const int64_t o = sK[t];
                                out[3 * np + o] = pr;
                                out[2 * np + o] = fma(sT[kTT + t], ce, pr);
                            }
                        }
                    // plane 0 needs no accumulator: a rolled loop (one copy of exp instead of ten)
                    const double rr = E / Ep, gg_shape = 1.0 - rr + rr * rr;
                    const double pp_pair = 1112.0 / (10125.0 * kPi) * (a2 * a2) / ((kMe2 * kMe2) * (kMe2 * kMe2)) * 8.0 * (kPi2 * kPi2) /
                                           63.0 * (Ep * Ep) * (gg_shape * gg_shape);
#pragma unroll 1
                    for (int ne2 = 0; ne2 < 2 * kMmNT; ++ne2) {
                        const int t = 8 * (ne2 >> 1) + 2 * q + (ne2 & 1);
                        if (st_ok && t < ntile) {
                            const double arg = Ep * sT[7 * kTT + t];
                            out[sK[t]] = fma(pp_pair * sT[3 * kTT + t], arg > -700.0 ? fexp(arg) : 0.0, sT[kTT + t] * cg);
                        }
                    }
                }
            }
        }   // passes
    }
}

}  // namespace acro
