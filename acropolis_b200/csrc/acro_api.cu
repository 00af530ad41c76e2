// acro_api.cu -- the C ABI declared in include/acropolis_b200.h: handle management, workspace carving, stage
// sequencing and the host-buffer convenience entry.  No torch types anywhere.
#include <cstdio>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "acro_internal.h"

using namespace acro;

struct acro_handle {
    int device = 0;
    DeviceTables tables{};
    acro_params_t params{};
    double* d_ratedb = nullptr;
    double* d_cosmo_lT = nullptr;
    double* d_cosmo_ld = nullptr;
    std::string err;
    LaunchCounters lc;
    cudaEvent_t ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    bool ev_valid = false;
    void* host_arena = nullptr;        // device storage of acro_run_batch_host calls: grow-only, freed with the handle
    size_t host_arena_bytes = 0;
};

static thread_local std::string g_create_err;

static int fail(acro_handle* h, int code, const std::string& msg) {
    if (h) h->err = msg; else g_create_err = msg;
    return code;
}
#define CU(h, call)                                                                                     \
    do {                                                                                                \
        cudaError_t _e = (call);                                                                        \
        if (_e != cudaSuccess)                                                                          \
            return fail((h), ACRO_ECUDA, std::string(#call) + ": " + cudaGetErrorString(_e));          \
    } while (0)

extern "C" {

int acro_abi_version(void) { return ACRO_ABI_VERSION; }

void acro_default_params(acro_params_t* p) {
    if (!p) return;
    p->quad_order = 64;
    p->pdi_cell_order = 8;
    p->Emin = 1.5;
    p->approx_zero = 1e-200;
    p->Ephb_T_max = 200.0;
    p->E_EC_max = 10.0;
    p->use_db = 1;
    p->kernel_mode = ACRO_KERNELS_FAST;
    p->wien_from = 7.38905609893065;
    p->universal = 0;
}

static int check_params(acro_handle* h, const acro_params_t* p) {
    if (!(p->quad_order == 16 || p->quad_order == 32 || p->quad_order == 64 || p->quad_order == 96))
        return fail(h, ACRO_EINVAL, "quad_order must be 16, 32, 64 or 96");
    if (!(p->pdi_cell_order == 8 || p->pdi_cell_order == 16)) return fail(h, ACRO_EINVAL, "pdi_cell_order must be 8 or 16");
    if (!(p->kernel_mode == ACRO_KERNELS_FAST || p->kernel_mode == ACRO_KERNELS_PLAIN_GL || p->kernel_mode == ACRO_KERNELS_PER_T)) return fail(h, ACRO_EINVAL, "bad kernel_mode");
#ifndef ACRO_CROSSCHECK
    if (p->kernel_mode != ACRO_KERNELS_FAST)
        return fail(h, ACRO_EINVAL, "kernel_mode PLAIN_GL / PER_T exist only in the cross-check build (libacropolis_b200_xcheck.so, -DACRO_CROSSCHECK)");
#endif
    if (!(p->wien_from >= 2.718281828459045 && p->wien_from <= 7.38905609893066)) return fail(h, ACRO_EINVAL, "wien_from must lie in [e, e^2]");
    if (!(p->Emin > 0 && p->approx_zero > 0 && p->Ephb_T_max > 0 && p->E_EC_max > 0)) return fail(h, ACRO_EINVAL, "bad params");
    return ACRO_OK;
}

int acro_create(acro_handle_t** out, int device, const acro_tables_t* t, const acro_params_t* params) {
    if (!out || !t || !t->ratedb || !t->cosmo_log10_T || !t->cosmo_log10_abs_dTdt || !t->Y0)
        return fail(nullptr, ACRO_EINVAL, "acro_create: null argument");
    if (t->ratedb_nT < 2 || t->ratedb_nE < 2 || t->n_cosmo < 2) return fail(nullptr, ACRO_EINVAL, "acro_create: table too small");
    acro_handle* h = new (std::nothrow) acro_handle();
    if (!h) return fail(nullptr, ACRO_ENOMEM, "out of host memory");
    h->device = device;
    acro_default_params(&h->params);
    if (params) {
        int rc = check_params(nullptr, params);
        if (rc) { delete h; return rc; }
        h->params = *params;
    }
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) { delete h; return fail(nullptr, ACRO_ECUDA, std::string("cudaSetDevice: ") + cudaGetErrorString(e)); }
    const size_t nb_db = sizeof(double) * (size_t)t->ratedb_nT * t->ratedb_nE * 2, nb_c = sizeof(double) * (size_t)t->n_cosmo;
    if ((e = cudaMalloc(&h->d_ratedb, nb_db)) != cudaSuccess || (e = cudaMalloc(&h->d_cosmo_lT, nb_c)) != cudaSuccess ||
        (e = cudaMalloc(&h->d_cosmo_ld, nb_c)) != cudaSuccess ||
        (e = cudaMemcpy(h->d_ratedb, t->ratedb, nb_db, cudaMemcpyHostToDevice)) != cudaSuccess ||
        (e = cudaMemcpy(h->d_cosmo_lT, t->cosmo_log10_T, nb_c, cudaMemcpyHostToDevice)) != cudaSuccess ||
        (e = cudaMemcpy(h->d_cosmo_ld, t->cosmo_log10_abs_dTdt, nb_c, cudaMemcpyHostToDevice)) != cudaSuccess ||
        (e = upload_quadrature_tables()) != cudaSuccess) {
        std::string m = std::string("acro_create: ") + cudaGetErrorString(e);
        acro_destroy(h);
        return fail(nullptr, ACRO_ECUDA, m);
    }
    DeviceTables& d = h->tables;
    d.ratedb = h->d_ratedb; d.nT = t->ratedb_nT; d.nE = t->ratedb_nE;
    d.Elog_min = t->ratedb_Elog_min; d.Elog_max = t->ratedb_Elog_max; d.Tlog_min = t->ratedb_Tlog_min; d.Tlog_max = t->ratedb_Tlog_max;
    d.cosmo_lT = h->d_cosmo_lT; d.cosmo_ldTdt = h->d_cosmo_ld; d.n_cosmo = t->n_cosmo;
    d.eta = t->eta; d.Yp = t->Y0[1 * 3 + 0]; d.YHe4 = t->Y0[5 * 3 + 0];
    for (int i = 0; i < 27; ++i) d.Y0[i] = t->Y0[i];
    for (int i = 0; i < 5; ++i) cudaEventCreate(&h->ev[i]);
    cudaEventCreate(&h->lc.split[0]); cudaEventCreate(&h->lc.split[1]); cudaEventCreate(&h->lc.split[2]);
    *out = h;
    return ACRO_OK;
}

int acro_destroy(acro_handle_t* h) {
    if (!h) return ACRO_OK;
    cudaSetDevice(h->device);
    cudaFree(h->d_ratedb); cudaFree(h->d_cosmo_lT); cudaFree(h->d_cosmo_ld);
    if (h->host_arena) cudaFree(h->host_arena);
    for (int i = 0; i < 5; ++i) if (h->ev[i]) cudaEventDestroy(h->ev[i]);
    for (int i = 0; i < 3; ++i) if (h->lc.split[i]) cudaEventDestroy(h->lc.split[i]);
    delete h;
    return ACRO_OK;
}

const char* acro_last_error(const acro_handle_t* h) { return h ? h->err.c_str() : g_create_err.c_str(); }

int acro_set_params(acro_handle_t* h, const acro_params_t* p) {
    if (!h || !p) return ACRO_EINVAL;
    int rc = check_params(h, p);
    if (rc) return rc;
    h->params = *p;
    return ACRO_OK;
}

// workspace layout: [5 K planes][pair closed forms][direct-rate items][cross-section table of the pdi cells][item counter
// (256 B)][per-system status]
static inline int64_t align256(int64_t v) { return (v + 255) & ~(int64_t)255; }
struct WsLayout { int64_t k, closed, items, sig, counter, status, total; };
static WsLayout ws_layout(const acro_batch_t* b) {
    WsLayout w;
    w.k = 0;
    w.closed = align256((int64_t)ACRO_NK * b->n_pairs_total * (int64_t)sizeof(double));
    w.items = w.closed + align256((int64_t)3 * b->n_pairs_points * (int64_t)sizeof(double));
    w.sig = w.items + align256(b->n_sys_energy_total * (int64_t)sizeof(DirectRateItem));
    w.counter = w.sig + align256((int64_t)16 * ACRO_NR * b->n_energy_total * (int64_t)sizeof(double));   // up to 16 nodes per cell
    w.status = w.counter + 256;
    w.total = w.status + align256((int64_t)b->n_systems * (int64_t)sizeof(int));
    return w;
}

int64_t acro_workspace_bytes(const acro_handle_t* h, const acro_batch_t* b) {
    if (!h || !b) return 0;
    return ws_layout(b).total;
}

int acro_run_batch(acro_handle_t* h, const acro_batch_t* b, const acro_out_t* o, void* workspace, int64_t workspace_bytes,
                   int32_t stages, void* stream) {
    if (!h || !b || !o) return ACRO_EINVAL;
    if (b->n_points < 0 || b->n_systems < 0) return fail(h, ACRO_EINVAL, "negative counts");
    if (b->n_points > 65535) return fail(h, ACRO_EINVAL, "n_points > 65535: split the batch (points are indexed by blockIdx.y)");
    if (b->n_points == 0 || b->n_systems == 0) return ACRO_OK;
    if (!b->pt_ne || !b->pt_nt || !b->pt_e_off || !b->pt_sys_off || !b->pt_e0 || !b->pt_t_at_tmax || !b->e_grid ||
        !b->sys_point || !b->sys_T || !b->sys_f_off || !b->sys_k_off || !b->s0 || !b->pt_pair_off)
        return fail(h, ACRO_EINVAL, "acro_run_batch: null batch array");
    if ((b->n_pairs_total & 3) != 0) return fail(h, ACRO_EINVAL, "n_pairs_total must be a multiple of 4 (every system's plane block is padded to 4 entries)");
    if (b->sc_mode != ACRO_SC_NONE && !b->sc) return fail(h, ACRO_EINVAL, "sc_mode set but sc is NULL");
    if (b->sc_mode == ACRO_SC_SEPARABLE && !b->sc_E) return fail(h, ACRO_EINVAL, "separable sc needs sc_E");
    if (b->ne_max < 2) return fail(h, ACRO_EINVAL, "ne_max < 2");
    const WsLayout w = ws_layout(b);
    if (!workspace || workspace_bytes < w.total) return fail(h, ACRO_EINVAL, "workspace too small");
    if ((stages & (ACRO_STAGE_RATES | ACRO_STAGE_SOLVE)) && !o->G) return fail(h, ACRO_EINVAL, "out.G is required");
    if ((stages & ACRO_STAGE_MATRIX) && !o->pdi) return fail(h, ACRO_EINVAL, "out.pdi is required for the matrix stage");
    CU(h, cudaSetDevice(h->device));
    cudaStream_t st = (cudaStream_t)stream;
    char* base = (char*)workspace;
    double* Kws = (double*)(base + w.k);
    DirectRateItem* items = (DirectRateItem*)(base + w.items);
    int* counter = (int*)(base + w.counter);
    int* status_sys = (int*)(base + w.status);
    CU(h, cudaEventRecord(h->ev[0], st));
    if (stages & ACRO_STAGE_RATES) {
        CU(h, cudaMemsetAsync(base + w.counter, 0, (size_t)(w.total - w.counter), st));
        CU(h, launch_rates(*b, *o, h->tables, h->params, items, counter, status_sys, st, h->lc));
        CU(h, launch_direct_rates(h->tables, h->params, o->G, items, counter, b->n_sys_energy_total, st, h->lc));
    }
    CU(h, cudaEventRecord(h->ev[1], st));
    if ((stages & ACRO_STAGE_KERNELS) && !h->params.universal)
        CU(h, launch_kernels(*b, h->tables, h->params, Kws, (double*)(base + w.closed), st, h->lc));
    CU(h, cudaEventRecord(h->ev[2], st));
    if (stages & ACRO_STAGE_SOLVE) CU(h, launch_solve_pdi(*b, *o, h->tables, h->params, Kws, (double*)(base + w.sig), st, h->lc));
    CU(h, cudaEventRecord(h->ev[3], st));
    if (stages & ACRO_STAGE_MATRIX)
        CU(h, launch_matrix(b->n_points, b->pt_nt, b->pt_sys_off, b->sys_T, o->pdi, nullptr, b->pt_t_at_tmax,
                            (stages & ACRO_STAGE_RATES) ? status_sys : nullptr, b->pt_e0, b->nt_max, o->mpdi, o->mdcy, o->Yf,
                            o->status, h->tables, h->params, st, h->lc));
    CU(h, cudaEventRecord(h->ev[4], st));
    h->ev_valid = true;
    return ACRO_OK;
}

int acro_count_work(acro_handle_t* h, const acro_batch_t* b, int64_t out[4], void* stream) {
    if (!h || !b || !out) return ACRO_EINVAL;
    CU(h, cudaSetDevice(h->device));
    cudaStream_t st = (cudaStream_t)stream;
    unsigned long long* d = nullptr;
    CU(h, cudaMalloc(&d, 4 * sizeof(unsigned long long)));
    cudaError_t e = cudaMemsetAsync(d, 0, 4 * sizeof(unsigned long long), st);
    if (e == cudaSuccess) e = launch_count_work(*b, h->params, d, st, h->lc);
    unsigned long long hv[4] = {0, 0, 0, 0};
    if (e == cudaSuccess) e = cudaMemcpyAsync(hv, d, sizeof(hv), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    cudaFree(d);
    if (e != cudaSuccess) return fail(h, ACRO_ECUDA, std::string("acro_count_work: ") + cudaGetErrorString(e));
    for (int i = 0; i < 4; ++i) out[i] = (int64_t)hv[i];
    return ACRO_OK;
}

int acro_last_stage_ms(acro_handle_t* h, float ms[4]) {
    if (!h || !ms) return ACRO_EINVAL;
    if (!h->ev_valid) return fail(h, ACRO_EINVAL, "no batch has been run");
    CU(h, cudaEventSynchronize(h->ev[4]));
    for (int i = 0; i < 4; ++i) CU(h, cudaEventElapsedTime(&ms[i], h->ev[i], h->ev[i + 1]));
    return ACRO_OK;
}

int acro_last_kernel_split_ms(acro_handle_t* h, float ms[4]) {
    if (!h || !ms) return ACRO_EINVAL;
    if (!h->ev_valid || !h->lc.split_valid) return fail(h, ACRO_EINVAL, "the production kernel builder has not run on this handle");
    CU(h, cudaEventSynchronize(h->ev[2]));
    CU(h, cudaEventElapsedTime(&ms[0], h->ev[1], h->lc.split[0]));
    CU(h, cudaEventElapsedTime(&ms[1], h->lc.split[0], h->lc.split[1]));
    CU(h, cudaEventElapsedTime(&ms[2], h->lc.split[1], h->lc.split[2]));
    CU(h, cudaEventElapsedTime(&ms[3], h->lc.split[2], h->ev[2]));
    return ACRO_OK;
}

int64_t acro_launch_count(const acro_handle_t* h) { return h ? h->lc.launches : 0; }

namespace {
struct DevBuf {
    void* p = nullptr;
    ~DevBuf() { if (p) cudaFree(p); }
    cudaError_t alloc(size_t n) { return cudaMalloc(&p, n ? n : 8); }
};

// bump allocator over the handle's host-path arena (256-byte aligned pieces)
struct Carver {
    char* base;
    size_t off = 0;
    explicit Carver(char* b) : base(b) {}
    static size_t al(size_t n) { return (n + 255) & ~(size_t)255; }
    void* take(size_t n) { void* p = base ? base + off : nullptr; off += al(n ? n : 8); return p; }
};
}  // namespace

// Host-pointer entry.  All device storage of a call (batch arrays, outputs, workspace) is carved out of ONE arena that the
// handle keeps and only ever grows, so a loop over chunks pays cudaMalloc once (it used to allocate and
// free ~25 buffers per call).  A batch whose workspace does not fit the device is refused with the size in the message:
// split it (acro_workspace_bytes tells how large a sub-batch is; acropolis_b200/engine.py::Engine.chunks is the host-side
// splitter the Python layer uses).
int acro_run_batch_host(acro_handle_t* h, const acro_batch_t* hb, const acro_out_t* ho, int32_t stages) {
    if (!h || !hb || !ho) return ACRO_EINVAL;
    if (hb->n_points <= 0 || hb->n_systems <= 0) return ACRO_OK;
    if (!hb->pt_pair_off) return fail(h, ACRO_EINVAL, "acro_run_batch_host: pt_pair_off is NULL");
    CU(h, cudaSetDevice(h->device));
    cudaStream_t st = nullptr;
    const size_t P = hb->n_points, S = hb->n_systems, NEt = hb->n_energy_total, NF = hb->n_sys_energy_total;
    const int64_t wsb = acro_workspace_bytes(h, hb);
    const size_t sc_bytes = hb->sc_mode == ACRO_SC_SEPARABLE ? S * 3 * 8 : (hb->sc_mode == ACRO_SC_DENSE ? NF * 3 * 8 : 0);
    const size_t scE_bytes = hb->sc_mode == ACRO_SC_SEPARABLE ? NEt * 3 * 8 : 0;
    acro_batch_t db = *hb;
    acro_out_t dout{};
    void *dF = nullptr, *ws = nullptr;
    // two passes over the same carving sequence: sizes first, pointers second
    for (int pass = 0; pass < 2; ++pass) {
        Carver c(pass ? (char*)h->host_arena : nullptr);
        db.pt_ne = (const int32_t*)c.take(P * 4);          db.pt_nt = (const int32_t*)c.take(P * 4);
        db.pt_e_off = (const int64_t*)c.take(P * 8);       db.pt_sys_off = (const int32_t*)c.take(P * 4);
        db.pt_pair_off = (const int64_t*)c.take(P * 8);    db.pt_e0 = (const double*)c.take(P * 8);
        db.pt_t_at_tmax = (const double*)c.take(P * 8);    db.e_grid = (const double*)c.take(NEt * 8);
        db.sys_point = (const int32_t*)c.take(S * 4);      db.sys_T = (const double*)c.take(S * 8);
        db.sys_f_off = (const int64_t*)c.take(S * 8);      db.sys_k_off = (const int64_t*)c.take(S * 8);
        db.s0 = (const double*)c.take(S * 3 * 8);
        db.sc = sc_bytes ? (const double*)c.take(sc_bytes) : nullptr;
        db.sc_E = scE_bytes ? (const double*)c.take(scE_bytes) : nullptr;
        dout.G = (double*)c.take(NF * 3 * 8);
        dF = ho->F ? c.take(NF * 3 * 8) : nullptr;         dout.F = (double*)dF;
        dout.pdi = (double*)c.take(S * ACRO_NR * 8);       dout.mpdi = (double*)c.take(P * 81 * 8);
        dout.mdcy = (double*)c.take(P * 81 * 8);           dout.Yf = (double*)c.take(P * 27 * 8);
        dout.status = (int32_t*)c.take(P * 4);
        ws = c.take((size_t)wsb);
        if (pass == 0 && c.off > h->host_arena_bytes) {
            size_t free_b = 0, total_b = 0;
            CU(h, cudaMemGetInfo(&free_b, &total_b));
            if (c.off > free_b + h->host_arena_bytes)
                return fail(h, ACRO_ENOMEM, "acro_run_batch_host: the batch needs " + std::to_string(c.off >> 20) + " MiB of device memory (" +
                            std::to_string((free_b + h->host_arena_bytes) >> 20) + " MiB available): split it into sub-batches");
            if (h->host_arena) { cudaFree(h->host_arena); h->host_arena = nullptr; h->host_arena_bytes = 0; }
            size_t want = c.off + c.off / 4;                       // 25 % headroom so that similar batches do not regrow it
            if (want > free_b) want = c.off;
            CU(h, cudaMalloc(&h->host_arena, want));
            h->host_arena_bytes = want;
        }
    }
#define UP(field, bytes) CU(h, cudaMemcpyAsync((void*)db.field, hb->field, (bytes), cudaMemcpyHostToDevice, st))
    UP(pt_ne, P * 4); UP(pt_nt, P * 4); UP(pt_e_off, P * 8); UP(pt_sys_off, P * 4); UP(pt_pair_off, P * 8); UP(pt_e0, P * 8);
    UP(pt_t_at_tmax, P * 8); UP(e_grid, NEt * 8); UP(sys_point, S * 4); UP(sys_T, S * 8); UP(sys_f_off, S * 8);
    UP(sys_k_off, S * 8); UP(s0, S * 3 * 8);
    if (sc_bytes) UP(sc, sc_bytes);
    if (scE_bytes) UP(sc_E, scE_bytes);
#undef UP
    int rc = acro_run_batch(h, &db, &dout, ws, wsb, stages, st);
    if (rc) return rc;
    if (ho->G && (stages & ACRO_STAGE_RATES)) CU(h, cudaMemcpyAsync(ho->G, dout.G, NF * 3 * 8, cudaMemcpyDeviceToHost, st));
    if (ho->F && (stages & ACRO_STAGE_SOLVE)) CU(h, cudaMemcpyAsync(ho->F, dout.F, NF * 3 * 8, cudaMemcpyDeviceToHost, st));
    if (ho->pdi && (stages & ACRO_STAGE_SOLVE)) CU(h, cudaMemcpyAsync(ho->pdi, dout.pdi, S * ACRO_NR * 8, cudaMemcpyDeviceToHost, st));
    if (stages & ACRO_STAGE_MATRIX) {
        if (ho->mpdi) CU(h, cudaMemcpyAsync(ho->mpdi, dout.mpdi, P * 81 * 8, cudaMemcpyDeviceToHost, st));
        if (ho->mdcy) CU(h, cudaMemcpyAsync(ho->mdcy, dout.mdcy, P * 81 * 8, cudaMemcpyDeviceToHost, st));
        if (ho->Yf) CU(h, cudaMemcpyAsync(ho->Yf, dout.Yf, P * 27 * 8, cudaMemcpyDeviceToHost, st));
        if (ho->status) CU(h, cudaMemcpyAsync(ho->status, dout.status, P * 4, cudaMemcpyDeviceToHost, st));
    }
    CU(h, cudaStreamSynchronize(st));
    return ACRO_OK;
}

int acro_unpack_kernels(acro_handle_t* h, const acro_batch_t* b, const void* workspace, int32_t system, double* K_dense,
                        void* stream) {
    if (!h || !b || !workspace || !K_dense) return ACRO_EINVAL;
    if (system < 0 || system >= b->n_systems) return fail(h, ACRO_EINVAL, "system out of range");
    CU(h, cudaSetDevice(h->device));
    // sys_k_off / sys_point / pt_ne live on the device: fetch the three scalars
    int64_t k_off = 0; int32_t pt = 0, ne = 0;
    CU(h, cudaMemcpy(&k_off, b->sys_k_off + system, 8, cudaMemcpyDeviceToHost));
    CU(h, cudaMemcpy(&pt, b->sys_point + system, 4, cudaMemcpyDeviceToHost));
    CU(h, cudaMemcpy(&ne, b->pt_ne + pt, 4, cudaMemcpyDeviceToHost));
    CU(h, launch_unpack_kernels((const double*)workspace, b->n_pairs_total, k_off, ne, K_dense, (cudaStream_t)stream, h->lc));
    return ACRO_OK;
}

int acro_matp_batch(acro_handle_t* h, int32_t n_points, int32_t nt_max, const int32_t* pt_nt, const int32_t* pt_sys_off,
                    const double* sys_T, const double* pdi, const double* T_lo, const double* pt_t_at_tmax, double* mpdi,
                    double* mdcy, double* Yf, int32_t* status, void* stream) {
    if (!h || !pt_nt || !pt_sys_off || !sys_T || !pdi) return ACRO_EINVAL;
    if (Yf && !pt_t_at_tmax) return fail(h, ACRO_EINVAL, "Yf needs pt_t_at_tmax");
    if (nt_max < 2) return fail(h, ACRO_EINVAL, "nt_max < 2");
    CU(h, cudaSetDevice(h->device));
    CU(h, launch_matrix(n_points, pt_nt, pt_sys_off, sys_T, pdi, T_lo, pt_t_at_tmax, nullptr, nullptr, nt_max, mpdi, mdcy, Yf,
                        status, h->tables, h->params, (cudaStream_t)stream, h->lc));
    return ACRO_OK;
}

int acro_transfer_batch(acro_handle_t* h, int32_t n, const double* mpdi, const double* mdcy, const int32_t* matrix_index,
                        const double* factor, const double* t_at_tmax, double* expm_out, double* Yf, int32_t* status, void* stream) {
    if (!h || !mpdi || !mdcy || !t_at_tmax || !Yf) return ACRO_EINVAL;
    CU(h, cudaSetDevice(h->device));
    CU(h, launch_transfer(n, mpdi, mdcy, matrix_index, factor, t_at_tmax, expm_out, Yf, status, h->tables, (cudaStream_t)stream, h->lc));
    return ACRO_OK;
}

int acro_total_rates(acro_handle_t* h, int32_t n, const double* E, const double* T, double* out, void* stream) {
    if (!h || !E || !T || !out) return ACRO_EINVAL;
    if (n <= 0) return ACRO_OK;
    CU(h, cudaSetDevice(h->device));
    cudaStream_t st = (cudaStream_t)stream;
    DevBuf items, counter;
    CU(h, items.alloc(sizeof(DirectRateItem) * (size_t)n));
    CU(h, counter.alloc(256));
    CU(h, cudaMemsetAsync(counter.p, 0, 256, st));
    CU(h, launch_total_rates(n, E, T, out, h->tables, h->params, (DirectRateItem*)items.p, (int*)counter.p, st, h->lc));
    CU(h, cudaStreamSynchronize(st));   // the temporaries die here
    return ACRO_OK;
}

int acro_direct_rates(acro_handle_t* h, int32_t n, const double* E, const double* T, double* out, void* stream) {
    if (!h || !E || !T || !out) return ACRO_EINVAL;
    if (n <= 0) return ACRO_OK;
    CU(h, cudaSetDevice(h->device));
    cudaStream_t st = (cudaStream_t)stream;
    DevBuf items, counter;
    CU(h, items.alloc(sizeof(DirectRateItem) * (size_t)n));
    CU(h, counter.alloc(256));
    CU(h, cudaMemsetAsync(counter.p, 0, 256, st));
    CU(h, launch_direct_rate_table(n, E, T, out, h->tables, h->params, (DirectRateItem*)items.p, (int*)counter.p, st, h->lc));
    CU(h, cudaStreamSynchronize(st));   // the temporaries die here
    return ACRO_OK;
}

int acro_measure_fp64_peak(acro_handle_t* h, double* tflops) {
    if (!h || !tflops) return ACRO_EINVAL;
    CU(h, cudaSetDevice(h->device));
    CU(h, run_dfma_benchmark(tflops, h->lc));
    return ACRO_OK;
}

}  // extern "C"
