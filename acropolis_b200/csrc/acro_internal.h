// acro_internal.h -- shared between the kernel TU and the C-ABI TU.  Not part of the public interface.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>

#include "../../include/acropolis_b200.h"

namespace acro {

// Device-side view of the uploaded tables.
struct DeviceTables {
    const double* ratedb;   // [nT][nE][2]
    int nT, nE;
    double Elog_min, Elog_max, Tlog_min, Tlog_max;
    const double* cosmo_lT;     // [n_cosmo] log10 T, descending
    const double* cosmo_ldTdt;  // [n_cosmo] log10 |dT/dt|
    int n_cosmo;
    double eta, Yp, YHe4;
    double Y0[27];              // [9][3]
};

// Work item of the direct (outside rates.db) rate integrals.
struct DirectRateItem {
    int64_t g_index;   // index of G[0][i] of the system in the G array (photon); electron/positron follow at +NE, +2NE
    int32_t ne;        // NE of the system
    int32_t kind;      // bit0: photon pair creation needed, bit1: inverse Compton needed
    double  E, T;
};

struct LaunchCounters {
    int64_t launches = 0;
    cudaEvent_t split[3] = {nullptr, nullptr, nullptr};   // recorded between the launch groups of the production kernel builder
    bool split_valid = false;
};

// Kernel launchers (acro_kernels.cu).  All asynchronous on `st`; return cudaGetLastError().
cudaError_t upload_quadrature_tables();   // fills the __constant__ Gauss-Legendre tables of the current device
cudaError_t launch_rates(const acro_batch_t& b, const acro_out_t& o, const DeviceTables& t, const acro_params_t& p,
                         DirectRateItem* items, int* item_count, int* status_sys, cudaStream_t st, LaunchCounters& lc);
cudaError_t launch_direct_rates(const DeviceTables& t, const acro_params_t& p, double* G, const DirectRateItem* items,
                                const int* item_count, int64_t max_items, cudaStream_t st, LaunchCounters& lc);
cudaError_t launch_kernels(const acro_batch_t& b, const DeviceTables& t, const acro_params_t& p, double* Kws, double* closed,
                           cudaStream_t st, LaunchCounters& lc);
cudaError_t launch_count_work(const acro_batch_t& b, const acro_params_t& p, unsigned long long* out, cudaStream_t st,
                              LaunchCounters& lc);
cudaError_t launch_solve_pdi(const acro_batch_t& b, const acro_out_t& o, const DeviceTables& t, const acro_params_t& p,
                             const double* Kws, double* sig, cudaStream_t st, LaunchCounters& lc);
cudaError_t launch_matrix(int n_points, const int32_t* pt_nt, const int32_t* pt_sys_off, const double* sys_T,
                          const double* pdi, const double* T_lo, const double* pt_t_at_tmax, const int* status_sys,
                          const double* pt_e0, int nt_max, double* mpdi, double* mdcy, double* Yf, int32_t* status,
                          const DeviceTables& t, const acro_params_t& p, cudaStream_t st, LaunchCounters& lc);
cudaError_t launch_transfer(int n, const double* mpdi, const double* mdcy, const int32_t* matrix_index, const double* factor, const double* t_at_tmax,
                            double* expm_out, double* Yf, int32_t* status, const DeviceTables& t, cudaStream_t st,
                            LaunchCounters& lc);
cudaError_t launch_total_rates(int n, const double* E, const double* T, double* out, const DeviceTables& t,
                               const acro_params_t& p, DirectRateItem* items, int* item_count, cudaStream_t st,
                               LaunchCounters& lc);
cudaError_t launch_direct_rate_table(int n, const double* E, const double* T, double* out, const DeviceTables& t,
                                     const acro_params_t& p, DirectRateItem* items, int* item_count, cudaStream_t st,
                                     LaunchCounters& lc);
cudaError_t launch_unpack_kernels(const double* Kws, int64_t n_pairs_total, int64_t k_off, int ne, double* K_dense,
                                  cudaStream_t st, LaunchCounters& lc);
cudaError_t run_dfma_benchmark(double* tflops, LaunchCounters& lc);

}  // namespace acro
