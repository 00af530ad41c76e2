/*
 * acropolis_b200.h -- C ABI of the B200-native non-thermal-nucleosynthesis hot path.
 *
 * Drop-in boundary for hep-mh/acropolis v1.3.1 (reference paths relative to its repo root).  The reference has
 * no FFI of its own (it is pure Python); the seam these entry points replace is the body of
 *     acropolis/models.py:112-130   AbstractModel._pdi_matrix()
 * i.e. everything below NuclearReactor.get_pdi_grids() / MatrixGenerator.get_final_matp() / expm, batched
 * over many (model point, temperature) systems.  INTEGRATION.md shows the ctypes stub a maintainer would add.
 *
 * Conventions
 *   - plain C, no torch types; every array is IEEE float64 or int32/int64, C-contiguous.
 *   - the caller owns every buffer (device buffers for acro_run_batch, host buffers for acro_run_batch_host);
 *     the library keeps only the opaque handle that holds the uploaded tables (rates.db, cosmo table, Y0, eta).
 *   - every function returns 0 on success or a negative ACRO_E* code; acro_last_error() gives the text.
 *     Nothing calls exit() (the reference's print_error does, pprint.py:82-100); per-point problems are
 *     reported in the status bitmask instead.
 *   - re-entrant per handle; one handle per device; all work is ordered on the stream passed in.
 *   - units as in the reference: energies/temperatures MeV, rates MeV, times s.
 */
#ifndef ACROPOLIS_B200_H
#define ACROPOLIS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ACRO_ABI_VERSION 4

#define ACRO_NX 3   /* cascade species: photon, electron, positron        (cascade.py:718)        */
#define ACRO_NY 9   /* nuclides n,p,d,t,He3,He4,Li6,Li7,Be7              (nucl.py:32-42)          */
#define ACRO_NR 17  /* photodisintegration reactions                     (nucl.py:49-67)          */
#define ACRO_NK 5   /* distinct kernel planes: gg, eg, ge-, ge+, ee      (cascade.py:439-455,644-700) */

/* error codes */
#define ACRO_OK            0
#define ACRO_EINVAL       -1
#define ACRO_ECUDA        -2
#define ACRO_ENOMEM       -3

/* stage mask for acro_run_batch */
#define ACRO_STAGE_RATES     1  /* G[3,NE] per system: closed forms + rates.db interpolation + direct 2-D integrals
                                   (cascade.py:285-374,469-498; db.py:50-124)                                    */
#define ACRO_STAGE_KERNELS   2  /* the five K planes per system (cascade.py:383-455,507-598,621-700)            */
#define ACRO_STAGE_SOLVE     4  /* cascade back-substitution + 17 pdi rates (cascade.py:180-239; nucl.py:398-467) */
#define ACRO_STAGE_MATRIX    8  /* T-integral, expm, pre/post decay, Yf (nucl.py:583-623; models.py:84-157)      */
#define ACRO_STAGE_ALL      15

/* per-point status bits (the reference prints warnings / raises instead: models.py:47-61, input.py:218-221) */
#define ACRO_ST_NONFINITE     1  /* a non-finite number reached the output                                   */
#define ACRO_ST_T_OUTSIDE     2  /* temperature range not covered by the cosmo table                         */
#define ACRO_ST_DIRECT_RATE   4  /* at least one rate came from the direct 2-D integral (outside rates.db)   */
#define ACRO_ST_E0_ABOVE_GEV  8  /* injection energy > 1 GeV ("results cannot be trusted")                   */

/* how the three thermal kernel integrals are evaluated (same integrals and limits either way) */
#define ACRO_KERNELS_FAST      0  /* temperature-shared panel grid (Bose table in shared memory) + Gauss-Laguerre on the Wien
                                     tail + reduced Gauss-Legendre elsewhere (default)                                */
/* the next two are cross-check modes of the test build (libacropolis_b200_xcheck.so, -DACRO_CROSSCHECK); the product
 * library rejects them in acro_set_params */
#define ACRO_KERNELS_PLAIN_GL  1  /* plain Q-point Gauss-Legendre in log(eps_bar) with the reference's F/G incl. range tests */
#define ACRO_KERNELS_PER_T     2  /* as FAST but every (pair,T) integral evaluated on its own (no temperature sharing)    */

/* continuous-source representation */
#define ACRO_SC_NONE       0  /* SC == 0 for all species (models.py:197-206 defaults)                        */
#define ACRO_SC_SEPARABLE  1  /* SC[X](E_i,T_s) = sc[s*3+X] * sc_E[(pt_e_off+i)*3+X]   (FSR form, models.py:276-287) */
#define ACRO_SC_DENSE      2  /* SC[X](E_i,T_s) = sc[(sys_f_off[s]*3) + X*NE + i]                            */

typedef struct acro_handle acro_handle_t;

/* Tables uploaded once per handle (host pointers). */
typedef struct {
    const double* ratedb;        /* float64[n_T, n_E, 2] log10 rates; [..,0] photon pair creation, [..,1] e IC (db.py:95-110) */
    int32_t       ratedb_nT, ratedb_nE;
    double        ratedb_Elog_min, ratedb_Elog_max, ratedb_Tlog_min, ratedb_Tlog_max; /* params.py:54-56 */
    const double* cosmo_log10_T;        /* float64[n_cosmo]  log10 of cosmo_file column 1, as stored (descending T) (input.py:137) */
    const double* cosmo_log10_abs_dTdt; /* float64[n_cosmo]  log10|column 2|                                            */
    int32_t       n_cosmo;
    double        eta;           /* baryon-to-photon ratio (param_file.dat)                                 */
    const double* Y0;            /* float64[9,3] initial abundances (mean, high, low) (input.py:263-268)    */
} acro_tables_t;

/* Algorithm parameters (reference: params.py:78-128).  acro_default_params() fills the reference defaults. */
typedef struct {
    int32_t quad_order;      /* Gauss-Legendre nodes per thermal kernel integral: 16, 32, 64 (default) or 96     */
    int32_t pdi_cell_order;  /* GL nodes per energy-grid cell in the pdi-rate integrals (default 8)              */
    double  Emin;            /* 1.5 MeV                                                                           */
    double  approx_zero;     /* 1e-200                                                                            */
    double  Ephb_T_max;      /* 200                                                                               */
    double  E_EC_max;        /* 10                                                                                */
    int32_t use_db;          /* flags.usedb                                                                       */
    int32_t kernel_mode;     /* ACRO_KERNELS_FAST (default) or ACRO_KERNELS_PLAIN_GL                              */
    double  wien_from;       /* ACRO_KERNELS_FAST: a thermal integral whose lower limit x_a exceeds wien_from * T
                              * is evaluated per (E,E',T) with Gauss-Laguerre on the Wien tail, the others on the
                              * temperature-shared grid.  e <= wien_from <= e^2.  Default e^2 = 7.389 (kernel entries
                              * within 1e-7 of the converged integrals); e = 2.718 gives 1e-9 at ~1.4x the time.    */
    int32_t universal;       /* flags.universal (flags.py:33-41): the photon spectrum is the universal one of
                              * cascade.py:774-812 instead of the cascade solution, and the pdi integrals stop at
                              * min(E0, me^2/(22T)) (nucl.py:416-422); the kernel and solve work is skipped.        */
    int32_t reserved;
} acro_params_t;

/*
 * One ragged batch of model points.  P points, S = sum_p NT_p systems (a system = one point at one temperature).
 * Pointers are DEVICE pointers for acro_run_batch and HOST pointers for acro_run_batch_host.
 */
typedef struct {
    int32_t n_points;            /* P <= 65535 per call (split larger scans into batches)   */
    int32_t n_systems;           /* S                                                        */
    int64_t n_energy_total;      /* sum_p NE_p   (length of e_grid)                          */
    int64_t n_sys_energy_total;  /* sum_s NE_s   (rows of G / F / dense SC)                  */
    int64_t n_pairs_total;       /* sum_s pad4(NE_s (NE_s+1)/2): entries of one K plane; every system's packed triangle is
                                    padded to a multiple of 4 entries (32-byte sectors), so this is a multiple of 4  */
    int64_t n_pairs_points;      /* sum_p pad4(NE_p (NE_p+1)/2): entries of one pair-constant block of the workspace */
    int32_t ne_max;              /* max_p NE_p                                               */
    int32_t nt_max;              /* max_p NT_p                                               */
    int32_t sc_mode;             /* ACRO_SC_*                                                */
    int32_t reserved;
    /* per point [P] */
    const int32_t* pt_ne;        /* NE_p  (cascade.py:736-739)                               */
    const int32_t* pt_nt;        /* NT_p  (nucl.py:473)                                      */
    const int64_t* pt_e_off;     /* offset of the point's grid in e_grid                     */
    const int32_t* pt_sys_off;   /* index of the point's first system                        */
    const int64_t* pt_pair_off;  /* offset of the point's pairs in a pair-constant block: prefix sum of pad4(tri_p) */
    const double*  pt_e0;        /* injection energy E0 (delta-source position)              */
    const double*  pt_t_at_tmax; /* cosmic time [s] at T_max, for the pre-decay matrix (models.py:140-145) */
    /* per energy [n_energy_total] */
    const double*  e_grid;       /* np.logspace(log Emin, log E0, NE, base=e) per point (cascade.py:745) */
    /* per system [S] */
    const int32_t* sys_point;    /* owning point                                             */
    const double*  sys_T;        /* temperature T_s                                          */
    const int64_t* sys_f_off;    /* row offset of the system in G / F / dense SC             */
    const int64_t* sys_k_off;    /* entry offset of the system in each K plane: prefix sum of pad4(tri_s), multiple of 4 */
    const double*  s0;           /* [S,3] monochromatic source terms S0_X(T_s) (cascade.py:763) */
    const double*  sc;           /* see ACRO_SC_*; may be NULL for ACRO_SC_NONE              */
    const double*  sc_E;         /* [n_energy_total,3] for ACRO_SC_SEPARABLE, else NULL       */
} acro_batch_t;

/* Outputs.  Any pointer may be NULL (that output is then not written), except where a requested stage needs it. */
typedef struct {
    double*  G;      /* [n_sys_energy_total*3]  per system [3][NE]: total rates Gamma_X(E_i,T)  (cascade.py:751)     */
    double*  F;      /* [n_sys_energy_total*3]  per system [3][NE]: spectra F_X(E_i) (cascade.py:227-237)            */
    double*  pdi;    /* [S,17]  photodisintegration rates per system (nucl.py:439-464)                                */
    double*  mpdi;   /* [P,9,9] (nucl.py:613)                                                                         */
    double*  mdcy;   /* [P,9,9] (nucl.py:614)                                                                         */
    double*  Yf;     /* [P,9,3] final abundances (models.py:87-89)                                                    */
    int32_t* status; /* [P] ACRO_ST_* bits                                                                            */
} acro_out_t;

/* ---- handle ---------------------------------------------------------------------------------------------- */
int  acro_abi_version(void);
void acro_default_params(acro_params_t* p);
int  acro_create(acro_handle_t** h, int device, const acro_tables_t* tables, const acro_params_t* params);
int  acro_destroy(acro_handle_t* h);
const char* acro_last_error(const acro_handle_t* h);   /* h may be NULL: last error of the calling thread's failed create */
int  acro_set_params(acro_handle_t* h, const acro_params_t* params);

/* ---- the hot path ---------------------------------------------------------------------------------------- */
/* Bytes of device workspace acro_run_batch needs for this batch (K planes + rate work list); 0 on error.      */
int64_t acro_workspace_bytes(const acro_handle_t* h, const acro_batch_t* batch_counts);

/* Device-pointer entry: replaces NuclearReactor.get_pdi_grids + MatrixGenerator.get_final_matp + expm + Yf for
 * every point of the batch (models.py:112-130, :84-89).  Asynchronous on `stream` (a cudaStream_t; 0 = legacy). */
int acro_run_batch(acro_handle_t* h, const acro_batch_t* batch, const acro_out_t* out,
                   void* workspace, int64_t workspace_bytes, int32_t stages, void* stream);

/* Host-pointer entry: same, with H2D of the batch, device allocation of workspace/outputs, D2H of the outputs
 * that are non-NULL, and a final synchronise.  This is the call the e2e benchmark times.                       */
int acro_run_batch_host(acro_handle_t* h, const acro_batch_t* batch, const acro_out_t* out, int32_t stages);

/* Copy one K plane set of one system out of a workspace (device->device), dense [5][NE][NE] with zeros below
 * the diagonal, i.e. the reference's K[X,Xp] for (X,Xp) in ((0,0),(0,1),(1,0),(2,0),(1,1)) (cascade.py:755).  */
int acro_unpack_kernels(acro_handle_t* h, const acro_batch_t* batch, const void* workspace, int32_t system,
                        double* K_dense, void* stream);

/* MatrixGenerator.get_matp(T_lo) for given pdi grids (nucl.py:583-623), plus expm/Yf: the integral runs from the
 * top of each point's T grid down to T_lo[p] (NULL => bottom of the grid = get_final_matp, nucl.py:658-659).
 * T [S], pdi [S,17]; device pointers.                                                                           */
int acro_matp_batch(acro_handle_t* h, int32_t n_points, int32_t nt_max, const int32_t* pt_nt, const int32_t* pt_sys_off,
                    const double* sys_T, const double* pdi, const double* T_lo, const double* pt_t_at_tmax,
                    double* mpdi, double* mdcy, double* Yf, int32_t* status, void* stream);

/* Fast-parameter path (scans.py:106-107,153-157; models.py:126-130): Yf = postd . expm(factor*mpdi + mdcy) . pred . Y0
 * for n jobs; job i uses the matrix pair matrix_index[i] of mpdi/mdcy [m,9,9] (matrix_index NULL: pair i, m = n), so the
 * many rescalings of one buffered pair share its storage; factor [n] (NULL: 1), t_at_tmax [n]; device pointers.       */
int acro_transfer_batch(acro_handle_t* h, int32_t n, const double* mpdi, const double* mdcy, const int32_t* matrix_index,
                        const double* factor, const double* t_at_tmax, double* expm_out /* [n,9,9] or NULL */, double* Yf,
                        int32_t* status, void* stream);

/* Total rates Gamma_X(E,T) for n arbitrary (E,T) pairs -> out[n,3] (SpectrumGenerator._rate_x, cascade.py:721-730). */
int acro_total_rates(acro_handle_t* h, int32_t n, const double* E, const double* T, double* out, void* stream);

/* The two TABULATED rates as converged direct 2-D integrals for n arbitrary (E,T) pairs -> out[n,3] = {thermal pair creation
 * gamma gamma_th -> e+ e- (0 below its threshold E < me^2/(50T)), inverse Compton, inverse Compton}: the numbers a rates.db
 * entry holds before log10 (reference tools/create_db.py:34-40, cascade.py:336-358,469-485).  tools/make_rates_db.py uses it
 * to regenerate / extend the table on the GPU.  Synchronous on `stream`.                                            */
int acro_direct_rates(acro_handle_t* h, int32_t n, const double* E, const double* T, double* out, void* stream);

/* Measured FP64 FMA throughput of the device (DFMA micro-benchmark), in TFLOP/s; the roofline denominator of the
 * kernel builder.  Synchronous.                                                                                 */
int acro_measure_fp64_peak(acro_handle_t* h, double* tflops);

/* Device time [ms] of the last acro_run_batch stages on this handle (CUDA events on the launch stream):
 * ms[0]=rates, ms[1]=kernels, ms[2]=solve+pdi, ms[3]=matrix.  Synchronises on the recorded events.             */
int acro_last_stage_ms(acro_handle_t* h, float ms[4]);

/* Split of ms[1] for kernel_mode = ACRO_KERNELS_FAST: ms[0] = pair_closed_forms_kernel, ms[1] = ic_photon_closed_kernel
 * (e -> gamma plane from tabulated one-variable functions), ms[2] = kernels_mma_kernel (temperature-shared grid, e -> e and
 * gamma -> e integrals), ms[3] = kernels_wien_kernel (their Wien tail). */
int acro_last_kernel_split_ms(acro_handle_t* h, float ms[4]);

/* Work census of a (device-resident) batch for the roofline's algorithmic-flop figure: out[0] = (system,i,j>=i) pairs,
 * out[1..3] = pairs with a non-empty range for the e->gamma, e->e and gamma->e thermal integrals (the reference's own
 * limit tests, cascade.py:416-427, 516-534, 572-592).  Synchronous.                                              */
int acro_count_work(acro_handle_t* h, const acro_batch_t* batch, int64_t out[4], void* stream);

/* Number of kernel launches issued through this handle since creation. */
int64_t acro_launch_count(const acro_handle_t* h);

#ifdef __cplusplus
}
#endif
#endif /* ACROPOLIS_B200_H */
