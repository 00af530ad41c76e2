"""ctypes binding of ``libacropolis_b200.so`` (C ABI: include/acropolis_b200.h).

The library is built in-tree by ``__graft_entry__.build()`` / ``python -m acropolis_b200.build``.  If it is missing
or does not load, importing the compute path raises -- there is no CPU fallback."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("ACROPOLIS_B200_LIB") or os.path.join(_HERE, "libacropolis_b200.so")   # env override: kernel A/B builds
# the same sources with -DACRO_CROSSCHECK: adds the per-(E,E',T) cross-check modes (kernel_mode PLAIN_GL / PER_T) for the tests
XCHECK_LIB_PATH = os.path.join(_HERE, "libacropolis_b200_xcheck.so")

NX, NY, NR, NK = 3, 9, 17, 5
STAGE_RATES, STAGE_KERNELS, STAGE_SOLVE, STAGE_MATRIX, STAGE_ALL = 1, 2, 4, 8, 15
ST_NONFINITE, ST_T_OUTSIDE, ST_DIRECT_RATE, ST_E0_ABOVE_GEV = 1, 2, 4, 8
SC_NONE, SC_SEPARABLE, SC_DENSE = 0, 1, 2
KERNELS_FAST, KERNELS_PLAIN_GL, KERNELS_PER_T = 0, 1, 2
ABI_VERSION = 4
KERNEL_SPLIT_NAMES = ("pair_closed_forms_kernel", "ic_photon_closed_kernel", "kernels_mma_kernel", "kernels_wien_kernel")

c_double_p = C.POINTER(C.c_double)
c_int32_p = C.POINTER(C.c_int32)
c_int64_p = C.POINTER(C.c_int64)


class Tables(C.Structure):
    _fields_ = [("ratedb", C.c_void_p), ("ratedb_nT", C.c_int32), ("ratedb_nE", C.c_int32),
                ("ratedb_Elog_min", C.c_double), ("ratedb_Elog_max", C.c_double),
                ("ratedb_Tlog_min", C.c_double), ("ratedb_Tlog_max", C.c_double),
                ("cosmo_log10_T", C.c_void_p), ("cosmo_log10_abs_dTdt", C.c_void_p), ("n_cosmo", C.c_int32),
                ("eta", C.c_double), ("Y0", C.c_void_p)]


class Params(C.Structure):
    _fields_ = [("quad_order", C.c_int32), ("pdi_cell_order", C.c_int32), ("Emin", C.c_double),
                ("approx_zero", C.c_double), ("Ephb_T_max", C.c_double), ("E_EC_max", C.c_double),
                ("use_db", C.c_int32), ("kernel_mode", C.c_int32), ("wien_from", C.c_double),
                ("universal", C.c_int32), ("reserved", C.c_int32)]


class Batch(C.Structure):
    _fields_ = [("n_points", C.c_int32), ("n_systems", C.c_int32), ("n_energy_total", C.c_int64),
                ("n_sys_energy_total", C.c_int64), ("n_pairs_total", C.c_int64), ("n_pairs_points", C.c_int64),
                ("ne_max", C.c_int32),
                ("nt_max", C.c_int32), ("sc_mode", C.c_int32), ("reserved", C.c_int32),
                ("pt_ne", C.c_void_p), ("pt_nt", C.c_void_p), ("pt_e_off", C.c_void_p), ("pt_sys_off", C.c_void_p),
                ("pt_pair_off", C.c_void_p), ("pt_e0", C.c_void_p), ("pt_t_at_tmax", C.c_void_p), ("e_grid", C.c_void_p),
                ("sys_point", C.c_void_p), ("sys_T", C.c_void_p), ("sys_f_off", C.c_void_p), ("sys_k_off", C.c_void_p),
                ("s0", C.c_void_p), ("sc", C.c_void_p), ("sc_E", C.c_void_p)]


class Out(C.Structure):
    _fields_ = [("G", C.c_void_p), ("F", C.c_void_p), ("pdi", C.c_void_p), ("mpdi", C.c_void_p), ("mdcy", C.c_void_p),
                ("Yf", C.c_void_p), ("status", C.c_void_p)]


EXPORTS = ("acro_abi_version", "acro_default_params", "acro_create", "acro_destroy", "acro_last_error", "acro_set_params",
           "acro_workspace_bytes", "acro_run_batch", "acro_run_batch_host", "acro_unpack_kernels", "acro_matp_batch",
           "acro_transfer_batch", "acro_total_rates", "acro_measure_fp64_peak", "acro_last_stage_ms", "acro_last_kernel_split_ms", "acro_launch_count",
           "acro_count_work", "acro_direct_rates")

_libs = {}


def load(variant="product"):
    """Load the shared library (once per variant) and declare the prototypes; raises if it has not been built.
    ``variant`` = "product" (what ships) or "xcheck" (test build with the cross-check kernel modes)."""
    if variant in _libs:
        return _libs[variant]
    path = LIB_PATH if variant == "product" else XCHECK_LIB_PATH
    if not os.path.exists(path):
        raise ImportError("acropolis_b200: %s is missing -- build it with `python -c 'import __graft_entry__ as g; "
                          "g.build()'` (nvcc, sm_100a). There is no CPU fallback." % path)
    lib = C.CDLL(path)
    vp, i32, i64 = C.c_void_p, C.c_int32, C.c_int64
    lib.acro_abi_version.restype = C.c_int
    lib.acro_default_params.argtypes = [C.POINTER(Params)]
    lib.acro_default_params.restype = None
    lib.acro_create.argtypes = [C.POINTER(vp), C.c_int, C.POINTER(Tables), C.POINTER(Params)]
    lib.acro_destroy.argtypes = [vp]
    lib.acro_last_error.argtypes = [vp]
    lib.acro_last_error.restype = C.c_char_p
    lib.acro_set_params.argtypes = [vp, C.POINTER(Params)]
    lib.acro_workspace_bytes.argtypes = [vp, C.POINTER(Batch)]
    lib.acro_workspace_bytes.restype = i64
    lib.acro_run_batch.argtypes = [vp, C.POINTER(Batch), C.POINTER(Out), vp, i64, i32, vp]
    lib.acro_run_batch_host.argtypes = [vp, C.POINTER(Batch), C.POINTER(Out), i32]
    lib.acro_unpack_kernels.argtypes = [vp, C.POINTER(Batch), vp, i32, vp, vp]
    lib.acro_matp_batch.argtypes = [vp, i32, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]
    lib.acro_transfer_batch.argtypes = [vp, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp]
    lib.acro_total_rates.argtypes = [vp, i32, vp, vp, vp, vp]
    lib.acro_direct_rates.argtypes = [vp, i32, vp, vp, vp, vp]
    lib.acro_measure_fp64_peak.argtypes = [vp, C.POINTER(C.c_double)]
    lib.acro_last_stage_ms.argtypes = [vp, C.POINTER(C.c_float * 4)]
    lib.acro_last_kernel_split_ms.argtypes = [vp, C.POINTER(C.c_float * 4)]
    lib.acro_count_work.argtypes = [vp, C.POINTER(Batch), C.POINTER(C.c_int64 * 4), vp]
    lib.acro_launch_count.argtypes = [vp]
    lib.acro_launch_count.restype = i64
    for name in EXPORTS:
        getattr(lib, name)
    if lib.acro_abi_version() != ABI_VERSION:
        raise ImportError("acropolis_b200: ABI version mismatch in %s" % path)
    _libs[variant] = lib
    return lib
