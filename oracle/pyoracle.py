"""ctypes wrapper of oracle/liboracle.so -- TEST INFRASTRUCTURE ONLY (see the header of acro_oracle.c).

Imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs; never by
acropolis_b200/."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "liboracle.so")


class _OracleStruct(C.Structure):
    _fields_ = [("ratedb", C.c_void_p), ("nT", C.c_int), ("nE", C.c_int),
                ("Elog_min", C.c_double), ("Elog_max", C.c_double), ("Tlog_min", C.c_double), ("Tlog_max", C.c_double),
                ("cosmo_lT", C.c_void_p), ("cosmo_ld", C.c_void_p), ("n_cosmo", C.c_int),
                ("eta", C.c_double), ("Yp", C.c_double), ("YHe4", C.c_double), ("Y0", C.c_double * 27),
                ("eps", C.c_double), ("Emin", C.c_double), ("approx_zero", C.c_double), ("Ephb_T_max", C.c_double),
                ("E_EC_max", C.c_double), ("use_db", C.c_int), ("neval", C.c_long)]


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def load_lib():
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(os.path.join(HERE, "acro_oracle.c")):
        subprocess.check_call(["make", "-s", "-C", HERE])
    lib = C.CDLL(LIB)
    assert lib.oracle_sizeof() == C.sizeof(_OracleStruct), "oracle_t layout mismatch"
    d, i, vp = C.c_double, C.c_int, C.c_void_p
    P = C.POINTER(_OracleStruct)
    for name, args, res in (
            ("oracle_F", [d, d, d], d), ("oracle_G", [d, d, d], d), ("oracle_dsdE_Z2", [d, d], d),
            ("oracle_total_rate", [P, i, d, d], d), ("oracle_rate_piece", [P, i, d, d], d),
            ("oracle_kernel_piece", [P, i, d, d, d], d), ("oracle_abs_dTdt", [P, d], d),
            ("oracle_cross_section", [i, d], d), ("oracle_quad_test", [P, d, d, d, d, d, i], d),
            ("oracle_rates_kernels", [P, i, vp, d, vp, vp], i),
            ("oracle_solve", [P, i, vp, vp, vp, vp, vp, vp], None),
            ("oracle_pdi_rates", [P, i, vp, vp, d, d, d, d, vp], i),
            ("oracle_matp", [P, i, vp, vp, d, vp, vp], None), ("oracle_expm", [vp, vp], None),
            ("oracle_final_abundances", [P, vp, vp, d, vp], None),
            ("oracle_run_point", [P, i, vp, i, vp, d, vp, vp, d, vp, vp, vp, vp, vp, vp], i),
            ("oracle_run_points", [P, i, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, i, vp], i)):
        f = getattr(lib, name)
        f.argtypes, f.restype = args, res
    return lib


class Oracle(object):
    """CPU oracle bound to one set of input tables.  `ii` is an acropolis_b200.input.InputInterface (data only),
    `ratedb` the float64[Tnum*Enum,2] table; both default to the packaged data files."""

    def __init__(self, ii=None, ratedb=None, eps=1e-3, use_db=True):
        self.lib = load_lib()
        if ii is None or ratedb is None:
            import sys
            sys.path.insert(0, os.path.dirname(HERE))
            from acropolis_b200.input import shared_input_interface
            from acropolis_b200.db import import_data_from_db
            ii = shared_input_interface() if ii is None else ii
            ratedb = import_data_from_db() if ratedb is None else ratedb
        self._ratedb = np.ascontiguousarray(ratedb, dtype=np.float64)
        lT, ld = ii.cosmo_log_tables()
        self._lT, self._ld = np.ascontiguousarray(lT), np.ascontiguousarray(ld)
        Y0 = np.ascontiguousarray(ii.bbn_abundances(), dtype=np.float64)
        s = _OracleStruct()
        s.ratedb, s.nT, s.nE = self._ratedb.ctypes.data, 750, 450
        s.Elog_min, s.Elog_max, s.Tlog_min, s.Tlog_max = 0.0, 3.0, -6.0, -1.0
        s.cosmo_lT, s.cosmo_ld, s.n_cosmo = self._lT.ctypes.data, self._ld.ctypes.data, len(self._lT)
        s.eta, s.Yp, s.YHe4 = float(ii.parameter("eta")), Y0[1, 0], Y0[5, 0]
        for k, v in enumerate(Y0[:, :3].reshape(-1)):
            s.Y0[k] = v
        s.eps, s.Emin, s.approx_zero, s.Ephb_T_max, s.E_EC_max = eps, 1.5, 1e-200, 200.0, 10.0
        s.use_db, s.neval = int(use_db), 0
        self.s = s

    def set_eps(self, eps):
        self.s.eps = eps

    def total_rate(self, X, E, T):
        return self.lib.oracle_total_rate(C.byref(self.s), X, E, T)

    def rates_kernels(self, E_grid, T, want_K=True):
        E_grid = np.ascontiguousarray(E_grid, dtype=np.float64)
        NE = len(E_grid)
        G = np.zeros((3, NE))
        K = np.zeros((5, NE, NE)) if want_K else None
        self.lib.oracle_rates_kernels(C.byref(self.s), NE, _ptr(E_grid), T, _ptr(G), _ptr(K) if want_K else None)
        return G, K

    def solve(self, E_grid, G, K, S0, SC):
        E_grid = np.ascontiguousarray(E_grid, dtype=np.float64)
        NE = len(E_grid)
        F = np.zeros((3, NE))
        G, K = np.ascontiguousarray(G), np.ascontiguousarray(K)
        S0, SC = np.ascontiguousarray(S0, dtype=np.float64), np.ascontiguousarray(SC, dtype=np.float64)
        self.lib.oracle_solve(C.byref(self.s), NE, _ptr(E_grid), _ptr(G), _ptr(K), _ptr(S0), _ptr(SC), _ptr(F))
        return F

    def pdi_rates(self, E_grid, Fph, E0, T, S0_photon, rate_photon_E0):
        E_grid, Fph = np.ascontiguousarray(E_grid, dtype=np.float64), np.ascontiguousarray(Fph, dtype=np.float64)
        r = np.zeros(17)
        w = self.lib.oracle_pdi_rates(C.byref(self.s), len(E_grid), _ptr(E_grid), _ptr(Fph), E0, T, S0_photon,
                                      rate_photon_E0, _ptr(r))
        return r, w

    def matp(self, Tr, pdi, T_lo=None):
        Tr, pdi = np.ascontiguousarray(Tr, dtype=np.float64), np.ascontiguousarray(pdi, dtype=np.float64)
        mp, md = np.zeros((9, 9)), np.zeros((9, 9))
        self.lib.oracle_matp(C.byref(self.s), len(Tr), _ptr(Tr), _ptr(pdi), Tr[0] if T_lo is None else T_lo, _ptr(mp), _ptr(md))
        return mp, md

    def expm(self, A):
        A = np.ascontiguousarray(A, dtype=np.float64)
        R = np.zeros((9, 9))
        self.lib.oracle_expm(_ptr(A), _ptr(R))
        return R

    def final_abundances(self, mpdi, mdcy, t_at_tmax):
        mp, md = np.ascontiguousarray(mpdi, dtype=np.float64), np.ascontiguousarray(mdcy, dtype=np.float64)
        Yf = np.zeros((9, 3))
        self.lib.oracle_final_abundances(C.byref(self.s), _ptr(mp), _ptr(md), t_at_tmax, _ptr(Yf))
        return Yf

    @staticmethod
    def _dense_sc(spec):
        NT, NE = len(spec.T), len(spec.E_grid)
        if spec.sc_mode == 0 or spec.sc is None:
            return None
        if spec.sc_mode == 1:
            return np.ascontiguousarray(np.asarray(spec.sc)[:, :, None] * np.asarray(spec.sc_E).T[None, :, :])
        return np.ascontiguousarray(np.asarray(spec.sc, dtype=np.float64).reshape(NT, 3, NE))

    def run_point(self, spec):
        """Full chain for one PointSpec -> dict(G, F, pdi, mpdi, mdcy, Yf, warnings)."""
        NT, NE = len(spec.T), len(spec.E_grid)
        G, F = np.zeros((NT, 3, NE)), np.zeros((NT, 3, NE))
        pdi, mp, md, Yf = np.zeros((NT, 17)), np.zeros((9, 9)), np.zeros((9, 9)), np.zeros((9, 3))
        SC = self._dense_sc(spec)
        E, T, S0 = (np.ascontiguousarray(a, dtype=np.float64) for a in (spec.E_grid, spec.T, spec.S0))
        w = self.lib.oracle_run_point(C.byref(self.s), NE, _ptr(E), NT, _ptr(T), spec.E0, _ptr(S0),
                                      _ptr(SC) if SC is not None else None, spec.t_at_tmax, _ptr(G), _ptr(F), _ptr(pdi),
                                      _ptr(mp), _ptr(md), _ptr(Yf))
        return dict(G=G, F=F, pdi=pdi, mpdi=mp, mdcy=md, Yf=Yf, warnings=w)

    def run_points(self, arrays, counts, n_threads=0):
        """Many points (flat arrays from Engine.assemble) on all host threads -> (Yf[P,9,3], integrand evaluations)."""
        P = counts["n_points"]
        Yf = np.zeros((P, 9, 3))
        nev = C.c_long(0)
        sep = counts["sc_mode"] == 1
        assert counts["sc_mode"] in (0, 1), "oracle.run_points: dense SC not supported, use run_point"
        a = arrays
        self.lib.oracle_run_points(C.byref(self.s), P, _ptr(a["pt_ne"]), _ptr(a["pt_nt"]), _ptr(a["pt_e_off"]),
                                   _ptr(a["pt_sys_off"]), _ptr(a["pt_e0"]), _ptr(a["pt_t_at_tmax"]), _ptr(a["e_grid"]),
                                   _ptr(a["sys_T"]), _ptr(a["s0"]), _ptr(a["sc"]) if sep else None,
                                   _ptr(a["sc_E"]) if sep else None, _ptr(Yf), n_threads, C.byref(nev))
        return Yf, nev.value
