#!/usr/bin/env python3
"""Generate the golden fixtures under tests/golden/ by running the UNMODIFIED reference.

This script is the only place that imports ``acropolis`` from ``/root/reference``.  It exists
because the reference (hep-mh/acropolis v1.3.1) ships no test-suite: the only
stored known answers are ``tools/data/*.dat``.  Everything else that pins the oracle (``oracle/``)
and the CUDA path is produced here, once, in the build container, and committed as ``.npz`` /
``.json`` so that the GPU box (which has no ``/root/reference``) can check against it.

    NUMBA_CACHE_DIR=/tmp/numba python tests/golden/make_golden.py [case ...] [-j 8]

Each case is independent and runs in its own process (the reference keeps per-function caches and
module-level resolution parameters, so cases that patch ``eps``/``NE_pd``/``NT_pd`` must not share
an interpreter).
"""
import argparse
import json
import os
import sys
import time
from multiprocessing import get_context

REF = os.environ.get("ACROPOLIS_REFERENCE", "/root/reference")
OUT = os.path.dirname(os.path.abspath(__file__))


def _import_reference(eps_cascade=None, eps_nucl=None, NE_pd=None, NT_pd=None):
    sys.path.insert(0, REF)
    import numpy as np  # noqa
    import acropolis.flags as flags
    flags.verbose = False
    import acropolis.cascade as cascade
    import acropolis.nucl as nucl
    # Resolution parameters are imported *by value* into the consumer modules
    # (cascade.py:22-23, nucl.py:21-22), so they are patched there.
    if eps_cascade is not None:
        cascade.eps = eps_cascade
    if eps_nucl is not None:
        nucl.eps = eps_nucl
    if NE_pd is not None:
        cascade.NE_pd = NE_pd
    if NT_pd is not None:
        nucl.NT_pd = NT_pd
    return cascade, nucl


DISTINCT = ((0, 0), (0, 1), (1, 0), (2, 0), (1, 1))  # gg, eg, ge-, ge+, ee


def _gk_at(sgen, E0, T, cascade):
    """G[3,NE] and the five distinct K[NE,NE] exactly as cascade.py:736-755 evaluates them."""
    import numpy as np
    from math import log, log10
    NE = max(int(log10(E0 / cascade.Emin) * cascade.NE_pd), cascade.NE_min)
    E_grid = np.logspace(log(cascade.Emin), log(E0), NE, base=np.e)
    G = np.array([[sgen._rate_x(X, E, T) for E in E_grid] for X in range(3)])
    K = np.zeros((5, NE, NE))
    for k, (X, Xp) in enumerate(DISTINCT):
        for i, E in enumerate(E_grid):
            for j, Ep in enumerate(E_grid):
                if Ep >= E:
                    K[k, i, j] = sgen._kernel_x_xp(X, Xp, E, Ep, T)
    return E_grid, G, K


def _full_point(model, cascade, nucl, k_at=()):
    """Run one model through the reference chain (models.py:112-130) recording every stage."""
    import numpy as np
    from scipy.linalg import expm
    rec = {"T": [], "sol": []}
    nr = nucl.NuclearReactor(model._sS0, model._sSc, model._sTrg, model._sE0, model._sII)
    orig = nr._sGen.get_spectrum

    def spy(E0, S0f, SCf, T, allX=False):
        sol = orig(E0, S0f, SCf, T, allX=True)
        rec["T"].append(T)
        rec["sol"].append(sol.copy())
        return sol[0:2, :] if not allX else sol

    nr._sGen.get_spectrum = spy
    t0 = time.time()
    temp, pdi_grids = nr.get_pdi_grids()
    mg = nucl.MatrixGenerator(temp, pdi_grids, model._sII)
    mpdi, mdcy = mg.get_final_matp()
    t1 = time.time()
    sol = np.array(rec["sol"])                     # [NT, 4, NE]
    E_grid = sol[0, 0, :]
    S0 = np.array([[f(T) for f in model._sS0] for T in temp])                        # [NT,3]
    SC = np.array([[[f(E, T) for E in E_grid] for f in model._sSc] for T in temp])   # [NT,3,NE]
    pdi = np.array([pdi_grids[r] for r in range(1, 18)]).T                            # [NT,17]
    rate_E0 = np.array([nr._sGen.rate_photon(model._sE0, T) for T in temp])
    model.set_matp_buffer((mpdi, mdcy))
    Yf = model.run_disintegration()
    out = dict(E0=model._sE0, Trg=np.array(model._sTrg), T=np.array(temp), E_grid=E_grid,
               S0=S0, SC=SC, F=sol[:, 1:, :], pdi=pdi, rate_photon_E0=rate_E0,
               mpdi=mpdi, mdcy=mdcy, expm=expm(mpdi + mdcy),
               pred=model._pred_matrix(), postd=model._postd_matrix(),
               t_at_Tmax=model._sII.time(max(model._sTrg)), Yf=Yf, wall_s=t1 - t0)
    G_all = np.array([[[nr._sGen._rate_x(X, E, T) for E in E_grid] for X in range(3)] for T in temp])
    out["G"] = G_all
    for it in k_at:
        _, G, K = _gk_at(nr._sGen, model._sE0, temp[it], cascade)
        out["K_it%d" % it] = K
    out["k_at"] = np.array(list(k_at), dtype=np.int64)
    return out


# ------------------------------------------------------------------------------------------ cases

def case_kat_scalars():
    """Scalar known answers for every closed form on the path (SURVEY.md 8c list, extended)."""
    import numpy as np
    cascade, nucl = _import_reference()
    from acropolis.models import DecayModel
    from acropolis.input import InputInterface, locate_sm_file
    from acropolis import db
    ii = InputInterface(locate_sm_file())
    sg = cascade.SpectrumGenerator(ii)
    ph, el, po = sg._sRW[0], sg._sRW[1], sg._sRW[2]
    out = {}
    out["eta"] = float(ii.parameter("eta"))
    out["Y0"] = ii.bbn_abundances().tolist()
    trips = [(2., 50., 1e-3), (1.6, 3., 2e-4), (10., 400., 5e-2), (0.02, 5., 1e-2)]
    out["JIT_F"] = [[*t, cascade._JIT_F(*t)] for t in trips]
    trips = [(3., 300., 1e-3), (20., 60., 1e-2), (2., 5., 0.08), (100., 250., 2e-3)]
    out["JIT_G"] = [[*t, cascade._JIT_G(*t)] for t in trips]
    pairs = [(2., 5.), (1.6, 300.), (30., 61.), (1.5, 2.6), (200., 1000.)]
    out["JIT_dsdE_Z2"] = [[*p, cascade._JIT_dsdE_Z2(*p)] for p in pairs]
    ET = [(E, T) for E in (1.5, 2.0, 2.05, 3.3, 10., 77., 300., 999.) for T in (2e-6, 1e-4, 1e-3, 1.3e-2, 9e-2)]
    out["ne"] = [[T, ph._ne(T)] for T in (1e-6, 1e-3, 0.1)]
    out["nNZ2"] = [[T, ph._nNZ2(T)] for T in (1e-6, 1e-3, 0.1)]
    out["rate_photon_photon"] = [[E, T, ph._rate_photon_photon(E, T)] for E, T in ET]
    out["rate_compton"] = [[E, T, ph._rate_compton(E, T)] for E, T in ET]
    out["rate_bethe_heitler"] = [[E, T, ph._rate_bethe_heitler(E, T)] for E, T in ET]
    out["rate_pair_creation_ae_db"] = [[E, T, ph._rate_pair_creation_ae_db(E, T)] for E, T in ET]
    out["rate_inverse_compton_db"] = [[E, T, el._rate_inverse_compton_db(E, T)] for E, T in ET]
    out["total_rate_photon"] = [[E, T, ph.total_rate(E, T)] for E, T in ET]
    EEpT = [(3., 300., 1e-3), (1.5, 1.5, 1e-3), (2., 2.3, 1e-4), (2., 2.6, 1e-4), (5., 40., 1.2e-2),
            (1.7, 500., 2e-6), (100., 101., 5e-3), (9., 10., 8e-2), (1.5, 2.5, 3e-5), (33., 47., 4e-4)]
    for name, fn in (("kernel_photon_photon", ph._kernel_photon_photon), ("kernel_compton_ph", ph._kernel_compton),
                     ("kernel_inverse_compton_ph", ph._kernel_inverse_compton),
                     ("kernel_inverse_compton_el", el._kernel_inverse_compton),
                     ("kernel_compton_el", el._kernel_compton), ("kernel_bethe_heitler", el._kernel_bethe_heitler),
                     ("kernel_pair_creation_ae", el._kernel_pair_creation_ae)):
        out[name] = [[*t, fn(*t)] for t in EEpT]
    out["time_of_T"] = [[T, ii.time(T)] for T in (1e-6, 1e-3, 0.5, 10.)]
    out["T_of_time"] = [[t, ii.temperature(t)] for t in (1e3, 1e5, 1e7, 1e10)]
    out["dTdt"] = [[T, ii.dTdt(T)] for T in (1e-6, 1e-3, 0.5)]
    out["scale_factor"] = [[T, ii.scale_factor(T)] for T in (1e-6, 1e-3, 0.5, 10.)]
    out["hubble_rate"] = [[T, ii.hubble_rate(T)] for T in (1e-6, 1e-3, 0.5)]
    out["temperature_range"] = list(ii.temperature_range())
    m = DecayModel(10., 1e5, 10., 1e-10, 0., 1.)
    nr = nucl.NuclearReactor(m._sS0, m._sSc, m._sTrg, m._sE0, m._sII)
    Es = (1.59, 2.3, 2.5, 3.0, 4.0, 5.0, 5.6, 5.7, 6.3, 7.3, 7.5, 8.0, 9.0, 9.4, 11.0, 15.0, 16.0, 19.0,
          20.0, 21.0, 24.0, 25.0, 27.0, 30.0, 43.0, 60.0, 100.0, 400.0)
    out["sigma"] = [[rid, E, nr.get_cross_section(rid, E)] for rid in range(1, 18) for E in Es]
    with open(os.path.join(OUT, "kat_scalars.json"), "w") as f:
        json.dump(out, f, indent=0)
    # raw db spot values (to pin the converted table + index convention)
    rdb = db.import_data_from_db()
    idx = [0, 1, 449, 450, 12345, 200000, 337499]
    np.savez_compressed(os.path.join(OUT, "ratedb_spots.npz"), idx=np.array(idx), vals=np.array([rdb[i] for i in idx]),
                        shape=np.array(rdb.shape))


def _decay_case(name, args, k_at=(), **patch):
    import numpy as np
    cascade, nucl = _import_reference(**patch)
    from acropolis.models import DecayModel
    m = DecayModel(*args)
    out = _full_point(m, cascade, nucl, k_at=k_at)
    out["ctor"] = np.array(args, dtype=float)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)


def case_cfg1():
    _decay_case("cfg1_decay", (10., 1e5, 10., 1e-10, 0., 1.), k_at=(0, 20, 39))


def case_cfg1_tight():
    _decay_case("cfg1_decay_tight", (10., 1e5, 10., 1e-10, 0., 1.), k_at=(0, 20, 39), eps_cascade=1e-9, eps_nucl=1e-9)


def case_decay_ee():
    # e+e- final state: monochromatic e-/e+ injection plus the FSR continuous photon source
    _decay_case("decay_ee", (30., 1e6, 10., 1e-9, 1., 0.), k_at=(5,))


def case_decay_ee_tight():
    _decay_case("decay_ee_tight", (30., 1e6, 10., 1e-9, 1., 0.), k_at=(), eps_cascade=1e-9, eps_nucl=1e-9)


def case_decay_nemin():
    # E0 = 1.6 MeV -> NE_min = 10 grid; only Be7 (1.5866) is above threshold
    _decay_case("decay_nemin", (3.2, 1e6, 10., 1e-6, 0., 1.), k_at=(10,))


def case_decay_mixed():
    _decay_case("decay_mixed", (80., 3e4, 10., 3e-8, 0.4, 0.6), k_at=())


def case_decay_lowT():
    # tau = 1e10 s: the T grid reaches below 1e-6 MeV => IC rate outside rates.db (cascade.py:488-493)
    _decay_case("decay_lowT", (12., 1e10, 10., 1e-7, 0., 1.), k_at=(0,))


def case_decay_highT():
    # tau = 1e3 s: the T grid reaches above 0.1 MeV => both db rates outside the table
    _decay_case("decay_highT", (6., 1e3, 10., 1e-3, 0., 1.), k_at=(39,))


def case_decay_big():
    _decay_case("decay_big", (400., 1e8, 10., 1e-11, 0., 1.), k_at=())


def case_cfg2():
    import numpy as np
    cascade, nucl = _import_reference()
    from acropolis.models import AnnihilationModel
    args = (10., 1e-25, 0., 0., 1., 0.)
    m = AnnihilationModel(*args)
    out = _full_point(m, cascade, nucl, k_at=(40,))
    out["ctor"] = np.array(args, dtype=float)
    np.savez_compressed(os.path.join(OUT, "cfg2_annih.npz"), **out)


def case_annih_pwave():
    import numpy as np
    cascade, nucl = _import_reference()
    from acropolis.models import AnnihilationModel
    args = (25., 0., 1e-17, 1e-3, 0.3, 0.7)
    m = AnnihilationModel(*args)
    out = _full_point(m, cascade, nucl)
    out["ctor"] = np.array(args, dtype=float)
    np.savez_compressed(os.path.join(OUT, "annih_pwave.npz"), **out)


def case_resonance():
    import numpy as np
    cascade, nucl = _import_reference()
    from acropolis.ext.benchmarks import BenchmarkModel3
    kw = dict(mchi=10., delta=1e-2, gammad=1e-3, gammav=1e-12)
    m = BenchmarkModel3(**kw)
    out = _full_point(m, cascade, nucl)
    out["ctor"] = np.array([kw["mchi"], kw["delta"], kw["gammad"], kw["gammav"]])
    out["tempkd"] = m._sTkd
    Ts = np.logspace(-6, -1.8, 12)
    out["sigma_v_T"] = Ts
    out["sigma_v"] = np.array([m._sigma_v(T) for T in Ts])
    np.savez_compressed(os.path.join(OUT, "resonance_b3.npz"), **out)


def case_resonance_b1():
    import numpy as np
    cascade, nucl = _import_reference()
    from acropolis.ext.benchmarks import BenchmarkModel1, BenchmarkModel2
    kw = dict(mchi=10., delta=1e-2, gammad=1e-3, gammav=1e-12)
    out = {}
    for tag, M in (("b1", BenchmarkModel1), ("b2", BenchmarkModel2)):
        m = M(**kw)
        Ts = np.logspace(-6, -1.8, 12)
        out[tag + "_tempkd"] = m._sTkd
        out[tag + "_sigma_v_T"] = Ts
        out[tag + "_sigma_v"] = np.array([m._sigma_v(T) for T in Ts])
        out[tag + "_S0"] = np.array([[f(T) for f in m._sS0] for T in Ts])
    np.savez_compressed(os.path.join(OUT, "resonance_b12_hooks.npz"), **out)


def _rates_direct(tag, eps):
    """Out-of-table 2-D rate integrals (cascade.py:336-358, 469-485), flags.usedb=False."""
    import numpy as np
    cascade, nucl = _import_reference(eps_cascade=eps)
    import acropolis.flags as flags
    flags.usedb = False
    from acropolis.input import InputInterface, locate_sm_file
    ii = InputInterface(locate_sm_file())
    sg = cascade.SpectrumGenerator(ii)
    ph, el = sg._sRW[0], sg._sRW[1]
    Es = np.array([1.5, 2.2, 5., 17., 60., 250., 999., 1500.])
    Ts = np.array([3e-7, 7e-7, 1e-5, 1e-3, 2e-2, 0.11, 0.14])
    pc = np.array([[ph._rate_pair_creation_ae_db(E, T) for T in Ts] for E in Es])
    ic = np.array([[el._rate_inverse_compton_db(E, T) for T in Ts] for E in Es])
    np.savez_compressed(os.path.join(OUT, "rates_direct_%s.npz" % tag), E=Es, T=Ts, pair_creation=pc,
                        inverse_compton=ic, eps=eps)


def case_rates_direct():
    _rates_direct("default", 1e-3)


def case_rates_direct_tight():
    _rates_direct("tight", 1e-9)


def case_scan5x5():
    """cfg3 sub-grid: same log ranges as the 200x200 scan (examples/scan_decay_*), 5x5 points."""
    import numpy as np
    _import_reference()
    from acropolis.models import DecayModel
    from acropolis.scans import BufferedScanner, ScanParameter
    t0 = time.time()
    res = BufferedScanner(DecayModel, mphi=ScanParameter(0, 3, 5), tau=ScanParameter(3, 10, 5), temp0=10.,
                          n0a=1e-10, bree=0., braa=1.).perform_scan(cores=-1)
    np.savez_compressed(os.path.join(OUT, "scan_decay_5x5.npz"), rows=res, wall_s=time.time() - t0,
                        cores=os.cpu_count())


def case_scan_fast():
    """Fast-parameter scan (scans.py:127-176): mphi x n0a(fast) at tau=1e7, 3x8 points."""
    import numpy as np
    _import_reference()
    from acropolis.models import DecayModel, AnnihilationModel
    from acropolis.scans import BufferedScanner, ScanParameter
    res = BufferedScanner(DecayModel, mphi=ScanParameter(0.1, 1.6, 4), tau=1e7, temp0=10.,
                          n0a=ScanParameter(-14, -3, 8, fast=True), bree=0., braa=1.).perform_scan(cores=4)
    res2 = BufferedScanner(AnnihilationModel, mchi=ScanParameter(0.5, 1.2, 2), a=0.,
                           b=ScanParameter(-23, -10, 6, fast=True), tempkd=1., bree=1., braa=0.).perform_scan(cores=2)
    np.savez_compressed(os.path.join(OUT, "scan_fast.npz"), decay_rows=res, annih_rows=res2)


def case_hires_rows():
    """Copy the reference's own stored known answers (tools/data/*.dat) as a fixture."""
    import numpy as np
    d = {}
    for k in ("NE_pd_decay", "NT_pd_decay", "NE_pd_annih", "NT_pd_annih"):
        d[k] = np.loadtxt(os.path.join(REF, "tools", "data", k + ".dat"))
    np.savez_compressed(os.path.join(OUT, "reference_convergence_rows.npz"), **d)


def case_hires_point():
    # one stored row re-run here (tools/data/NE_pd_decay.dat: NE_pd=60 at NT_pd=50) incl. all stages
    _decay_case("decay_hires_ne60_nt50", (50., 1e5, 10., 1e-8, 0., 1.), k_at=(), NE_pd=60, NT_pd=50)


# ------------------------------------------------------------------------------------------ round-2 cases

def _annih_case(name, args, **patch):
    import numpy as np
    cascade, nucl = _import_reference(**patch)
    from acropolis.models import AnnihilationModel
    m = AnnihilationModel(*args)
    out = _full_point(m, cascade, nucl)
    out["ctor"] = np.array(args, dtype=float)
    out.pop("G", None)          # rates.db interpolation does not depend on eps: already pinned by the default fixtures
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)


def case_cfg2_tight():
    # BASELINE config 2 with the reference's quadratures run to eps = 1e-9 (cascade AND nucl)
    _annih_case("cfg2_annih_tight", (10., 1e-25, 0., 0., 1., 0.), eps_cascade=1e-9, eps_nucl=1e-9)


def case_annih_pwave_tight():
    _annih_case("annih_pwave_tight", (25., 0., 1e-17, 1e-3, 0.3, 0.7), eps_cascade=1e-9, eps_nucl=1e-9)


def case_resonance_tight():
    import numpy as np
    cascade, nucl = _import_reference(eps_cascade=1e-9, eps_nucl=1e-9)
    from acropolis.ext.benchmarks import BenchmarkModel3
    kw = dict(mchi=10., delta=1e-2, gammad=1e-3, gammav=1e-12)
    m = BenchmarkModel3(**kw)
    out = _full_point(m, cascade, nucl)
    out["ctor"] = np.array([kw["mchi"], kw["delta"], kw["gammad"], kw["gammav"]])
    out["tempkd"] = m._sTkd
    out.pop("G", None)
    np.savez_compressed(os.path.join(OUT, "resonance_b3_tight.npz"), **out)


def _resonance_full(tag, cls_name):
    """BenchmarkModel1 / 2 (ext/benchmarks.py:50-51) through the whole chain at the reference defaults."""
    import numpy as np
    cascade, nucl = _import_reference()
    import acropolis.ext.benchmarks as bm
    kw = dict(mchi=10., delta=1e-2, gammad=1e-3, gammav=1e-12)
    m = getattr(bm, cls_name)(**kw)
    out = _full_point(m, cascade, nucl)
    out["ctor"] = np.array([kw["mchi"], kw["delta"], kw["gammad"], kw["gammav"]])
    out["tempkd"] = m._sTkd
    for k in ("G", "SC", "expm"):
        out.pop(k, None)
    np.savez_compressed(os.path.join(OUT, "resonance_%s_full.npz" % tag), **out)


def case_resonance_b1_full():
    _resonance_full("b1", "BenchmarkModel1")


def case_resonance_b2_full():
    _resonance_full("b2", "BenchmarkModel2")


def case_all_matp():
    """MatrixGenerator.get_all_matp (nucl.py:626-655) for BASELINE config 1: the rate matrices integrated from Tmax down to
    every grid temperature (the abundance evolution Y(T)), default eps and eps = 1e-9 for the T-integrals."""
    import numpy as np
    cascade, nucl = _import_reference()
    from acropolis.models import DecayModel
    m = DecayModel(10., 1e5, 10., 1e-10, 0., 1.)
    nr = nucl.NuclearReactor(m._sS0, m._sSc, m._sTrg, m._sE0, m._sII)
    temp, pdi_grids = nr.get_pdi_grids()
    mg = nucl.MatrixGenerator(temp, pdi_grids, m._sII)
    _, (all_mpdi, all_mdcy) = mg.get_all_matp()
    nucl.eps = 1e-9
    _, (all_mpdi_t, all_mdcy_t) = nucl.MatrixGenerator(temp, pdi_grids, m._sII).get_all_matp()
    pdi = np.array([pdi_grids[r] for r in range(1, 18)]).T
    np.savez_compressed(os.path.join(OUT, "cfg1_all_matp.npz"), T=np.array(temp), pdi=pdi, all_mpdi=all_mpdi,
                        all_mdcy=all_mdcy, all_mpdi_tight=all_mpdi_t, all_mdcy_tight=all_mdcy_t,
                        t_at_Tmax=m._sII.time(max(m._sTrg)))


def case_universal():
    """flags.universal = True (cascade.py:774-812, nucl.py:412-422): the universal photon spectrum instead of the cascade
    solution, for BASELINE config 1 and one e+e- point; spectra, pdi rates, matrices and abundances."""
    import numpy as np
    cascade, nucl = _import_reference()
    import acropolis.flags as flags
    flags.set_universal(True)
    from acropolis.models import DecayModel
    out = {}
    for tag, args in (("cfg1", (10., 1e5, 10., 1e-10, 0., 1.)), ("ee", (30., 1e6, 10., 1e-9, 1., 0.))):
        m = DecayModel(*args)
        nr = nucl.NuclearReactor(m._sS0, m._sSc, m._sTrg, m._sE0, m._sII)
        temp, pdi_grids = nr.get_pdi_grids()
        mpdi, mdcy = nucl.MatrixGenerator(temp, pdi_grids, m._sII).get_final_matp()
        m.set_matp_buffer((mpdi, mdcy))
        Yf = m.run_disintegration()
        spec = np.array([nr._sGen.get_universal_spectrum(m._sE0, m._sS0, m._sSc, T, offset=5e-2) for T in temp])
        out[tag + "_ctor"] = np.array(args)
        out[tag + "_T"] = np.array(temp)
        out[tag + "_E_grid"] = spec[0, 0]
        out[tag + "_F"] = spec[:, 1]
        out[tag + "_pdi"] = np.array([pdi_grids[r] for r in range(1, 18)]).T
        out[tag + "_mpdi"], out[tag + "_mdcy"], out[tag + "_Yf"] = mpdi, mdcy, Yf
    np.savez_compressed(os.path.join(OUT, "universal.npz"), **out)


def case_pdi_exact():
    """The 17 pdi-rate integrals of nucl.py:439-464 evaluated to convergence ON THE REFERENCE'S OWN photon spectra: the
    integrand Fph_s of nucl.py:428-431 (LogInterp of the reference spectrum x E x get_cross_section) handed to
    scipy.integrate.quad with break points at the grid knots, epsrel 1e-12, limit 2000.  Settles which of two 'converged'
    numbers is right where the reference at eps = 1e-9 (QAGS, limit 50, no break points, IntegrationWarning) and the cell-wise
    rule of the CUDA path / the C oracle differ (decay_ee, 1.8e-4)."""
    import numpy as np
    import warnings
    from math import log, exp
    from scipy.integrate import quad
    cascade, nucl = _import_reference()
    from acropolis.models import DecayModel
    from acropolis.utils import LogInterp
    out = {}
    for tag in ("decay_ee_tight", "cfg1_decay_tight"):
        fx = np.load(os.path.join(OUT, tag + ".npz"))
        m = DecayModel(*fx["ctor"])
        nr = nucl.NuclearReactor(m._sS0, m._sSc, m._sTrg, m._sE0, m._sII)
        E_grid, T, F, S0 = fx["E_grid"], fx["T"], fx["F"], fx["S0"]
        pdi = np.zeros((len(T), 17))
        for it, Ti in enumerate(T):
            EC = nucl.me2 / (22. * Ti)
            Emax = min(m._sE0, nucl.E_EC_max * EC)
            Fph = LogInterp(E_grid, F[it, 0])
            rate_E0 = fx["rate_photon_E0"][it]
            for rid in range(1, 18):
                val = S0[it, 0] * nr.get_cross_section(rid, m._sE0) / rate_E0
                Eth = nucl._eth[rid]
                if Emax > Eth:
                    knots = [log(E) for E in E_grid if Eth < E < Emax]
                    if rid == 15:
                        knots.append(log(1.586627 + 2.2951))   # near the sign change of the closing polynomial: extra break point
                    knots = sorted(k for k in knots if log(Eth) < k < log(Emax))
                    with warnings.catch_warnings():
                        warnings.simplefilter("ignore")
                        I = quad(lambda y: Fph(exp(y)) * exp(y) * nr.get_cross_section(rid, exp(y)), log(Eth), log(Emax),
                                 epsrel=1e-12, epsabs=0, limit=2000, points=knots if knots else None)[0]
                    val += I
                pdi[it, rid - 1] = max(nucl.approx_zero, val)
        out[tag + "_pdi_exact"] = pdi
    np.savez_compressed(os.path.join(OUT, "pdi_exact.npz"), **out)


def case_converged():
    """'Converged reference' answers for the five tight points: every quadrature downstream of the (tight, eps = 1e-9) cascade
    spectra redone with the REFERENCE'S OWN integrand functions handed to scipy.integrate.quad with break points at the grid
    knots and a subdivision limit that is never reached -- pdi rates (nucl.py:428-451: Fph_s, limit 2000), the T-integrals of
    MatrixGenerator (nucl.py:563-614: _pdi_kernel_ij / _dcy_kernel_ij, limit 20000, break points at the T-grid knots), then
    expm and the decay matrices as in models.py:84-89.  The reference itself at eps = 1e-9 keeps quad's default limit = 50
    (pdi) / 100 (T) without break points and stops at 2e-4 / 3e-6 on these kinked integrands (IntegrationWarning); the
    fixture separates that from the CUDA path's own error."""
    import numpy as np
    import warnings
    from math import log, exp
    from scipy.integrate import quad
    from scipy.linalg import expm
    cascade, nucl = _import_reference()
    from acropolis.utils import LogInterp
    from acropolis.input import InputInterface, locate_sm_file
    ii = InputInterface(locate_sm_file())
    nr = nucl.NuclearReactor(None, None, (1e-3, 1e-2), 10., ii)
    out = {}
    for tag in ("cfg1_decay_tight", "decay_ee_tight", "cfg2_annih_tight", "annih_pwave_tight", "resonance_b3_tight"):
        fx = np.load(os.path.join(OUT, tag + ".npz"))
        E0, E_grid, T, F, S0 = float(fx["E0"]), fx["E_grid"], fx["T"], fx["F"], fx["S0"]
        pdi = np.zeros((len(T), 17))
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            for it, Ti in enumerate(T):
                EC = nucl.me2 / (22. * Ti)
                Emax = min(E0, nucl.E_EC_max * EC)
                Fph = LogInterp(E_grid, F[it, 0])
                for rid in range(1, 18):
                    val = S0[it, 0] * nr.get_cross_section(rid, E0) / fx["rate_photon_E0"][it]
                    Eth = nucl._eth[rid]
                    if Emax > Eth:
                        knots = sorted(log(E) for E in E_grid if Eth < E < Emax)
                        if rid == 15 and Eth < 5.412617347027785 < Emax:
                            knots = sorted(knots + [log(5.412617347027785)])   # sign change of the closing polynomial (nucl.py:328)
                        val += quad(lambda y: Fph(exp(y)) * exp(y) * nr.get_cross_section(rid, exp(y)), log(Eth), log(Emax),
                                    epsrel=1e-12, epsabs=0, limit=2000, points=knots if knots else None)[0]
                    pdi[it, rid - 1] = max(nucl.approx_zero, val)
            mg = nucl.MatrixGenerator(T, {rid: pdi[:, rid - 1] for rid in range(1, 18)}, ii)
            knotsT = [log(t) for t in T[1:-1]]
            mpdi, mdcy = np.zeros((9, 9)), np.zeros((9, 9))
            for i in range(9):
                for j in range(9):
                    if any(mg._pref_ij(nucl._rsig[rid], i, j) != 0 for rid in nucl._lrid):
                        mpdi[i, j] = quad(lambda y: mg._pdi_kernel_ij(i, j, exp(y)) * exp(y), log(T[-1]), log(T[0]),
                                          epsrel=1e-11, epsabs=0, limit=20000, points=knotsT)[0]
                    if any(mg._pref_ij(nucl._dsig[did], i, j) != 0 for did in nucl._ldid):
                        mdcy[i, j] = quad(lambda y: mg._dcy_kernel_ij(i, j, exp(y)) * exp(y), log(T[-1]), log(T[0]),
                                          epsrel=1e-11, epsabs=0, limit=20000, points=knotsT)[0]
        Y0 = ii.bbn_abundances()
        Yf = fx["postd"] @ expm(mpdi + mdcy) @ fx["pred"] @ Y0
        key = tag.replace("_tight", "")
        out[key + "_pdi"], out[key + "_mpdi"], out[key + "_mdcy"], out[key + "_Yf"] = pdi, mpdi, mdcy, Yf
    np.savez_compressed(os.path.join(OUT, "converged_reference.npz"), **out)


def case_scan_deviations():
    """Consumer side of the scan rows (plots.py:40-101): abundance deviations in sigma for the committed reference scan rows
    (scan_decay_5x5.npz, scan_fast.npz) against each observation set of obs.py.  matplotlib is absent in this image and
    plots.py imports it at module level, so empty stand-in modules are registered first (nothing of them is called by
    _get_abundance/_get_deviations)."""
    import types
    import numpy as np
    mpl = types.ModuleType("matplotlib")
    plt = types.ModuleType("matplotlib.pyplot")
    plt.rc = lambda *a, **k: None
    plt.rcParams = {}
    tick = types.ModuleType("matplotlib.ticker")
    tick.FixedLocator = tick.FixedFormatter = object
    mpl.pyplot, mpl.ticker = plt, tick
    sys.modules.update({"matplotlib": mpl, "matplotlib.pyplot": plt, "matplotlib.ticker": tick})
    _import_reference()
    import acropolis.plots as rp
    import acropolis.obs as ro
    out = {}
    rows = {"decay5x5": np.load(os.path.join(OUT, "scan_decay_5x5.npz"))["rows"],
            "fast_decay": np.load(os.path.join(OUT, "scan_fast.npz"))["decay_rows"],
            "fast_annih": np.load(os.path.join(OUT, "scan_fast.npz"))["annih_rows"]}
    for tag, data in rows.items():
        for i in range(9):
            m, d = rp._get_abundance(data, i)
            out["%s_abundance_%d" % (tag, i)] = np.array([m, d])
        for oname in ("pdg2020", "pdg2021", "pdg2022", "pdg2023", "pdg2024", "pdg"):
            with np.errstate(all="ignore"):
                out["%s_%s" % (tag, oname)] = np.array(rp._get_deviations(data.copy(), getattr(ro, oname)))
    out["obs_table"] = np.array([[getattr(ro, o)[k].mean, getattr(ro, o)[k].err] for o in ("pdg2020", "pdg2021", "pdg2022", "pdg2023", "pdg2024")
                                 for k in ("Yp", "DH", "HeD", "LiH")])
    out["tex"] = np.array([rp.tex_title(mphi=10., tau=1e5, temp0=10., n0a=1e-10), rp.tex_label("mchi"), rp.tex_label("n0a"),
                           rp.tex_label("nokey"), rp.tex_title(mchi=10., b=1e-20, tempkd=1.), rp.tex_title(braa=1)])
    np.savez_compressed(os.path.join(OUT, "scan_deviations.npz"), **out)


CASES = {k[5:]: v for k, v in list(globals().items()) if k.startswith("case_")}


def _run(name):
    os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/numba_cache")
    t0 = time.time()
    CASES[name]()
    return name, time.time() - t0


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("cases", nargs="*")
    ap.add_argument("-j", type=int, default=4)
    a = ap.parse_args()
    names = a.cases or list(CASES)
    ctx = get_context("spawn")
    # scans fork their own pools -> run them serially after the single-point cases
    solo = [n for n in names if n.startswith("scan")]
    para = [n for n in names if not n.startswith("scan")]
    with ctx.Pool(a.j, maxtasksperchild=1) as pool:
        for n, dt in pool.imap_unordered(_run, para):
            print("done %-24s %.1fs" % (n, dt), flush=True)
    for n in solo:
        import subprocess
        t0 = time.time()
        subprocess.check_call([sys.executable, __file__, "--solo", n] if False else
                              [sys.executable, "-c", "import sys; sys.path.insert(0, %r); import make_golden as g; g._run(%r)" % (OUT, n)])
        print("done %-24s %.1fs" % (n, time.time() - t0), flush=True)
