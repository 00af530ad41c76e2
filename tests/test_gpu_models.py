"""The reference-facing Python surface on the GPU: models, SpectrumGenerator / NuclearReactor / MatrixGenerator views,
BufferedScanner (plain and fast-parameter scans), and the reference's own stored convergence rows."""
from functools import partial

import numpy as np
import pytest

from conftest import golden, relerr

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _quiet():
    import acropolis_b200.flags as flags
    flags.verbose = False
    yield


def test_decay_model_cfg1(gpu_engine):
    """BASELINE config 1: ./decay 10 1e5 10 1e-10 0 1."""
    from acropolis_b200.models import DecayModel
    z = golden("cfg1_decay.npz")
    Yf = DecayModel(10., 1e5, 10., 1e-10, 0., 1.).run_disintegration()
    assert Yf.shape == (9, 3)
    assert relerr(Yf, z["Yf"], 1e-30) < 1e-4
    # SURVEY 8(c) known answers (mean column)
    assert Yf[2, 0] == pytest.approx(1.83746573e-05, rel=1e-4) and Yf[6, 0] == pytest.approx(8.10470514e-15, rel=1e-4)


def test_annihilation_model_cfg2(gpu_engine):
    from acropolis_b200.models import AnnihilationModel
    z = golden("cfg2_annih.npz")
    Yf = AnnihilationModel(10., 1e-25, 0., 0., 1., 0.).run_disintegration()
    assert relerr(Yf, z["Yf"], 1e-30) < 5e-4   # reference T-quadrature noise, see tests/test_oracle.py
    assert relerr(Yf[[1, 2, 4, 5, 7]], z["Yf"][[1, 2, 4, 5, 7]], 1e-30) < 1e-4


def test_resonance_model(gpu_engine):
    """BASELINE config 4 point: BenchmarkModel3(mchi=10, delta=1e-2, gammad=1e-3, gammav=1e-12) (52 s in the reference)."""
    from acropolis_b200.ext.benchmarks import BenchmarkModel3
    z = golden("resonance_b3.npz")
    Yf = BenchmarkModel3(mchi=10., delta=1e-2, gammad=1e-3, gammav=1e-12).run_disintegration()
    assert relerr(Yf, z["Yf"], 1e-30) < 5e-4 and relerr(Yf[[1, 2, 4, 5]], z["Yf"][[1, 2, 4, 5]], 1e-30) < 1e-4


def test_trivial_model_below_threshold(gpu_engine):
    from acropolis_b200.models import DecayModel
    m = DecayModel(2.0, 1e5, 10., 1e-10, 0., 1.)      # E0 = 1 MeV <= Emin (models.py:65-71)
    Y0 = m._sII.bbn_abundances()
    Yf = m.run_disintegration()
    assert Yf[0, 0] == 0 and Yf[1, 0] == Y0[0, 0] + Y0[1, 0] and Yf[7, 0] == Y0[7, 0] + Y0[8, 0]


def test_user_subclass_scalar_hooks(gpu_engine):
    """A user model that overrides scalar hooks goes through the generic (dense SC) path and still matches."""
    from acropolis_b200.models import DecayModel
    z = golden("decay_ee.npz")

    class MyDecay(DecayModel):
        def _source_photon_c(self, E, T):
            return DecayModel._source_photon_c(self, E, T)

    Yf = MyDecay(30., 1e6, 10., 1e-9, 1., 0.).run_disintegration()
    Yv = DecayModel(30., 1e6, 10., 1e-9, 1., 0.).run_disintegration()
    assert relerr(Yf, Yv, 1e-30) < 1e-12 and relerr(Yf, z["Yf"], 1e-30) < 1e-4


def test_spectrum_generator_view(gpu_engine):
    from acropolis_b200.models import DecayModel
    from acropolis_b200.cascade import SpectrumGenerator
    z = golden("cfg1_decay.npz")
    m = DecayModel(10., 1e5, 10., 1e-10, 0., 1.)
    sg = SpectrumGenerator(m._sII)
    T = float(z["T"][20])
    sol = sg.get_spectrum(m._sE0, m._sS0, m._sSc, T, allX=True)
    assert sol.shape == (4, 62) and np.array_equal(sol[0], z["E_grid"])
    assert relerr(sol[1:], z["F"][20]) < 1e-5
    assert sg.get_spectrum(m._sE0, m._sS0, m._sSc, T).shape == (2, 62)
    assert sg.rate_photon(m._sE0, T) == pytest.approx(float(z["rate_photon_E0"][20]), rel=1e-11)
    E, G, K = sg.get_kernels(m._sE0, T)
    assert relerr(K, z["K_it20"]) < 1e-5 and relerr(G, z["G"][20]) < 1e-11


def test_nuclear_reactor_and_matrix_generator_views(gpu_engine):
    from acropolis_b200.models import DecayModel
    from acropolis_b200.nucl import NuclearReactor, MatrixGenerator
    z = golden("cfg1_decay.npz")
    m = DecayModel(10., 1e5, 10., 1e-10, 0., 1.)
    Tr, grids = NuclearReactor(m._sS0, m._sSc, m._sTrg, m._sE0, m._sII).get_pdi_grids()
    assert np.array_equal(Tr, z["T"]) and sorted(grids) == list(range(1, 18))
    pdi = np.column_stack([grids[r] for r in range(1, 18)])
    assert relerr(pdi, z["pdi"], 1e-100) < 1.5e-3
    mg = MatrixGenerator(Tr, grids, m._sII)
    mpdi, mdcy = mg.get_final_matp()
    assert relerr(mdcy, z["mdcy"], 1e-100) < 1e-6 and relerr(mpdi, z["mpdi"], 1e-100) < 1e-3
    Ts, (all_mpdi, all_mdcy) = mg.get_all_matp()
    assert all_mpdi.shape == (40, 9, 9) and np.allclose(all_mpdi[0], mpdi, rtol=1e-12) and np.all(all_mpdi[-1] == 0)


def test_scan_decay_5x5(gpu_engine):
    """cfg3 sub-grid: same log ranges as the 200x200 scan, against the reference's multiprocessing scan."""
    from acropolis_b200.models import DecayModel
    from acropolis_b200.scans import BufferedScanner, ScanParameter
    z = golden("scan_decay_5x5.npz")
    rows = BufferedScanner(DecayModel, mphi=ScanParameter(0, 3, 5), tau=ScanParameter(3, 10, 5), temp0=10., n0a=1e-10,
                           bree=0., braa=1.).perform_scan(cores=-1)
    ref = z["rows"]
    assert rows.shape == ref.shape == (25, 29)
    assert np.array_equal(rows[:, :2], ref[:, :2])
    assert relerr(rows[:, 2:], ref[:, 2:], 1e-30) < 1e-4


def test_scan_fast_parameter(gpu_engine):
    """Fast-parameter scans (scans.py:127-176).  In the strongly excluded corner (n0a -> 1e-3: ||f*mpdi|| up to 7e4) the
    final abundances depend exponentially on mpdi, so the reference's own ~5e-5 quadrature noise on mpdi is amplified
    to ~2e-3 in Yf (the oracle shows the same distance to the reference at eps=1e-3 AND at eps=1e-9).  Rows in the
    weak regime are held to 1.5e-4 against the reference; every row is held to 1e-4 against the converged oracle."""
    from acropolis_b200.models import DecayModel, AnnihilationModel
    from acropolis_b200.scans import BufferedScanner, ScanParameter
    from oracle_helpers import get_oracle
    z = golden("scan_fast.npz")
    rows = BufferedScanner(DecayModel, mphi=ScanParameter(0.1, 1.6, 4), tau=1e7, temp0=10.,
                           n0a=ScanParameter(-14, -3, 8, fast=True), bree=0., braa=1.).perform_scan(cores=4)
    ref = z["decay_rows"]
    assert rows.shape == ref.shape and np.array_equal(rows[:, :2], ref[:, :2])
    err = np.array([relerr(rows[k, 2:], ref[k, 2:], 1e-30) for k in range(len(ref))]).reshape(4, 8)
    assert err[:, :5].max() < 1.5e-4 and err.max() < 5e-3
    o = get_oracle(1e-9)
    n0 = np.logspace(-14, -3, 8)
    for im, mphi in enumerate(np.logspace(0.1, 1.6, 4)):
        m = DecayModel(mphi, 1e7, 10., n0[0], 0., 1.)
        if m.is_trivial():
            continue
        r = o.run_point(m.point_spec())
        for k in (0, 4, 5, 7):
            Yo = o.final_abundances(r["mpdi"] * n0[k] / n0[0], r["mdcy"], m._t_at_tmax())
            assert relerr(rows[im * 8 + k, 2:].reshape(3, 9).T, Yo, 1e-30) < 1e-4
    rows = BufferedScanner(partial(AnnihilationModel, omegah2=0.12), mchi=ScanParameter(0.5, 1.2, 2), a=0.,
                           b=ScanParameter(-23, -10, 6, fast=True), tempkd=1., bree=1., braa=0.).perform_scan()
    ref = z["annih_rows"]
    assert rows.shape == ref.shape and np.array_equal(rows[:, :2], ref[:, :2])
    err = np.array([relerr(rows[k, 2:], ref[k, 2:], 1e-30) for k in range(len(ref))]).reshape(2, 6)
    assert err[:, :3].max() < 1.5e-4 and err.max() < 5e-3


def test_reference_convergence_rows(gpu_engine):
    """The reference's only stored known answers (tools/data/{NE,NT}_pd_{decay,annih}.dat, 98 rows of Y_D mean/high/low
    for `decay 50 1e5 10 1e-8 0 1` and `annihilation 10 1e-24 0 0 1 0`), at resolutions up to NE_pd=500 / NT_pd=200 --
    the full-size configurations (NE up to 610, NT up to 800), all in one batch."""
    import acropolis_b200.params as params
    from acropolis_b200.models import DecayModel, AnnihilationModel
    from acropolis_b200.scans import run_models
    z = golden("reference_convergence_rows.npz")
    old = (params.NE_pd, params.NT_pd)
    specs, refs = [], []
    try:
        for key, cls, args in (("decay", DecayModel, (50., 1e5, 10., 1e-8, 0., 1.)),
                               ("annih", AnnihilationModel, (10., 1e-24, 0., 0., 1., 0.))):
            for which in ("NE", "NT"):
                for row in z["%s_pd_%s" % (which, key)]:
                    params.NE_pd, params.NT_pd = (int(row[0]), 50) if which == "NE" else (150, int(row[0]))
                    m = cls(*args)
                    specs.append(m.point_spec())
                    refs.append(row[1:4])
    finally:
        params.NE_pd, params.NT_pd = old
    out = gpu_engine.run(specs, want=("Yf", "status"))
    Y_D = out["Yf"][:, 2, :]
    refs = np.array(refs)
    assert len(refs) == 98
    # the stored rows have 6 significant digits; the reference's own quadrature noise on Y_D is ~1e-5 relative
    err = np.abs(Y_D / refs - 1)
    assert np.max(err) < 1e-4 and np.median(err) < 1e-5


def test_high_resolution_point_vs_oracle(gpu_engine):
    """BASELINE config 5 flavour: cfg1 at 2x the default resolution (NE_pd=240, NT_pd=40 -> NE=125, NT=80: four
    temperature tiles in the shared-grid kernel), against the oracle at the same resolution."""
    import acropolis_b200.params as params
    from acropolis_b200.models import DecayModel
    from oracle_helpers import get_oracle
    old = (params.NE_pd, params.NT_pd)
    try:
        params.NE_pd, params.NT_pd = 240, 40
        sp = DecayModel(10., 1e5, 10., 1e-10, 0., 1.).point_spec()
    finally:
        params.NE_pd, params.NT_pd = old
    assert len(sp.E_grid) == 125 and len(sp.T) in (79, 80)
    out = gpu_engine.run([sp], want=("F", "pdi", "Yf"))
    ref = get_oracle(1e-3).run_point(sp)
    F = out["F"].reshape(len(sp.T), 3, len(sp.E_grid))
    assert relerr(F, ref["F"]) < 1e-5
    assert relerr(out["pdi"], ref["pdi"], 1e-100) < 1.5e-3
    assert relerr(out["Yf"][0], ref["Yf"], 1e-30) < 1e-4


def test_kernel_modes_agree_on_abundances(gpu_engine, xcheck_engine):
    """The three evaluations of the thermal kernel integrals (production, per-(pair,T), plain Gauss-Legendre) give the
    same spectra and abundances for a ragged batch."""
    import acropolis_b200.params as params
    from oracle_helpers import spec_from_golden
    specs = [spec_from_golden(golden(n + ".npz")) for n in ("decay_nemin", "cfg1_decay", "decay_ee", "decay_highT", "cfg2_annih")]
    res = {0: gpu_engine.run(specs, want=("F", "Yf"))}
    try:
        for mode in (1, 2):          # cross-check modes: test build of the library only
            params.kernel_mode = mode
            res[mode] = xcheck_engine.run(specs, want=("F", "Yf"))
        with pytest.raises(RuntimeError, match="cross-check build"):
            gpu_engine.run(specs[:1], want=("Yf",))        # the product library refuses them loudly
    finally:
        params.kernel_mode = 0
    for mode in (1, 2):
        live = res[0]["F"] > 1e-150
        assert np.max(np.abs(res[mode]["F"][live] / res[0]["F"][live] - 1)) < 1e-7
        assert relerr(res[mode]["Yf"], res[0]["Yf"], 1e-30) < 1e-8


def test_edge_cases(gpu_engine):
    """Empty input, all-trivial scan, a point just above Emin (NE_min grid), and an injection energy above 1 GeV
    (status bit instead of the reference's printed warning, models.py:47-51)."""
    from acropolis_b200 import _lib
    from acropolis_b200.models import DecayModel
    from acropolis_b200.scans import BufferedScanner, ScanParameter, run_points, run_models
    assert run_models([]).shape == (0, 9, 3)
    assert gpu_engine.run([], want=("Yf",))["Yf"].size == 0
    rows = BufferedScanner(DecayModel, mphi=ScanParameter(0, 0.4, 3), tau=ScanParameter(4, 6, 2), temp0=10., n0a=1e-10,
                           bree=0., braa=1.).perform_scan()
    Y0 = DecayModel(1., 1e5, 10., 1e-10, 0., 1.).run_disintegration()
    assert rows.shape == (6, 29) and np.array_equal(rows[:, 2:], np.tile(Y0.T.reshape(-1), (6, 1)))
    m = DecayModel(3.0002, 1e6, 10., 1e-8, 0., 1.)          # E0 = 1.5001 MeV: NE = NE_min = 10 over a 7e-5 wide range
    sp = m.point_spec()
    assert len(sp.E_grid) == 10
    out = gpu_engine.run([sp], want=("Yf", "F", "status"))
    assert np.all(np.isfinite(out["Yf"])) and out["status"][0] == 0
    big = DecayModel(2600., 1e6, 10., 1e-12, 0., 1.)        # E0 = 1.3 GeV: top of the grid leaves rates.db
    out = gpu_engine.run([big.point_spec()], want=("Yf", "status"))
    assert out["status"][0] & _lib.ST_E0_ABOVE_GEV and out["status"][0] & _lib.ST_DIRECT_RATE
    assert np.all(np.isfinite(out["Yf"]))
    r = run_points(DecayModel, [dict(mphi=20., tau=1e6, temp0=10., n0a=1e-9, bree=0., braa=1.)], group=1)
    assert r.shape == (1, 6 + 27)


def test_full_size_slice_invariants(gpu_engine):
    """One 1000-point slice of BASELINE config 3 (the 200x200 mphi x tau scan, as bench.py cuts it) through the public
    scan entry, checked with size-independent properties: (1) nucleon number sum_i A_i Y_i is conserved by the whole chain
    (pre-decay, expm of the photodisintegration + decay matrix, post-decay): 5e-13 measured with the free-neutron decay
    folded out of the matrix before the exponential (csrc/acro_kernels.cu: apply_transfer); without the fold scaling and
    squaring of a matrix with ||mdcy||_1 = 1.45e9 loses it at 7.7e-8, SciPy's expm on the same matrices at 3e-9; (2) points
    below all thresholds return the decayed BBN abundances; (3) a point's row does not depend on what else is in the
    batch (grouping, cost sorting and un-permutation of run_points / Engine.stage)."""
    from acropolis_b200.models import DecayModel
    from acropolis_b200.scans import run_points
    fixed = dict(temp0=10., n0a=1e-10, bree=0., braa=1.)
    grid = [dict(mphi=float(m), tau=float(t)) for m in np.logspace(0, 3, 200) for t in np.logspace(3, 10, 200)]
    par = grid[7::40]
    assert len(par) == 1000
    rows = run_points(partial(DecayModel, **fixed), par)
    assert rows.shape == (1000, 29) and np.all(np.isfinite(rows))
    assert np.array_equal(rows[:, 0], [p["mphi"] for p in par]) and np.array_equal(rows[:, 1], [p["tau"] for p in par])
    Y = rows[:, 2:].reshape(1000, 3, 9)                        # [point, (mean, high, low), nuclide]
    A = np.array([1., 1., 2., 3., 3., 4., 6., 7., 7.])
    m0 = DecayModel(1.0, 1e5, **fixed)
    Y0 = m0._sII.bbn_abundances()                              # [nuclide, 3]
    assert np.max(np.abs((Y @ A) / (A @ Y0)[None, :] - 1.)) < 1e-9
    trivial = np.array([p["mphi"] / 2. <= 1.5 for p in par])
    assert 100 < trivial.sum() < 300
    Yt = m0._squeeze_decays(Y0)
    assert np.array_equal(Y[trivial], np.broadcast_to(Yt.T, (int(trivial.sum()), 3, 9)))
    rng = np.random.default_rng(11)
    for k in rng.choice(np.flatnonzero(~trivial), 12, replace=False):
        Yk = DecayModel(par[k]["mphi"], par[k]["tau"], **fixed).run_disintegration()
        assert relerr(Y[k].T, Yk, 1e-30) < 1e-12, par[k]


def test_cfg5_resolution_heaviest_point(gpu_engine, xcheck_engine):
    """BASELINE config 5 at full size: NE_pd=480, NT_pd=80 and the heaviest point of the scan grid (mphi = 1e3 MeV: NE=1210,
    NT=160, 117 M (E,E',T) entries, four temperature tiles).  The oracle would need hours here, so the production kernel
    builder is checked against the library's independent per-(E,E',T) evaluation (kernel_mode = PER_T: Gauss-Legendre /
    Gauss-Laguerre per entry, no shared grid) and against nucleon-number conservation."""
    import acropolis_b200.params as params
    from acropolis_b200.models import DecayModel
    old = (params.NE_pd, params.NT_pd)
    try:
        params.NE_pd, params.NT_pd = 480, 80
        m = DecayModel(1e3, 1e6, 10., 1e-10, 0., 1.)
        sp = m.point_spec()
    finally:
        params.NE_pd, params.NT_pd = old
    assert len(sp.E_grid) in (1210, 1211) and len(sp.T) in (159, 160)
    res = {0: gpu_engine.run([sp], want=("F", "pdi", "Yf", "status"))}
    try:
        params.kernel_mode = 2
        res[2] = xcheck_engine.run([sp], want=("F", "pdi", "Yf", "status"))
    finally:
        params.kernel_mode = 0
    assert res[0]["status"][0] == 0 and np.all(np.isfinite(res[0]["F"])) and np.all(np.isfinite(res[0]["Yf"]))
    live = res[2]["F"] > 1e-150
    assert np.max(np.abs(res[0]["F"][live] / res[2]["F"][live] - 1)) < 1e-6
    assert relerr(res[0]["pdi"], res[2]["pdi"], 1e-100) < 1e-6
    assert relerr(res[0]["Yf"], res[2]["Yf"], 1e-30) < 1e-7
    A = np.array([1., 1., 2., 3., 3., 4., 6., 7., 7.])
    Y0 = m._sII.bbn_abundances()
    assert np.max(np.abs((A @ res[0]["Yf"][0]) / (A @ Y0) - 1.)) < 1e-8


# ---------------------------------------------------------------------------------------------------------------
# round 2: BenchmarkModel1/2, get_all_matp, flags.universal -- against fixtures generated from the reference
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("tag, cls", [("b1", "BenchmarkModel1"), ("b2", "BenchmarkModel2")])
def test_benchmark_models_1_and_2(gpu_engine, tag, cls):
    """BASELINE config 4 points: BenchmarkModel1/2(mchi=10, delta=1e-2, gammad=1e-3, gammav=1e-12) (ext/benchmarks.py:50-51)
    through the model hooks, all stages against the reference at its defaults."""
    import acropolis_b200.ext.benchmarks as bm
    z = golden("resonance_%s_full.npz" % tag)
    m = getattr(bm, cls)(mchi=10., delta=1e-2, gammad=1e-3, gammav=1e-12)
    assert m._sTkd == pytest.approx(float(z["tempkd"]), rel=1e-8)
    sp = m.point_spec()
    assert np.array_equal(sp.E_grid, z["E_grid"]) and np.allclose(sp.T, z["T"], rtol=1e-14)
    # the model's source hooks: the reference averages sigma v with quad (ext/models.py:306-316, epsrel 1e-3: 1e-7 of noise),
    # here a fixed-order rule (acropolis_b200/ext/models.py)
    assert relerr(sp.S0, z["S0"], 1e-300) < 1e-6
    r = gpu_engine.run([sp], want=("F", "pdi", "mpdi", "mdcy", "Yf"))
    assert relerr(r["F"].reshape(len(sp.T), 3, -1), z["F"]) < 1e-5
    assert relerr(r["pdi"], z["pdi"], 1e-100) < 7e-4          # the reference's pdi quadrature noise at eps = 1e-3
    assert relerr(r["mdcy"][0], z["mdcy"], 1e-100) < 1e-6
    Yf = m.run_disintegration()
    assert relerr(Yf, r["Yf"][0], 1e-30) < 1e-12
    # as for BenchmarkModel3 / cfg2: 1e-4 on the nuclides the reference's own T-quadrature noise (eps = 1e-3) does not
    # dominate, 5e-4 on Li6 / Li7 (the tight-reference fixtures of the same model family hold 1e-6, test_gpu_parity.py)
    assert relerr(Yf, z["Yf"], 1e-30) < 5e-4 and relerr(Yf[[1, 2, 4, 5]], z["Yf"][[1, 2, 4, 5]], 1e-30) < 1e-4


def test_get_all_matp_vs_reference(gpu_engine):
    """MatrixGenerator.get_all_matp (nucl.py:626-655): rate matrices integrated from Tmax down to every grid temperature,
    on the reference's rates.  1e-5 against the reference with nucl.eps = 1e-9, 5e-4 against its default (eps = 1e-3: the
    T-quadrature noise of the reference itself, 2e-4 between the two fixtures)."""
    from acropolis_b200.nucl import MatrixGenerator, _lrid
    from acropolis_b200.input import shared_input_interface
    z = golden("cfg1_all_matp.npz")
    grids = {rid: z["pdi"][:, rid - 1] for rid in _lrid}
    T, (mp, md) = MatrixGenerator(z["T"], grids, shared_input_interface()).get_all_matp()
    assert np.array_equal(T, z["T"]) and mp.shape == (len(T), 9, 9) and md.shape == (len(T), 9, 9)
    assert np.all(mp[-1] == 0) and np.all(md[-1] == 0)          # T = Tmax: empty range
    assert relerr(mp, z["all_mpdi_tight"], 1e-100) < 1e-5 and relerr(md, z["all_mdcy_tight"], 1e-100) < 1e-7
    assert relerr(mp, z["all_mpdi"], 1e-100) < 5e-4 and relerr(md, z["all_mdcy"], 1e-100) < 1e-6


def test_universal_spectrum_flag(gpu_engine):
    """flags.universal (flags.py:33-41): universal photon spectrum instead of the cascade solution (cascade.py:774-812), pdi
    integrals cut at min(E0, EC) (nucl.py:416-422); spectra, rates, matrices and abundances against the reference."""
    import acropolis_b200.flags as flags
    from acropolis_b200.models import DecayModel
    from acropolis_b200.cascade import SpectrumGenerator
    z = golden("universal.npz")
    try:
        flags.set_universal(True)
        for tag in ("cfg1", "ee"):
            m = DecayModel(*z[tag + "_ctor"])
            sp = m.point_spec()
            r = gpu_engine.run([sp], want=("F", "pdi", "mpdi", "mdcy", "Yf", "status"))
            F = r["F"].reshape(len(sp.T), 3, -1)[:, 0]
            assert relerr(F, z[tag + "_F"], 1e-190) < 1e-10
            assert relerr(r["pdi"], z[tag + "_pdi"], 1e-100) < 7e-4        # reference pdi quadrature at eps = 1e-3
            assert relerr(r["mdcy"][0], z[tag + "_mdcy"], 1e-100) < 1e-6
            assert relerr(r["mpdi"][0], z[tag + "_mpdi"], 1e-100) < 2e-3
            assert relerr(m.run_disintegration(), z[tag + "_Yf"], 1e-30) < 1e-4
            # the reference-facing view (any offset) gives the same numbers
            it = len(sp.T) // 2
            u = SpectrumGenerator(m._sII).get_universal_spectrum(m._sE0, m._sS0, m._sSc, float(sp.T[it]), offset=5e-2)
            assert relerr(u[1], z[tag + "_F"][it], 1e-190) < 1e-10
    finally:
        flags.set_universal(False)
    # and the flag really switches: the cascade solution differs
    r0 = gpu_engine.run([DecayModel(*z["cfg1_ctor"]).point_spec()], want=("pdi",))
    assert relerr(r0["pdi"], z["cfg1_pdi"], 1e-100) > 1e-2


def test_rate_table_regeneration(gpu_engine):
    """the rates.db entries as converged direct integrals on the GPU (acro_direct_rates, tools/make_rates_db.py)
    against (a) the reference's dblquad at eps = 1e-9 and (b) a sub-grid of the shipped table, whose own accuracy the
    reference states as 'interpolation error ~5e-5'  -- the old table generator used QUADPACK at 1e-3."""
    import acropolis_b200.params as params
    from acropolis_b200.db import import_data_from_db
    t = golden("rates_direct_tight.npz")
    E, T = np.meshgrid(t["E"], t["T"], indexing="ij")
    r = gpu_engine.direct_rates(E.reshape(-1), T.reshape(-1)).reshape(len(t["E"]), len(t["T"]), 2)
    thr = E >= params.me2 / (50. * T)               # pair creation is hard-zero below (cascade.py:340)
    assert np.all(r[..., 0][~thr] == 0)
    big = thr & (t["pair_creation"] > 1e-150)
    assert np.max(np.abs(r[..., 0][big] / t["pair_creation"][big] - 1)) < 1e-6
    ok = E <= 1e3                                   # the reference's IC integrand is only trusted to 1 GeV (models.py:47)
    assert np.max(np.abs(r[..., 1][ok] / t["inverse_compton"][ok] - 1)) < 1e-6
    ref = import_data_from_db().reshape(params.Tnum, params.Enum, 2)
    jT, jE = np.arange(0, params.Tnum, 37), np.arange(0, params.Enum, 23)
    Tg = np.logspace(params.Tmin_log, params.Tmax_log, params.Tnum)[jT]
    Eg = np.logspace(params.Emin_log, params.Emax_log, params.Enum)[jE]
    TT, EE = np.meshgrid(Tg, Eg, indexing="ij")
    g = gpu_engine.direct_rates(EE.reshape(-1), TT.reshape(-1)).reshape(len(jT), len(jE), 2)
    sub = ref[np.ix_(jT, jE)]
    below = EE < params.me2 / (50. * TT)            # the shipped table (an older generator) has entries there; the reference
    assert np.all(g[..., 0][below] == 0)            # returns 0 before looking at them (cascade.py:362)
    for c in range(2):
        live = (sub[..., c] > -150) & (g[..., c] > 1e-150)
        assert live.sum() > (400 if c else 150)
        assert np.max(np.abs(g[..., c][live] / 10.0 ** sub[..., c][live] - 1)) < 5e-5     # measured over the whole table: 2.1e-5


def test_scan_rows_are_savetxt_compatible(gpu_engine, tmp_path):
    """The example scripts write the scan table with np.savetxt (examples/scan_pwave_ee:34-37) and plots.py reads it back
    with np.loadtxt: same layout [p1, p2, 9 mean, 9 high, 9 low], exact round trip."""
    from acropolis_b200.models import AnnihilationModel
    from acropolis_b200.scans import BufferedScanner, ScanParameter
    rows = BufferedScanner(AnnihilationModel, mchi=ScanParameter(0.5, 1.5, 3), a=0., b=ScanParameter(-20, -14, 4, fast=True),
                           tempkd=1., bree=1., braa=0.).perform_scan(cores=-1)
    assert rows.shape == (12, 29) and rows.dtype == np.float64
    f = tmp_path / "annih_pwave_Tkd_1e+00MeV_ee.dat"
    np.savetxt(f, rows)
    back = np.loadtxt(f)
    assert np.array_equal(back, rows)
    # sorted-key column order (scans.py:122-124): b, mchi
    assert np.allclose(np.unique(rows[:, 0]), np.logspace(-20, -14, 4)) and np.allclose(np.unique(rows[:, 1]), np.logspace(0.5, 1.5, 3))


def test_run_models_two_devices_one_process(gpu_engine):
    """one process driving several GPUs (run_models(devices=[0, 1]): a ThreadPoolExecutor with one Engine
    per device).  Function attributes (the 212 kB dynamic shared memory of kernels_mma_kernel) are per device, so the second
    device used to fail with 'invalid argument'.  Needs two visible GPUs (gpurun --gpus 2)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from acropolis_b200.models import DecayModel
    from acropolis_b200.scans import run_models
    models = [DecayModel(m, t, 10., 1e-10, 0., 1.) for m in (4., 20., 150., 700.) for t in (1e4, 1e6, 1e8)]
    one = run_models(models, devices=[0])
    two = run_models(models, devices=[0, 1])
    other = run_models(models, devices=[1])
    assert np.array_equal(one, two) and np.array_equal(one, other)
