"""CPU-only checks: the C-ABI library loads and exports every declared symbol, the host-side mirror of the reference
interface (input tables, model hooks, grids, scan bookkeeping, sharding) reproduces the reference's numbers, and the
product path refuses to run without the CUDA device/library."""
import json
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from conftest import GOLDEN, ROOT, golden, relerr


@pytest.fixture(autouse=True)
def _quiet():
    import acropolis_b200.flags as flags
    flags.verbose = False
    yield


def test_library_exports_every_declared_symbol():
    from acropolis_b200 import _lib
    lib = _lib.load()
    hdr = open(os.path.join(ROOT, "include", "acropolis_b200.h")).read()
    declared = set(re.findall(r"\b(acro_[a-z0-9_]+)\s*\(", hdr)) - {"acro_batch_t", "acro_out_t"}
    assert declared == set(_lib.EXPORTS)
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.acro_abi_version() == 4 == _lib.ABI_VERSION
    p = _lib.Params()
    lib.acro_default_params(p)
    assert (p.quad_order, p.pdi_cell_order, p.Emin, p.approx_zero, p.Ephb_T_max, p.E_EC_max, p.use_db, p.kernel_mode, p.universal) == \
        (64, 8, 1.5, 1e-200, 200.0, 10.0, 1, 0, 0)
    # the test build (cross-check kernel modes) exports the same ABI
    x = _lib.load("xcheck")
    assert x.acro_abi_version() == _lib.ABI_VERSION and all(getattr(x, n) is not None for n in declared)


def test_product_library_ships_without_crosscheck_kernels():
    """The product .so holds the production kernels only; the cross-check instantiations live in the
    test build."""
    from acropolis_b200 import _lib
    def kernels(path):
        out = subprocess.run(["cuobjdump", "-lelf", path], capture_output=True, text=True).stdout  # noqa: F841
        sy = subprocess.run(["cuobjdump", "-res-usage", path], capture_output=True, text=True).stdout
        return sy
    prod, xc = kernels(_lib.LIB_PATH), kernels(_lib.XCHECK_LIB_PATH)
    assert "kernels_mma_kernel" in prod and "kernels_wien_kernel" in prod and "acro14kernels_kernelI" not in prod
    assert "acro14kernels_kernelI" in xc and "kernels_mma_kernel" in xc


def test_library_is_built_for_sm_100a():
    from acropolis_b200 import _lib
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_no_cpu_fallback_and_no_oracle_in_product():
    """The product never imports oracle/, and the engine refuses to start without a CUDA device."""
    import torch
    pkg = os.path.join(ROOT, "acropolis_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert "pyoracle" not in src and "liboracle" not in src and "acro_oracle" not in src, f
    if not torch.cuda.is_available():
        from acropolis_b200.engine import Engine
        from acropolis_b200.input import shared_input_interface
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            Engine.get(shared_input_interface())


def test_input_interface_known_answers():
    from acropolis_b200.input import shared_input_interface
    ii = shared_input_interface()
    k = json.load(open(os.path.join(GOLDEN, "kat_scalars.json")))
    assert float(ii.parameter("eta")) == k["eta"]
    assert np.array_equal(ii.bbn_abundances(), np.array(k["Y0"]))
    for T, v in k["time_of_T"]:
        assert ii.time(T) == v
    for t, v in k["T_of_time"]:
        assert ii.temperature(t) == v
    for T, v in k["dTdt"]:
        assert ii.dTdt(T) == v
    for T, v in k["scale_factor"]:
        assert ii.scale_factor(T) == v
    for T, v in k["hubble_rate"]:
        assert ii.hubble_rate(T) == v
    assert list(ii.temperature_range()) == k["temperature_range"]
    Ts = np.array([T for T, _ in k["scale_factor"]])
    assert np.allclose(ii.scale_factor(Ts), [v for _, v in k["scale_factor"]], rtol=1e-14)
    with pytest.raises(ValueError):
        ii.time(1e9)


def test_rate_table_conversion():
    from acropolis_b200.db import import_data_from_db, in_rate_db
    z = golden("ratedb_spots.npz")
    db = import_data_from_db()
    assert db.shape == tuple(z["shape"]) == (750 * 450, 2)
    assert np.array_equal(db[z["idx"]], z["vals"])
    assert in_rate_db(0.0, -6.0) and in_rate_db(3.0, -1.0) and not in_rate_db(3.0000001, -2) and not in_rate_db(1, -6.1)


DECAY = ["cfg1_decay", "decay_ee", "decay_nemin", "decay_mixed", "decay_lowT", "decay_highT", "decay_big"]


@pytest.mark.parametrize("name", DECAY + ["cfg2_annih", "annih_pwave"])
def test_model_grids_and_sources_match_reference(name):
    """E grid, T grid (incl. the 39/40 truncation), S0, SC and t(Tmax) as the reference produces them."""
    from acropolis_b200.models import DecayModel, AnnihilationModel
    z = golden(name + ".npz")
    cls = DecayModel if name in DECAY else AnnihilationModel
    m = cls(*z["ctor"])
    sp = m.point_spec()
    assert np.array_equal(sp.E_grid, z["E_grid"]) and np.array_equal(sp.T, z["T"])
    assert sp.E0 == float(z["E0"]) and sp.t_at_tmax == float(z["t_at_Tmax"])
    assert tuple(m._sTrg) == tuple(z["Trg"])
    assert relerr(sp.S0, z["S0"], 1e-300) < 1e-11   # exp(-dt/tau) with dt/tau ~ 700 amplifies one ulp of t(T)
    NT, NE = len(sp.T), len(sp.E_grid)
    if sp.sc_mode == 0:
        assert not np.any(z["SC"])
    else:
        SC = sp.sc[:, :, None] * sp.sc_E.T[None, :, :]
        assert SC.shape == (NT, 3, NE) and relerr(SC, z["SC"], 1e-300) < 1e-11
        assert np.array_equal(SC == 0, z["SC"] == 0)
    # the scalar hooks (what user code may call) agree with the vectorised arrays
    it = NT // 2
    assert m._sS0[0](sp.T[it]) == pytest.approx(z["S0"][it, 0], rel=1e-13, abs=0)
    assert m._sSc[0](sp.E_grid[3], sp.T[it]) == pytest.approx(z["SC"][it, 0, 3], rel=1e-13, abs=0)
    assert np.array_equal(m._pred_matrix(), z["pred"]) and np.array_equal(m._postd_matrix(), z["postd"])


def test_hires_grid_sizes():
    import acropolis_b200.params as params
    from acropolis_b200.models import DecayModel
    z = golden("decay_hires_ne60_nt50.npz")
    old = (params.NE_pd, params.NT_pd)
    try:
        params.NE_pd, params.NT_pd = 60, 50
        sp = DecayModel(*z["ctor"]).point_spec()
    finally:
        params.NE_pd, params.NT_pd = old
    assert np.array_equal(sp.E_grid, z["E_grid"]) and np.array_equal(sp.T, z["T"])


def test_scan_bookkeeping():
    from functools import partial
    from acropolis_b200.models import DecayModel
    from acropolis_b200.scans import BufferedScanner, ScanParameter
    from acropolis_b200.pprint import AcropolisError
    z = golden("scan_decay_5x5.npz")
    sc = BufferedScanner(DecayModel, mphi=ScanParameter(0, 3, 5), tau=ScanParameter(3, 10, 5), temp0=10., n0a=1e-10,
                         bree=0., braa=1.)
    pts = sc.scan_points()
    assert len(pts) == 25
    assert np.array_equal(np.array([[p["mphi"], p["tau"]] for p in pts]), z["rows"][:, :2])
    zf = golden("scan_fast.npz")
    sf = BufferedScanner(partial(DecayModel, temp0=10.), mphi=ScanParameter(0.1, 1.6, 4), tau=1e7,
                         n0a=ScanParameter(-14, -3, 8, fast=True), bree=0., braa=1.)
    pts = sf.scan_points()
    assert np.array_equal(np.array([[p["mphi"], p["n0a"]] for p in pts]), zf["decay_rows"][:, :2])
    assert np.array_equal(ScanParameter(0, 1, 3, spacing="lin").get_range(), np.linspace(0, 1, 3))
    with pytest.raises(AcropolisError):
        BufferedScanner(DecayModel, mphi=ScanParameter(0, 3, 5), tau=1e5, temp0=10., n0a=1e-10, bree=0., braa=1.)
    with pytest.raises(AcropolisError):
        BufferedScanner(dict, mphi=ScanParameter(0, 3, 5), tau=ScanParameter(3, 10, 5))
    # trivial rows (E0 <= Emin) need no device: mphi = 1 MeV
    m = DecayModel(1.0, 1e5, 10., 1e-10, 0., 1.)
    assert m.is_trivial()
    row = np.array(sc._row({"mphi": 1.0, "tau": 1e3}, m.run_disintegration()))
    assert relerr(row[2:], z["rows"][0, 2:], 1e-30) < 1e-15


def test_assemble_offsets():
    from acropolis_b200.engine import Engine, PointSpec
    rng = np.random.default_rng(0)
    pts = []
    for ne, nt in ((10, 3), (17, 5), (33, 2)):
        pts.append(PointSpec(5.0, np.sort(rng.random(ne)) + 1.5, np.sort(rng.random(nt)), rng.random((nt, 3)), 1.0))
    a, c = Engine.assemble(pts)
    assert c["n_points"] == 3 and c["n_systems"] == 10 and c["n_energy_total"] == 60 and c["ne_max"] == 33 and c["nt_max"] == 5
    assert c["n_sys_energy_total"] == 10 * 3 + 17 * 5 + 33 * 2
    # every system's packed triangle is padded to a multiple of 4 entries: 55 -> 56, 153 -> 156, 561 -> 564
    assert c["n_pairs_total"] == 56 * 3 + 156 * 5 + 564 * 2 and c["n_pairs_points"] == 56 + 156 + 564
    assert list(a["pt_pair_off"]) == [0, 56, 56 + 156]
    assert np.all(a["sys_k_off"] % 4 == 0)
    assert list(a["pt_e_off"]) == [0, 10, 27] and list(a["pt_sys_off"]) == [0, 3, 8]
    assert list(a["sys_point"]) == [0] * 3 + [1] * 5 + [2] * 2
    assert list(a["sys_f_off"][:5]) == [0, 10, 20, 30, 47] and a["sys_k_off"][3] == 168 and a["sys_k_off"][4] == 168 + 156
    assert c["sc_mode"] == 0


def test_deal_snake_partition_and_balance():
    """Cost-balanced dealing of a plain scan to the ranks (BufferedScanner._scan_plain under torchrun)."""
    from acropolis_b200.scans import deal_snake, BufferedScanner, ScanParameter
    from acropolis_b200.models import DecayModel, AnnihilationModel
    b = BufferedScanner(DecayModel, mphi=ScanParameter(0., 3., 40), tau=ScanParameter(3., 10., 30), temp0=10., n0a=1e-10, bree=0., braa=1.)
    pts = b.scan_points()
    costs = b._point_costs(pts)
    assert costs.shape == (1200,) and costs.min() == 1.0          # below-threshold points cost (almost) nothing
    for world in (2, 3, 8):
        sh = deal_snake(costs, world)
        assert sorted(np.concatenate(sh).tolist()) == list(range(1200))
        loads = np.array([costs[x].sum() for x in sh])
        assert loads.max() / loads.mean() < 1.02
    # the proxy is the grid the model will really use
    from acropolis_b200.engine import energy_grid
    for i in (0, 400, 777, 1199):
        m = b._build(pts[i])
        ne = 0 if m.is_trivial() else len(energy_grid(m._sE0))
        assert costs[i] == 40. * ne * ne + 1.
    a = BufferedScanner(AnnihilationModel, mchi=ScanParameter(0., 3., 5), a=ScanParameter(-30., -25., 4), b=0., tempkd=1., bree=1., braa=0.)
    assert a._point_costs(a.scan_points()).shape == (20,)

    class Mine(DecayModel):
        pass
    assert BufferedScanner(Mine, mphi=ScanParameter(0., 3., 3), tau=ScanParameter(3., 10., 3), temp0=10., n0a=1e-10, bree=0.,
                           braa=1.)._point_costs([dict(mphi=1., tau=1.)]) is None      # unknown class: round-robin


def test_shard_indices_partition_and_balance():
    from acropolis_b200.scans import shard_indices, point_cost
    rng = np.random.default_rng(1)
    ne = rng.integers(10, 303, size=500)
    costs = [point_cost(n, 40) for n in ne]
    for k in (1, 2, 4, 8):
        sh = shard_indices(costs, k)
        assert sorted(np.concatenate(sh).tolist()) == list(range(500))
        loads = np.array([sum(costs[i] for i in s) for s in sh])
        assert loads.max() / loads.mean() < 1.02


def _gloo_worker(rank, world, port, q, n=11):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from acropolis_b200.scans import gather_rows, deal_snake
    dist.init_process_group("gloo", rank=rank, world_size=world)
    width = 29
    mine = deal_snake(np.arange(n, dtype=np.float64) + 1.0, world)[rank]      # the dealing perform_scan uses under torchrun
    rows = np.array([[i * 100.0 + c for c in range(width)] for i in mine]).reshape(len(mine), width)
    full = gather_rows(mine, rows, n, width)
    q.put((rank, full))
    dist.barrier()
    dist.destroy_process_group()


def test_gather_rows_gloo_world2():
    """N>1 path of perform_scan: ranks compute interleaved shards, rows are all-gathered (gloo, world_size 2)."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
    expect = np.array([[i * 100.0 + c for c in range(29)] for i in range(11)])
    assert np.array_equal(got[0], expect) and np.array_equal(got[1], expect)


def test_gather_rows_gloo_with_an_empty_rank():
    """more ranks than points -- the rank without points must still enter the gather (it used to raise in
    reshape(0, -1) and leave the others hanging in all_gather).  World size 3, two rows."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_gloo_worker, args=(r, 3, port, q, 2)) for r in range(3)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(3))
    for p in procs:
        p.join(timeout=60)
    expect = np.array([[i * 100.0 + c for c in range(29)] for i in range(2)])
    assert all(np.array_equal(got[r], expect) for r in range(3))


def test_fsr_shape_matches_scalar_hook():
    from acropolis_b200.models import DecayModel, _fsr_shape
    m = DecayModel(30., 1e6, 10., 1e-9, 1., 0.)
    E = np.array([1.5, 3.0, 14.9, 14.99, 15.0])
    T = 1e-3
    ref = np.array([m._source_photon_c(e, T) for e in E])
    assert np.allclose(_fsr_shape(E, 15.0) * m._source_electron_0(T), ref, rtol=1e-14, atol=0)
    assert ref[-1] == 0.0


def test_resonance_model_hooks_match_reference():
    """ResonanceModel (ext/models.py:139-325): T_kd from estimate_tempkd_ee, the Breit-Wigner thermal average and the
    resulting source arrays, fixed-order vectorised quadrature vs the reference's adaptive quad calls."""
    from acropolis_b200.ext.benchmarks import BenchmarkModel1, BenchmarkModel2, BenchmarkModel3
    kw = dict(mchi=10., delta=1e-2, gammad=1e-3, gammav=1e-12)
    z, z12 = golden("resonance_b3.npz"), golden("resonance_b12_hooks.npz")
    m3 = BenchmarkModel3(**kw)
    assert m3._sTkd == pytest.approx(float(z["tempkd"]), rel=1e-7)
    assert np.allclose(m3._sigma_v_array(z["sigma_v_T"]), z["sigma_v"], rtol=1e-6, atol=0)
    assert m3._sigma_v(float(z["sigma_v_T"][3])) == pytest.approx(float(z["sigma_v"][3]), rel=1e-6)
    sp = m3.point_spec()
    assert np.array_equal(sp.E_grid, z["E_grid"]) and np.array_equal(sp.T, z["T"])
    assert relerr(sp.S0, z["S0"], 1e-300) < 1e-5          # the reference's own quad noise (eps=1e-3) is ~5e-7 here
    SC = sp.sc[:, :, None] * sp.sc_E.T[None, :, :]
    assert relerr(SC, z["SC"], 1e-300) < 1e-5
    for tag, M in (("b1", BenchmarkModel1), ("b2", BenchmarkModel2)):
        m = M(**kw)
        assert m._sTkd == pytest.approx(float(z12[tag + "_tempkd"]), rel=1e-7)
        assert np.allclose(m._sigma_v_array(z12[tag + "_sigma_v_T"]), z12[tag + "_sigma_v"], rtol=1e-6, atol=0)
    from acropolis_b200.pprint import AcropolisError
    from acropolis_b200.ext.models import ResonanceModel
    with pytest.raises(AcropolisError):
        ResonanceModel(10., 2.0, 1e-3, 1e-12, 1, 1e-3)     # delta > 1


def test_unpermute_outputs_of_a_cost_sorted_chunk():
    """Engine.stage sorts a chunk heaviest-first; Engine.unpermute must put per-point, per-system and per-(system,E)
    outputs back into the caller's order."""
    from acropolis_b200.engine import Engine
    rng = np.random.default_rng(3)
    ne_in = np.array([5, 9, 3, 7]); nt_in = np.array([2, 1, 4, 3])          # caller's order
    perm = np.argsort(-(nt_in * ne_in ** 2), kind="stable")                  # as in Engine.stage
    ne, nt = ne_in[perm], nt_in[perm]                                        # sorted order (what the device saw)
    per_point = [rng.normal(size=(9, 3)) for _ in range(4)]
    per_sys = [rng.normal(size=(t, 17)) for t in nt_in]
    per_f = [rng.normal(size=3 * e * t) for e, t in zip(ne_in, nt_in)]
    Yf = np.stack([per_point[i] for i in perm])
    pdi = np.concatenate([per_sys[i] for i in perm])
    F = np.concatenate([per_f[i] for i in perm])
    assert np.array_equal(Engine.unpermute("Yf", Yf, perm, ne, nt), np.stack(per_point))
    assert np.array_equal(Engine.unpermute("pdi", pdi, perm, ne, nt), np.concatenate(per_sys))
    assert np.array_equal(Engine.unpermute("F", F, perm, ne, nt), np.concatenate(per_f))
    assert np.array_equal(Engine.unpermute("Yf", Yf, None, ne, nt), Yf)


def test_vectorised_point_specs_equal_the_per_model_ones():
    """DecayModel.point_specs (one numpy pass over all points of a scan group) against AbstractModel.point_spec (the
    reference's per-model source hooks): same grids and source arrays bit for bit, photon and e+e- injection."""
    from acropolis_b200.models import DecayModel
    rng = np.random.default_rng(5)
    models = []
    for mphi, tau in zip(np.logspace(0.6, 3, 9), 10.**rng.uniform(3, 10, 9)):
        models.append(DecayModel(float(mphi), float(tau), 10., 1e-10, 0., 1.))
        models.append(DecayModel(float(mphi), float(tau), 5., 3e-9, 0.7, 0.3))
    vec = DecayModel.point_specs(models)
    assert len(vec) == len(models)
    for m, v in zip(models, vec):
        s = m.point_spec()
        assert v.E0 == s.E0 and v.t_at_tmax == s.t_at_tmax and v.sc_mode == s.sc_mode
        for name in ("E_grid", "T", "S0", "sc", "sc_E"):
            a, b = getattr(v, name), getattr(s, name)
            assert (a is None) == (b is None), name
            if a is not None:
                assert np.array_equal(np.asarray(a), np.asarray(b)), name


def test_chunks_respect_budget_and_point_limit():
    """Engine.chunks: greedy split by workspace bytes, and never more points per chunk than the kernels' blockIdx.y can
    index."""
    from acropolis_b200.engine import Engine, PointSpec
    eng = object.__new__(Engine)
    E, T = np.linspace(1.5, 5., 10), np.linspace(1e-3, 2e-3, 2)
    pts = [PointSpec(5., E, T, np.zeros((2, 3)), 0.)] * 70000
    ch = eng.chunks(pts, budget=1 << 40)
    assert [len(c) for c in ch] == [32768, 32768, 4464] and ch[1][0] == 32768
    one = Engine.point_cost_bytes(10, 2)
    ch = eng.chunks(pts[:10], budget=3 * one)
    assert [len(c) for c in ch] == [3, 3, 3, 1]


def test_neutron_fold_leaves_the_transfer_unchanged():
    """models._fold_neutron: post-decay x expm(M) x pre-decay equals the same product with the free-neutron decay folded out
    of M (column n := 0, row p += row n, row n := 0) -- checked with SciPy's expm on reference matrices whose norm is small
    enough for the plain exponential to be accurate -- and a model with its own pre-/post-decay is never folded."""
    from scipy.linalg import expm
    from acropolis_b200.models import DecayModel, _fold_neutron, _neutron_foldable
    z = golden("cfg1_decay.npz")
    mpdi, mdcy = z["mpdi"], z["mdcy"]
    m = DecayModel(10., 1e5, 10., 1e-10, 0., 1.)
    assert _neutron_foldable(m, mpdi, mdcy)
    pred, postd = m._pred_matrix(), m._postd_matrix()
    Y0 = m._sII.bbn_abundances()
    for f in (1., 1e3, 1e6):
        M = f * mpdi + mdcy
        Mf = f * _fold_neutron(mpdi) + _fold_neutron(mdcy)
        assert np.abs(Mf).sum(0).max() < np.abs(M).sum(0).max()
        a, b = postd @ expm(M) @ pred @ Y0, postd @ expm(Mf) @ pred @ Y0
        assert relerr(b, a, 1e-30) < 1e-11
    bad = mpdi.copy()
    bad[5, 0] = 1e-3                                # a reaction with the neutron as target: not foldable
    assert not _neutron_foldable(m, bad, mdcy)

    class OwnDecays(DecayModel):
        def _pred_matrix(self):
            return np.identity(9)
    assert not _neutron_foldable(OwnDecays(10., 1e5, 10., 1e-10, 0., 1.), mpdi, mdcy)


def test_annihilation_point_specs_equal_the_per_model_ones():
    """AnnihilationModel.point_specs (group form) against point_spec(): bit-equal grids and source arrays, s- and p-wave,
    photon and e+e- channels, kinetic decoupling inside and outside the temperature range."""
    from acropolis_b200.models import AnnihilationModel
    models = []
    for mchi in np.logspace(0.4, 2.5, 5):
        models.append(AnnihilationModel(float(mchi), 1e-25, 0., 1., 1., 0.))
        models.append(AnnihilationModel(float(mchi), 0., 3e-21, 1e-3, 0.4, 0.6, omegah2=0.1))
        models.append(AnnihilationModel(float(mchi), 2e-26, 1e-22, 1e-7, 0., 1.))
    vec = AnnihilationModel.point_specs(models)
    assert len(vec) == len(models)
    for m, v in zip(models, vec):
        s = m.point_spec()
        assert v.E0 == s.E0 and v.t_at_tmax == s.t_at_tmax and v.sc_mode == s.sc_mode
        for name in ("E_grid", "T", "S0", "sc", "sc_E"):
            a, b = getattr(v, name), getattr(s, name)
            assert (a is None) == (b is None), name
            if a is not None:
                assert np.array_equal(np.asarray(a), np.asarray(b)), name


# ---- consumer side of the scan rows: plots.py / obs.py numbers ---------------------------------

def test_scan_deviations_match_reference():
    """_get_abundance/_get_deviations on the committed reference scan rows against what the reference's plots.py returned for
    them (tests/golden/make_golden.py case scan_deviations), every observation set of obs.py; labels as the reference's."""
    import acropolis_b200.plots as plots
    import acropolis_b200.obs as obs
    z = np.load(os.path.join(GOLDEN, "scan_deviations.npz"))
    rows = {"decay5x5": np.load(os.path.join(GOLDEN, "scan_decay_5x5.npz"))["rows"],
            "fast_decay": np.load(os.path.join(GOLDEN, "scan_fast.npz"))["decay_rows"],
            "fast_annih": np.load(os.path.join(GOLDEN, "scan_fast.npz"))["annih_rows"]}
    sets = ("pdg2020", "pdg2021", "pdg2022", "pdg2023", "pdg2024")
    table = np.array([[getattr(obs, o)[k].mean, getattr(obs, o)[k].err] for o in sets for k in ("Yp", "DH", "HeD", "LiH")])
    assert np.array_equal(table, z["obs_table"]) and obs.pdg is obs.pdg2024
    for tag, data in rows.items():
        for i in range(9):
            assert np.array_equal(np.array(plots._get_abundance(data, i)), z["%s_abundance_%d" % (tag, i)])
        for o in sets + ("pdg",):
            got = np.array(plots._get_deviations(data.copy(), getattr(obs, o)))
            assert np.allclose(got, z["%s_%s" % (tag, o)], rtol=1e-13, atol=0, equal_nan=True), (tag, o)
    tex = [plots.tex_title(mphi=10., tau=1e5, temp0=10., n0a=1e-10), plots.tex_label("mchi"), plots.tex_label("n0a"),
           plots.tex_label("nokey"), plots.tex_title(mchi=10., b=1e-20, tempkd=1.), plots.tex_title(braa=1)]
    assert tex == [str(t) for t in z["tex"]]


def test_scan_deviations_grid_and_exclusion():
    """scan_deviations(): shapes follow the slow-parameter-first row order, 'overall' combines |D/H|, |Yp| and the one-sided
    He3/D, fix_helium closes holes down a column, and a table without hydrogen maps to D/H = -10."""
    import acropolis_b200.plots as plots
    data = np.load(os.path.join(GOLDEN, "scan_fast.npz"))["decay_rows"]          # 4 mphi x 8 n0a
    d = plots.scan_deviations(data)
    assert d["x"].shape == (4, 8) and np.all(d["x"][:, 0] == d["x"][:, -1]) and np.all(d["y"][0] == d["y"][-1])
    assert np.array_equal(d["overall"], np.maximum(np.maximum(np.abs(d["DH"]), np.abs(d["Yp"])), d["HeD"]))
    assert np.array_equal(d["excluded"], d["overall"] > 1.95996)
    # n0a = 1e-3 destroys everything once E0 = mphi/2 is above the D threshold (the two heavier masses), 1e-14 nothing
    assert d["excluded"][2:, -1].all() and not d["excluded"][:2].any() and not d["excluded"][:, 0].any()
    H = np.array([[0., 3., 1., 3., 1.], [0., 0., 3., 0., 0.]]).T                  # columns = fixed second parameter
    F = plots._fix_helium(H.copy())
    assert np.array_equal(F[:, 0], [0., 3., 10., 3., 10.]) and np.array_equal(F[:, 1], [0., 0., 3., 10., 10.])
    empty = data.copy()
    empty[:, 2:] = 0.
    with np.errstate(all="ignore"):
        Yp, DH, HeD, LiH = plots._get_deviations(empty, plots.pdg2022)
    assert np.all(DH == -10) and np.all(np.isnan(HeD))        # as the reference: only D/H is repaired (plots.py:96-98)
    nod = data.copy()
    nod[:, [4, 13, 22]] = 0.                                   # no deuterium left: D/H below its error -> He3/D excluded
    with np.errstate(all="ignore"):
        assert np.all(plots._get_deviations(nod, plots.pdg2022)[2] == 10)
    with pytest.raises(ImportError, match="matplotlib"):
        plots.plot_scan_results(data)


def test_icphoton_tables_against_quadrature():
    """The generated tables of the closed-form e -> gamma kernel (csrc/acro_icphoton_tab.h): the double-precision evaluation
    that ic_photon_closed_kernel performs (tools/make_icphoton_tables.py::phi_double, same operation order) against adaptive
    quadrature of the original integrals  int F-part * v/(e^v - 1) dv, in the regimes the kernel uses them: shared-grid range
    (v_a < e^2), Wien tail, upper-limit bracket at the thermal cut-off, short range."""
    from scipy.integrate import quad
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import make_icphoton_tables as g
    tabs = g.load_header()
    assert tabs["A"].shape == (43, 12) and tabs["M"].shape == (7, 12)

    def exact(va, vb):
        h = lambda q: 2 * q * np.log(q) + (1 + 2 * q) * (1 - q)
        bose = lambda s: 1.0 / (-np.expm1(-(va + s)))
        S = vb - va
        brk = [x for x in (0.25, 1.0, 4.0, 12.0, 30.0, 70.0) if x < S]
        f0 = quad(lambda s: h(va / (va + s)) * (va + s) * np.exp(-s) * bose(s), 0, S, points=brk, epsabs=0, epsrel=1e-13, limit=400)[0]
        f1 = quad(lambda s: s * np.exp(-s) * bose(s), 0, S, points=brk, epsabs=0, epsrel=1e-13, limit=400)[0]
        return f0 * np.exp(-va), f1 * np.exp(-va)

    for va, vb in ((1e-7, 200.0), (3e-3, 150.0), (0.4, 30.0), (2.5, 200.0), (7.0, 40.0), (7.0, 200.0), (12.0, 200.0), (60.0, 200.0),
                   (150.0, 200.0), (180.0, 200.0), (199.0, 200.0), (199.7, 200.0), (5.0, 9.0), (0.05, 8.0)):
        e0, e1 = exact(va, vb)
        g0, g1 = g.phi_double(tabs, va, vb)
        tol = 1e-9 if vb - va < 2 else 1e-11
        assert abs(g0 / e0 - 1) < tol and abs(g1 / e1 - 1) < tol, (va, vb, g0 / e0 - 1, g1 / e1 - 1)


def test_bench_reference_arm_contract():
    """bench.py --impl reference needs no GPU (the CPU port of the reference's algorithm on the host cores): one JSON line with
    the driver's keys, e2e equal to the line's own value with zero transfer bytes, a cpu_baseline that describes the run; under
    torchrun every rank but 0 exits 0 without a line."""
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads(r.stdout.strip().splitlines()[-1])
    assert d["impl"] == "reference" and d["unit"] == "points/s" and d["higher_is_better"] is True and d["n_gpus"] == 1
    assert d["metric"].startswith("scan points/sec") and d["config"]["name"] == "cfg3" and d["value"] > 0
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "27" in cb["sample"]
    env1 = dict(env, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT="29533")
    r1 = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                        capture_output=True, text=True, timeout=120, env=env1, cwd=ROOT)
    assert r1.returncode == 0 and r1.stdout.strip() == ""


def test_run_points_groups_and_lanes_with_a_stub_engine(monkeypatch):
    """scans.run_points without a GPU: the engine is replaced by a stub that records what is submitted to which lane and
    answers Yf[p] = f(E0, number of temperatures).  Checks the pipelining logic itself: the first group is the tail of the list,
    the rest goes heaviest-first in groups of growth x the previous size, consecutive groups alternate between the lanes,
    every GPU point is submitted exactly once, trivial points never reach the engine, and the rows come back in input order."""
    from functools import partial
    import acropolis_b200.scans as scans
    from acropolis_b200.models import DecayModel
    calls = []

    class StubEngine(object):
        engines = {}

        @classmethod
        def get(cls, ii, device=None, variant="product", lane=0):
            return cls.engines.setdefault(lane, cls(lane))

        def __init__(self, lane):
            self.lane, self.last_h2d_bytes, self.last_d2h_bytes = lane, 0, 0

        def submit(self, specs, want=("Yf", "status")):
            calls.append((self.lane, [(s.E0, len(s.T)) for s in specs]))
            self.last_h2d_bytes += 8 * len(specs)
            return [specs]

        def collect(self, pend, want=("Yf", "status")):
            specs = pend[0]
            Y = np.array([[[s.E0 * (1 + i) + 0.001 * len(s.T) + 10 * c for c in range(3)] for i in range(9)] for s in specs])
            return dict(Yf=Y, status=np.zeros(len(specs), dtype=np.int32))

    monkeypatch.setattr(scans, "Engine", StubEngine)
    mphi = np.logspace(0, 3, 30)                      # the first points are below every threshold (E0 = mphi/2 < 1.5)
    tau = np.logspace(3, 10, 13)
    par = [dict(mphi=float(m), tau=float(t)) for m in mphi for t in tau]
    model = partial(DecayModel, temp0=10., n0a=1e-10, bree=0., braa=1.)
    rows = scans.run_points(model, par, devices=[0], first_group=16, growth=4, lanes=2)
    n_trivial = sum(1 for p in par if p["mphi"] / 2 <= 1.5)
    sizes = [len(c[1]) for c in calls]
    assert sizes[0] == 16 and sizes[1] == 64 and sum(sizes) == len(par) - n_trivial and n_trivial > 0
    assert [c[0] for c in calls] == [k % 2 for k in range(len(calls))]                  # lanes alternate
    first = calls[0][1]
    assert sorted(e0 for e0, _ in first) == sorted(p["mphi"] / 2 for p in par[-16:])  # the tail of the list goes first
    cost = [nt * max(int(np.log10(e0 / 1.5) * 120), 10) ** 2 for e0, nt in calls[1][1]]
    assert cost == sorted(cost, reverse=True)                                          # heaviest first within the rest
    assert rows.shape == (len(par), 2 + 27)
    for r, p in zip(rows, par):
        assert r[0] == p["mphi"] and r[1] == p["tau"]
        if p["mphi"] / 2 > 1.5:
            assert abs(r[2] - (p["mphi"] / 2 + 0.001 * 40)) < 0.0011 and abs(r[2 + 9] - r[2] - 10.0) < 1e-9   # Yf[0, col 0] and [0, col 1] of the stub
    assert StubEngine.engines[0].last_h2d_bytes == 8 * (len(par) - n_trivial)          # lane 1's counters are folded into lane 0's
