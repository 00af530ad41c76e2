"""Parameter scans: ``ScanParameter`` / ``BufferedScanner`` with the reference's interface (acropolis/scans.py:16-245)
and result layout (rows ``[scan params sorted by key, 9 mean, 9 high, 9 low]``, scans.py:122-124).

The reference fans the grid out over a ``multiprocessing.Pool`` (scans.py:210-227).  Here the grid is flattened,
trivial points (E0 <= Emin) are answered on the spot, and all remaining points go through the batched CUDA engine
in a few launches per stage; with several GPUs the point list is dealt out cost-balanced (cost ~ NT*NE^2) and the
rows are gathered at the end -- no collective on the hot path.  Two ways to use several GPUs:

* one process, ``perform_scan(devices=[0,1,...])`` (default: every visible device);
* one process per GPU under ``torchrun``: every rank calls ``perform_scan()``; ranks compute their shard and the
  rows are all-gathered so every rank returns the full table (``ACROPOLIS_B200_DIST=0`` disables this)."""
from functools import partial
import gc
import os
from time import time

import numpy as np

from acropolis_b200.pprint import print_info, print_error
from acropolis_b200.models import AbstractModel
from acropolis_b200.engine import Engine
import acropolis_b200.params as params


class ScanParameter(object):

    def __init__(self, ivalue, fvalue, num, spacing="log", fast=False):
        self._sInitialValue = ivalue
        self._sFinalValue = fvalue
        self._sSpacingMode = spacing
        self._sFastParameter = fast
        self._sNumPoints = num

    def get_range(self):
        if self._sSpacingMode == "log":
            return np.logspace(self._sInitialValue, self._sFinalValue, self._sNumPoints)
        elif self._sSpacingMode == "lin":
            return np.linspace(self._sInitialValue, self._sFinalValue, self._sNumPoints)

    def is_fast(self):
        return self._sFastParameter


# ---------------------------------------------------------------------------------------------------------------
# sharding helpers (host logic; covered by gloo world_size-2 tests)
# ---------------------------------------------------------------------------------------------------------------
def point_cost(ne, nt):
    """Relative cost of one point: NT * NE(NE+1)/2 kernel entries (SURVEY.md section 6 cost model)."""
    return float(nt) * ne * (ne + 1) / 2.0


def shard_indices(costs, n_shards):
    """Deal items to shards, heaviest first, always to the currently lightest shard (LPT).  Returns a list of
    index arrays (each sorted ascending) that partition range(len(costs))."""
    costs = np.asarray(costs, dtype=np.float64)
    order = np.argsort(-costs, kind="stable")
    load = np.zeros(n_shards)
    shards = [[] for _ in range(n_shards)]
    for i in order:
        k = int(np.argmin(load))
        shards[k].append(int(i))
        load[k] += costs[i]
    return [np.array(sorted(s), dtype=np.int64) for s in shards]


def deal_snake(costs, n_shards):
    """Deal items to shards heaviest first in snake order (0..n-1, n-1..0, ...): vectorised near-LPT -- every shard's load is
    within one item of the mean for the smooth cost profile of a scan grid.  Returns a list of ascending index arrays that
    partition range(len(costs))."""
    costs = np.asarray(costs, dtype=np.float64)
    order = np.argsort(-costs, kind="stable")
    k = np.arange(len(costs))
    lap, pos = k // n_shards, k % n_shards
    owner = np.where(lap % 2 == 0, pos, n_shards - 1 - pos)
    return [np.sort(order[owner == r]) for r in range(n_shards)]


def _dist_world():
    if os.environ.get("ACROPOLIS_B200_DIST", "1") == "0":
        return None
    try:
        import torch.distributed as dist
    except ImportError:
        return None
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        return dist
    return None


def gather_rows(local_idx, local_rows, n_total, width, dist=None):
    """All-gather per-rank result rows into the full [n_total, width] table (every rank gets it)."""
    dist = _dist_world() if dist is None else dist
    full = np.zeros((n_total, width))
    if dist is None:
        full[local_idx] = local_rows
        return full
    import torch
    ws = dist.get_world_size()
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    n_loc = torch.tensor([len(local_idx)], dtype=torch.int64, device=dev)
    counts = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(ws)]
    dist.all_gather(counts, n_loc)
    n_max = int(max(int(c.item()) for c in counts))
    buf = torch.zeros((n_max, width + 1), dtype=torch.float64, device=dev)
    if len(local_idx):
        buf[:len(local_idx), 0] = torch.from_numpy(np.asarray(local_idx, dtype=np.float64)).to(dev)
        buf[:len(local_idx), 1:] = torch.from_numpy(np.asarray(local_rows, dtype=np.float64).reshape(-1, width)).to(dev)
    bufs = [torch.zeros_like(buf) for _ in range(ws)]
    dist.all_gather(bufs, buf)
    for c, b in zip(counts, bufs):
        n = int(c.item())
        if n:
            b = b[:n].cpu().numpy()
            full[b[:, 0].astype(np.int64)] = b[:, 1:]
    return full


_CHAIN_HOOKS = ("run_disintegration", "_pdi_matrix", "_pred_matrix", "_postd_matrix", "_squeeze_decays")


def uses_builtin_chain(model):
    """True if the model takes the stock route from (mpdi, mdcy) to the abundances (reference models.py:43-165), which the
    device computes for a whole batch.  A user subclass that overrides one of these hooks is honoured in the reference by
    ``_run_single`` calling ``model.run_disintegration()`` (scans.py:118); here such a model is run through its own
    ``run_disintegration()`` too (one engine call per model -- slow but faithful)."""
    cls = type(model)
    return all(getattr(cls, n) is getattr(AbstractModel, n) for n in _CHAIN_HOOKS)


def report_status(status, where):
    """Decode the per-point status words of a batch (include/acropolis_b200.h ACRO_ST_*) into the reference's messages:
    E0 > 1 GeV and an uncovered temperature range are warnings there (models.py:47-61); non-finite abundances have no
    counterpart in the reference (it would print NaN rows) -- they are reported with their count."""
    from acropolis_b200 import _lib
    from acropolis_b200.pprint import print_warning
    status = np.asarray(status, dtype=np.int64).reshape(-1)
    if status.size == 0:
        return 0
    n_e0 = int(np.count_nonzero(status & _lib.ST_E0_ABOVE_GEV))
    n_t = int(np.count_nonzero(status & _lib.ST_T_OUTSIDE))
    n_bad = int(np.count_nonzero(status & _lib.ST_NONFINITE))
    if n_e0:
        print_warning("Injection energy > 1 GeV for %d point(s). Results cannot be trusted." % n_e0, where)
    if n_t:
        print_warning("Temperature range not covered by input data for %d point(s). Results cannot be trusted." % n_t, where)
    if n_bad:
        print_warning("Non-finite abundances for %d point(s)." % n_bad, where)
    return n_bad


def run_models(models, devices=None, want_matp=False):
    """Final abundances of many models at once -> Yf[N, 9, ncol]  (N x ``run_disintegration``, reference
    models.py:43-91).  Trivial models (E0 <= Emin) get the post-decayed input abundances.  With ``want_matp`` the
    per-model (mpdi, mdcy) are returned as well (None for trivial models)."""
    N = len(models)
    if N == 0:
        return np.zeros((0, 9, 3))
    ncol = models[0]._sII.bbn_abundances().shape[1]
    Yf = np.zeros((N, 9, ncol))
    matp = [None] * N
    work = []
    for i, m in enumerate(models):
        if not uses_builtin_chain(m):
            Yf[i] = m.run_disintegration()      # the model's own hooks decide (reference scans.py:118)
            matp[i] = m.get_matp_buffer()
        elif m.is_trivial():
            Yf[i] = m._squeeze_decays(m._sII.bbn_abundances())
        else:
            work.append(i)
    if not work:
        return (Yf, matp) if want_matp else Yf
    wm = [models[i] for i in work]
    specs = type(wm[0]).point_specs(wm) if len({type(m) for m in wm}) == 1 else [m.point_spec() for m in wm]
    import torch
    if devices is None:
        dist = _dist_world()
        devices = [torch.cuda.current_device()] if dist is not None else list(range(torch.cuda.device_count()))
    devices = list(devices)[:max(1, len(work))]
    want = ("Yf", "mpdi", "mdcy", "status") if want_matp else ("Yf", "status")
    if len(devices) == 1:
        out = Engine.get(models[work[0]]._sII, devices[0]).run(specs, want=want)
        parts = [(np.arange(len(work)), out)]
    else:
        shards = shard_indices([point_cost(len(s.E_grid), len(s.T)) for s in specs], len(devices))
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(len(devices)) as ex:
            futs = [ex.submit(Engine.get(models[work[0]]._sII, d).run, [specs[j] for j in sh], want=want)
                    for d, sh in zip(devices, shards) if len(sh)]
            parts = [(sh, f.result()) for sh, f in zip([s for s in shards if len(s)], futs)]
    for sh, out in parts:
        report_status(out["status"], "acropolis_b200.scans.run_models")
        for k, j in enumerate(sh):
            Yf[work[j]] = out["Yf"][k][:, :ncol]
            if want_matp:
                matp[work[j]] = (out["mpdi"][k], out["mdcy"][k])
    return (Yf, matp) if want_matp else Yf


def run_points(model, param_sets, devices=None, group=4096, first_group=40, growth=5, lanes=2):
    """Rows ``[sorted params..., 9 mean, 9 high, 9 low]`` for an arbitrary list of parameter dicts of one model class
    (or ``functools.partial``): the batched form of ``BufferedScanner._run_single`` (reference scans.py:110-124).
    Builds the models, stages their grids and source arrays, runs the CUDA path and reads the abundances back.  On one
    device the list is processed in groups: while the GPU works on one group the host constructs the models and source
    arrays of the next.  The first group is the LAST `first_group` GPU points of the list, launched before anything else
    is constructed (scans are ordered by increasing mass, so the tail is the expensive end); the rest is sorted
    heaviest-first by the cost proxy NT * NE^2 and goes in groups that grow by `growth` up to `group`.  Consecutive groups go to
    different engine lanes (``Engine.get(..., lane=k)``: own library handle, stream and workspace), so their kernels OVERLAP on
    the GPU: a small group alone leaves 40 % of the device idle (wave quantisation and tails of the long-block builder kernels),
    the next group fills it.  Any order is valid -- a list whose tail is cheap only pipelines less well.  Measured on 1000-point
    slices (tools/gpu_e2e_timeline.py; host ~12 us per point, GPU ~45 us on average), wall per call: one lane, groups 64 + 776:
    44.0 ms (two launch sets one after the other cost 3.5 ms more GPU time than one); two lanes, 64 + 256 + 520: 43.0;
    48 + 192 + 600: 42.4; **40 + 200 + 600: 41.8**; 32 + 160 + 648: 42.2; 24 + 120 + 600 + 96: 42.9; 16 + 80 + 400 + 344: 44.4;
    three lanes: no better; one lane with three
    groups: 48 ms.  A single launch set of all 840 points is 38.1 ms of GPU time."""
    import torch
    if devices is None:
        dist = _dist_world()
        devices = [torch.cuda.current_device()] if dist is not None else list(range(max(1, torch.cuda.device_count())))
    devices = list(devices)
    N = len(param_sets)
    if len(devices) != 1 or N == 0:
        models = [model(**p) for p in param_sets]
        Yf = run_models(models, devices=devices)
        return np.array([[*[p[k] for k in sorted(p)], *Y.transpose().reshape(Y.size)] for p, Y in zip(param_sets, Yf)])
    lanes = int(os.environ.get("ACROPOLIS_B200_LANES", lanes))      # 1 = every group on one stream, one after the other
    rows, eng, in_flight = [None] * N, None, []
    # The loop below allocates thousands of small objects per call; a generation-2 collection in the middle of it walks
    # every cached grid and model (tens of ms, seen as 30-50 ms holes in tools/gpu_e2e_timeline.py) while the GPU runs dry.
    # Nothing here creates reference cycles, so the collector rests until the rows are assembled.
    gc_was_on = gc.isenabled()
    gc.disable()
    try:
        return _run_points_pipelined(model, param_sets, devices, group, first_group, growth, rows, eng, in_flight, lanes)
    finally:
        if gc_was_on:
            gc.enable()


def _run_points_pipelined(model, param_sets, devices, group, first_group, growth, rows, eng, in_flight, lanes=2):
    N = len(param_sets)

    # rows [N, n_params + 9 ncol]: the parameter columns are filled while the GPU works, the abundance columns of a group in
    # one vectorised assignment when its results arrive (a Python loop over 1000 rows was ~2 ms after the last kernel)
    keys = sorted(param_sets[0])
    width = [None]
    table = [None]

    def ensure_table(nc):
        if table[0] is None:
            width[0] = len(keys) + 9 * nc
            table[0] = np.empty((N, width[0]))
            table[0][:, :len(keys)] = [[p[key] for key in keys] for p in param_sets]
        return table[0]

    def finish(entry):
        idx, pend, e = entry
        out = e.collect(pend, want=("Yf", "status"))
        statuses.append(out["status"])
        t = ensure_table(ncol)
        t[np.asarray(idx), len(keys):] = out["Yf"][:, :, :ncol].transpose(0, 2, 1).reshape(len(idx), 9 * ncol)

    models, ncol, statuses, direct = [None] * N, 3, [], []
    engines, n_submitted = [], [0]

    def build(i):
        """Model of point i; a point below all thresholds is answered on the spot (a scan may start with hundreds)."""
        nonlocal ncol, eng
        m = models[i] = model(**param_sets[i])
        own = not uses_builtin_chain(m)
        if own or m.is_trivial():
            Y = m.run_disintegration() if own else m._squeeze_decays(m._sII.bbn_abundances())
            ncol = Y.shape[1]
            direct.append((i, Y.transpose().reshape(Y.size)))     # written into the table at the end (keeps the start lean)
            return False
        if eng is None:
            ncol = m._sII.bbn_abundances().shape[1]
            eng = Engine.get(m._sII, devices[0])
            engines.append(eng)
            engines.extend(Engine.get(m._sII, devices[0], lane=k) for k in range(1, max(1, lanes)))
            for e in engines:
                e.last_h2d_bytes = e.last_d2h_bytes = 0
        return True

    def submit(idx):
        todo = [models[i] for i in idx]
        specs = type(todo[0]).point_specs(todo) if len({type(mm) for mm in todo}) == 1 else [mm.point_spec() for mm in todo]
        e = engines[n_submitted[0] % len(engines)]      # consecutive groups on different lanes: their kernels overlap on the GPU
        n_submitted[0] += 1
        in_flight.append((idx, e.submit(specs, want=("Yf", "status")), e))
        while len(in_flight) > 2 * len(engines):
            finish(in_flight.pop(0))

    # 1. first group: the last `first_group` GPU points of the list, before anything else is even constructed
    head, i = [], N - 1
    while i >= 0 and len(head) < min(group, first_group):
        if build(i):
            head.append(i)
        i -= 1
    if head:
        submit(head)
    # 2. the rest: all models, then heaviest first by the cost proxy NT * NE^2 of the grids they imply
    gpu_idx = [k for k in range(i + 1) if build(k)]
    if gpu_idx:
        E0 = np.array([models[k]._sE0 for k in gpu_idx], dtype=np.float64)
        Tr = np.array([models[k]._sTrg for k in gpu_idx], dtype=np.float64)
        ne = np.maximum((np.log10(E0 / params.Emin) * params.NE_pd).astype(np.int64), params.NE_min)
        nt = (np.log10(Tr[:, 1] / Tr[:, 0]) * params.NT_pd).astype(np.int64)
        order = [gpu_idx[k] for k in np.argsort(-(nt * ne * ne), kind="stable")]
        k, target = 0, min(group, growth * max(1, len(head)))
        while k < len(order):
            submit(order[k:k + target])
            k += target
            target = min(group, growth * target)
    for entry in in_flight:
        finish(entry)
    for e in engines[1:]:                                # the caller reads the transfer counters of the primary engine
        eng.last_h2d_bytes += e.last_h2d_bytes
        eng.last_d2h_bytes += e.last_d2h_bytes
    if statuses:
        report_status(np.concatenate(statuses), "acropolis_b200.scans.run_points")
    # A model holds lists of its own bound methods (get_source_terms, as in the reference): a reference cycle per point that
    # only the cyclic collector could free.  These models never leave this function, so the cycles are cut here and the
    # thousands of objects of a call are released by reference counting instead of piling up for a later collection.
    for m in models:
        if m is not None:
            m._sS0 = m._sSc = None
    t = ensure_table(ncol)
    if direct:
        t[np.array([i for i, _ in direct]), len(keys):] = np.array([y for _, y in direct])
    return t


class BufferedScanner(object):

    def __init__(self, model, **kwargs):
        self._sModelClass = model.func if type(model) is partial else model
        if not issubclass(self._sModelClass, AbstractModel):
            print_error(self._sModelClass.__name__ + " is not a subclass of AbstractModel",
                        "acropolis_b200.scans.BufferedScanner.__init__")
        self._sModel = model
        self._sFixed = {}   # fixed parameters
        self._sScanp = {}   # scan parameters
        self._sNP = 0
        self._sNP_fast = 0
        self._sFast_id = None
        self._parse_arguments(**kwargs)
        self._sScanp_id = list(self._sScanp.keys())

    def _parse_arguments(self, **kwargs):
        for key, param in kwargs.items():
            if isinstance(param, ScanParameter):
                self._sNP += 1
                self._sScanp[key] = param.get_range()
                if param.is_fast():
                    self._sNP_fast += 1
                    self._sFast_id = key
            else:
                self._sFixed[key] = param
        if self._sNP_fast > 1 or self._sNP != 2:
            print_error("The class BufferedScanner only supports 2d parameter scans with <= 1 fast parameters!",
                        "acropolis_b200.scans.BufferedScanner._parse_arguments")

    def _rescale_matp_buffer(self, buffer, factor):
        return (factor * buffer[0], buffer[1])

    def _build(self, scanp_set):
        fullp = dict(scanp_set)
        fullp.update(self._sFixed)
        return self._sModel(**fullp)

    @staticmethod
    def _row(scanp_set, Yb):
        return [*[scanp_set[key] for key in sorted(scanp_set)], *Yb.transpose().reshape(Yb.size)]

    @staticmethod
    def _ncol():
        """Abundance columns of the shared input file (rows are [2 + 9*ncol] wide)."""
        from acropolis_b200.input import shared_input_interface, locate_sm_file
        return shared_input_interface(locate_sm_file()).bbn_abundances().shape[1]

    def scan_points(self):
        """The flattened list of parameter dicts in the reference's order (scans.py:193-208)."""
        key1, key2 = self._sScanp_id
        if self._sNP_fast == 0:
            return [{key1: v1, key2: v2} for v1 in self._sScanp[key1] for v2 in self._sScanp[key2]]
        keyp = key1 if key1 != self._sFast_id else key2
        return [{keyp: vp, self._sFast_id: vf} for vp in self._sScanp[keyp] for vf in self._sScanp[self._sFast_id]]

    def perform_scan(self, cores=1, devices=None):
        """``cores`` is accepted for compatibility (the reference's Pool size); the work runs on the GPU(s)."""
        assert self._sNP_fast <= 1 and self._sNP == 2
        start_time = time()
        print_info("Running scan for {} on the GPU.".format(self._sModelClass.__name__),
                   "acropolis_b200.scans.BufferedScanner.perform_scan", verbose_level=3)
        pts = self.scan_points()
        dist = _dist_world()
        rank, world = (dist.get_rank(), dist.get_world_size()) if dist is not None else (0, 1)
        if self._sNP_fast == 0:
            rows = self._scan_plain(pts, rank, world, devices)
        else:
            rows = self._scan_fast(pts, rank, world, devices)
        print_info("Finished after {:.1f}min.".format((time() - start_time) / 60),
                   "acropolis_b200.scans.BufferedScanner.perform_scan", verbose_level=3)
        return rows

    # -- no fast parameter: every point is a full computation (scans.py:110-124) -------------------------------
    def _point_costs(self, pts):
        """Cost proxy NT * NE^2 per grid point WITHOUT building models, for the classes whose injection energy is a
        constructor parameter (DecayModel: mphi / 2, AnnihilationModel: mchi); None for any other class."""
        from acropolis_b200.models import DecayModel, AnnihilationModel
        key, scale, nt = {DecayModel: ("mphi", 0.5, 2. * params.NT_pd), AnnihilationModel: ("mchi", 1.0, 4. * params.NT_pd)}.get(
            self._sModelClass, (None, None, None))
        if key is None or not pts:
            return None
        src = np.array([p[key] for p in pts]) if key in pts[0] else np.full(len(pts), float(self._sFixed.get(key, np.nan)))
        E0 = scale * src
        if not np.all(np.isfinite(E0)):
            return None
        ne = np.where(E0 > params.Emin, np.maximum(np.floor(np.log10(np.maximum(E0, 1e-300) / params.Emin) * params.NE_pd), params.NE_min), 0.)
        return nt * ne * ne + 1.0

    def _scan_plain(self, pts, rank, world, devices):
        N = len(pts)
        mine = np.arange(N)
        if world > 1:
            costs = self._point_costs(pts)
            if costs is None:      # unknown model class: the grid size is not known before the model exists; neighbouring
                mine = np.arange(rank, N, world)      # points cost about the same, so deal them round-robin
            else:                  # heaviest first, dealt in snake order: every rank gets the same load to within one point
                mine = deal_snake(costs, world)[rank]
        width = 2 + 9 * self._ncol()
        rows = (run_points(lambda **sp: self._build(sp), [pts[i] for i in mine], devices=devices).reshape(len(mine), width)
                if len(mine) else np.zeros((0, width)))      # a rank without points still enters the gather
        return gather_rows(mine, rows, N, width) if world > 1 else rows

    # -- one fast parameter: full computation for the first fast value, rescaled mpdi for the rest (scans.py:127-176)
    def _scan_fast(self, pts, rank, world, devices):
        fast = self._sScanp[self._sFast_id]
        nf = len(fast)
        n_slow = len(pts) // nf
        slow = np.arange(n_slow) if world == 1 else np.arange(rank, n_slow, world)
        ncol = self._ncol()
        width = 2 + 9 * ncol
        idx = (slow[:, None] * nf + np.arange(nf)[None, :]).reshape(-1).astype(np.int64)
        from acropolis_b200.models import DecayModel, AnnihilationModel
        # constructor parameters that only scale the source terms, for EXACTLY these classes (a user subclass, whose
        # constructor may do anything, takes the generic path)
        invariant = {DecayModel: ("n0a",), AnnihilationModel: ("a", "b")}.get(self._sModelClass, ())
        if self._sFast_id in invariant:
            Yf = self._fast_rows_invariant(pts, slow, fast, devices, ncol)
        else:
            Yf = self._fast_rows_generic(pts, slow, fast, devices, ncol)
        keys = sorted(pts[0]) if pts else []
        rows = np.empty((len(idx), width))
        for c, key in enumerate(keys):
            rows[:, c] = [pts[i][key] for i in idx]
        rows[:, 2:] = Yf.reshape(len(idx), 9, ncol).transpose(0, 2, 1).reshape(len(idx), 9 * ncol)
        return gather_rows(idx, rows, len(pts), width) if world > 1 else rows

    def _fast_rows_invariant(self, pts, slow, fast, devices, ncol):
        """Built-in model classes whose grids, thresholds and decay matrices do not depend on the fast parameter (it only
        scales the source terms: DecayModel n0a, AnnihilationModel a / b): ONE model per
        slow value instead of one per grid point (the reference builds all of them, scans.py:140-160; 40 000 constructions
        were 70 % of the wall time of the shipped example scans here), then one batched device call for all rescaled matrices."""
        nf = len(fast)
        Yf = np.zeros((len(slow), nf, 9, ncol))
        models = [self._build(pts[s * nf]) for s in slow]
        live = [r for r, m in enumerate(models) if not m.is_trivial()]
        for r, m in enumerate(models):
            if m.is_trivial():
                Yf[r, :] = m._squeeze_decays(m._sII.bbn_abundances())
        if live:
            _, matp = run_models([models[r] for r in live], devices=devices, want_matp=True)
            eng = Engine.get(models[live[0]]._sII, None if devices is None else list(devices)[0])
            mp = np.array([b[0] for b in matp])
            md = np.array([b[1] for b in matp])
            index = np.repeat(np.arange(len(live), dtype=np.int32), nf)
            factor = np.tile(fast / fast[0], len(live))
            tt = np.repeat(np.array([models[r]._t_at_tmax() for r in live]), nf)
            Y = eng.transfer(mp, md, factor, tt, index=index)
            report_status(eng.last_transfer_status, "acropolis_b200.scans.BufferedScanner.perform_scan")
            Yf[live] = Y.reshape(len(live), nf, 9, 3)[:, :, :, :ncol]
        return Yf

    def _fast_rows_generic(self, pts, slow, fast, devices, ncol):
        """Any model class: one model per grid point, as the reference does (scans.py:127-176)."""
        nf = len(fast)
        models = [[self._build(pts[s * nf + f]) for f in range(nf)] for s in slow]
        # the first fast value whose model is non-trivial supplies the buffer (reference: matpb stays None before)
        lead = [next((f for f in range(nf) if not row[f].is_trivial()), None) for row in models]
        lead_models = [row[f] for row, f in zip(models, lead) if f is not None]
        Yl, matp = run_models(lead_models, devices=devices, want_matp=True) if lead_models else (None, [])
        it = iter(matp)
        bufs = [next(it) if f is not None else None for f in lead]
        Yf = np.zeros((len(slow), nf, 9, ncol))
        jobs = []   # (row, f, buffer index, factor, t_at_tmax)
        for r, row in enumerate(models):
            for f, m in enumerate(row):
                if bufs[r] is None or f < lead[r] or m.is_trivial():
                    # before the buffer exists the reference runs the model itself: trivial by construction
                    Yf[r, f] = m._squeeze_decays(m._sII.bbn_abundances()) if m.is_trivial() else np.nan
                elif not uses_builtin_chain(m):      # user hooks between (mpdi, mdcy) and Yf: the model's own chain
                    m.set_matp_buffer(self._rescale_matp_buffer(bufs[r], fast[f] / fast[lead[r]]))
                    Yf[r, f] = m.run_disintegration()
                else:
                    m.set_matp_buffer(self._rescale_matp_buffer(bufs[r], fast[f] / fast[lead[r]]))
                    jobs.append((r, f, r, fast[f] / fast[lead[r]], m._t_at_tmax()))
        if jobs:
            eng = Engine.get(models[0][0]._sII, None if devices is None else list(devices)[0])
            zero = np.zeros((9, 9))
            mp = np.array([b[0] if b is not None else zero for b in bufs])
            md = np.array([b[1] if b is not None else zero for b in bufs])
            Y = eng.transfer(mp, md, np.array([j[3] for j in jobs]), np.array([j[4] for j in jobs]),
                             index=np.array([j[2] for j in jobs], dtype=np.int32))
            report_status(eng.last_transfer_status, "acropolis_b200.scans.BufferedScanner.perform_scan")
            for (r, f, *_), y in zip(jobs, Y):
                Yf[r, f] = y[:, :ncol]
        return Yf
