"""Batched engine: turns a list of model points into the ragged batch of include/acropolis_b200.h, owns the device
buffers (torch tensors) and calls the C ABI.  One ``Engine`` per (device, input tables).

This is the B200 replacement of the body of ``AbstractModel._pdi_matrix`` (reference models.py:112-130) for MANY
models at once; ``acropolis_b200.models`` / ``.scans`` / ``.cascade`` / ``.nucl`` are thin views onto it.
PyTorch is used for device memory, streams and pinned staging only."""
import ctypes as C
from math import log, log10

import numpy as np
import torch

from acropolis_b200 import _lib
import acropolis_b200.params as params
import acropolis_b200.flags as flags
from acropolis_b200.db import import_data_from_db


class PointSpec(object):
    """Inputs of one model point (everything ``NuclearReactor``/``MatrixGenerator`` read from the model)."""
    __slots__ = ("E0", "E_grid", "T", "S0", "sc_mode", "sc", "sc_E", "t_at_tmax")

    def __init__(self, E0, E_grid, T, S0, t_at_tmax, sc_mode=_lib.SC_NONE, sc=None, sc_E=None):
        self.E0 = float(E0)
        self.E_grid = np.ascontiguousarray(E_grid, dtype=np.float64)
        self.T = np.ascontiguousarray(T, dtype=np.float64)
        self.S0 = np.ascontiguousarray(S0, dtype=np.float64).reshape(len(self.T), 3)
        self.t_at_tmax = float(t_at_tmax)
        self.sc_mode = sc_mode
        self.sc = sc        # SEPARABLE: [NT,3]; DENSE: [NT,3,NE]
        self.sc_E = sc_E    # SEPARABLE: [NE,3]

    @classmethod
    def trusted(cls, E0, E_grid, T, S0, t_at_tmax, sc_mode=_lib.SC_NONE, sc=None, sc_E=None):
        """Constructor without conversions for callers that already hold contiguous float64 arrays of the right shapes
        (the vectorised ``point_specs`` of the models: one call per scan point)."""
        self = object.__new__(cls)
        self.E0, self.E_grid, self.T, self.S0, self.t_at_tmax = E0, E_grid, T, S0, t_at_tmax
        self.sc_mode, self.sc, self.sc_E = sc_mode, sc, sc_E
        return self


_GRID_MEMO = {}     # the points of a grid scan share their energy and temperature grids (read-only arrays)


def _memo_grid(key, make):
    g = _GRID_MEMO.get(key)
    if g is None:
        if len(_GRID_MEMO) >= 8192:
            _GRID_MEMO.clear()
        g = _GRID_MEMO[key] = make()
        g.setflags(write=False)
    return g


def energy_grid(E0):
    """NE and E_grid exactly as cascade.py:736-745 (same numpy/math expressions => bit-identical grid)."""
    def make():
        NE = max(int(log10(E0 / params.Emin) * params.NE_pd), params.NE_min)
        return np.logspace(log(params.Emin), log(E0), NE, base=np.e)
    return _memo_grid(("E", float(E0), params.Emin, params.NE_pd, params.NE_min), make)


def temperature_grid(Tmin, Tmax):
    """NT and Tr exactly as nucl.py:473-477."""
    def make():
        NT = int(log10(Tmax / Tmin) * params.NT_pd)
        return np.logspace(log10(Tmin), log10(Tmax), NT)
    return _memo_grid(("T", float(Tmin), float(Tmax), params.NT_pd), make)


class _Arena(object):
    """Packs numpy arrays into one pinned host buffer and one device buffer (a single H2D per chunk)."""

    def __init__(self, device):
        self.device = device
        self.host = None
        self.dev = None
        self.busy = None     # event after the last H2D from this arena's pinned buffer

    def upload(self, arrays, stream):
        offs, total = [], 0
        for a in arrays:
            offs.append(total)
            total += (a.nbytes + 255) & ~255
        total = max(total, 256)
        if self.host is None or self.host.numel() < total:
            cap = int(total * 1.25)
            self.host = torch.empty(cap, dtype=torch.uint8, pin_memory=True)
            self.dev = torch.empty(cap, dtype=torch.uint8, device=self.device)
        hv = self.host.numpy()
        for a, o in zip(arrays, offs):
            if a.nbytes:
                hv[o:o + a.nbytes] = a.reshape(-1).view(np.uint8)
        with torch.cuda.stream(stream):
            self.dev[:total].copy_(self.host[:total], non_blocking=True)
            self.busy = torch.cuda.Event()
            self.busy.record(stream)
        base = self.dev.data_ptr()
        return [base + o for o in offs], total


class Engine(object):
    _cache = {}

    @classmethod
    def get(cls, ii, device=None, variant="product", lane=0):
        """The engine of (device, input tables).  ``variant="xcheck"`` binds the test build of the library, which also has
        the per-(E,E',T) cross-check kernel modes (params.kernel_mode = PLAIN_GL / PER_T).  ``lane`` > 0 gives a sibling engine
        on the same device with its own library handle, stream, staging arenas and workspace: what two engines launch overlaps
        on the GPU (``scans.run_points`` alternates its groups between lanes 0 and 1)."""
        if device is None:
            device = torch.cuda.current_device() if torch.cuda.is_available() else 0
        device = torch.device("cuda", device) if not isinstance(device, torch.device) else device
        key = (device.index, id(ii), flags.usedb, variant, lane)
        eng = cls._cache.get(key)
        if eng is None:
            eng = cls._cache[key] = cls(ii, device, variant)
            if lane:
                eng._primary = cls.get(ii, device, variant)
                eng._primary._siblings.append(eng)
        return eng

    def __init__(self, ii, device, variant="product"):
        if not torch.cuda.is_available():
            raise RuntimeError("acropolis_b200: no CUDA device visible; the hot path has no CPU fallback")
        self.lib = _lib.load(variant)
        self.variant = variant
        self.ii = ii
        self.device = device
        self.stream = torch.cuda.Stream(device=device)
        rdb = import_data_from_db()
        self._use_db = rdb is not None
        if rdb is None:   # flags.usedb False or file absent: every rate comes from the direct integrals
            rdb = np.full((params.Tnum * params.Enum, 2), -200.0)
        lT, ldT = ii.cosmo_log_tables()
        Y0 = np.ascontiguousarray(ii.bbn_abundances(), dtype=np.float64)
        if Y0.shape[1] < 3:   # fewer than 3 abundance columns: pad with the mean column
            Y0 = np.ascontiguousarray(np.column_stack([Y0] + [Y0[:, :1]] * (3 - Y0.shape[1])))
        if ii.bbn_abundances().shape[1] > 3:
            raise NotImplementedError("acropolis_b200: abundance files with more than 3 columns (mean, high, low) are not "
                                      "supported by the device path (include/acropolis_b200.h: Y0 is float64[9,3])")
        self.ncol = ii.bbn_abundances().shape[1]
        t = _lib.Tables()
        t.ratedb = rdb.ctypes.data
        t.ratedb_nT, t.ratedb_nE = params.Tnum, params.Enum
        t.ratedb_Elog_min, t.ratedb_Elog_max = params.Emin_log, params.Emax_log
        t.ratedb_Tlog_min, t.ratedb_Tlog_max = params.Tmin_log, params.Tmax_log
        t.cosmo_log10_T, t.cosmo_log10_abs_dTdt, t.n_cosmo = lT.ctypes.data, ldT.ctypes.data, lT.shape[0]
        t.eta = float(ii.parameter("eta"))
        Y0c = np.ascontiguousarray(Y0[:, :3])      # must outlive acro_create: the struct holds its address
        t.Y0 = Y0c.ctypes.data
        self._keep = (rdb, lT, ldT, Y0, Y0c)
        h = C.c_void_p()
        rc = self.lib.acro_create(C.byref(h), device.index, C.byref(t), None)
        if rc != 0:
            raise RuntimeError("acro_create failed: %s" % self.lib.acro_last_error(None).decode())
        self.h = h
        self._arena = _Arena(device)
        self._ws = None
        self._params_key = None
        self._siblings = []          # engines of other lanes on this device (their launches count as this engine's)
        self._primary = None         # set on a sibling: the lane-0 engine of its device

    def __del__(self):
        try:
            if getattr(self, "h", None):
                self.lib.acro_destroy(self.h)
                self.h = None
        except Exception:
            pass

    # ------------------------------------------------------------------------------------------------------
    def _check(self, rc, what):
        if rc != 0:
            raise RuntimeError("%s failed (%d): %s" % (what, rc, self.lib.acro_last_error(self.h).decode()))

    def _sync_params(self):
        key = (params.quad_order, params.pdi_cell_order, params.Emin, params.approx_zero, params.Ephb_T_max,
               params.E_EC_max, bool(flags.usedb and self._use_db), int(params.kernel_mode), float(params.wien_from),
               bool(flags.universal))
        if key != self._params_key:
            p = _lib.Params(*key[:6], int(key[6]), key[7], key[8], int(key[9]), 0)
            self._check(self.lib.acro_set_params(self.h, C.byref(p)), "acro_set_params")
            self._params_key = key

    def launch_count(self):
        return int(self.lib.acro_launch_count(self.h)) + sum(e.launch_count() for e in self._siblings)

    def measure_fp64_peak(self):
        v = C.c_double()
        self._check(self.lib.acro_measure_fp64_peak(self.h, C.byref(v)), "acro_measure_fp64_peak")
        return v.value

    # ------------------------------------------------------------------------------------------------------
    @staticmethod
    def assemble(points):
        """Concatenate PointSpecs into the flat arrays of acro_batch_t (host side, numpy)."""
        P = len(points)
        ne = np.fromiter((len(p.E_grid) for p in points), dtype=np.int32, count=P)
        nt = np.fromiter((len(p.T) for p in points), dtype=np.int32, count=P)
        pt_e_off = np.zeros(P, dtype=np.int64)
        np.cumsum(ne[:-1], out=pt_e_off[1:])
        pt_sys_off = np.zeros(P, dtype=np.int32)
        np.cumsum(nt[:-1], out=pt_sys_off[1:])
        sys_point = np.repeat(np.arange(P, dtype=np.int32), nt)
        sys_ne = ne[sys_point].astype(np.int64)
        S = int(sys_point.shape[0])
        sys_f_off = np.zeros(S, dtype=np.int64)
        np.cumsum(sys_ne[:-1], out=sys_f_off[1:])
        # every system's packed triangle is padded to a multiple of 4 entries (32 bytes): the builder's 16-byte stores of
        # neighbouring pairs stay aligned and every 64-byte run it writes covers whole DRAM sectors
        sys_k_off = np.zeros(S, dtype=np.int64)
        tri = (sys_ne * (sys_ne + 1) // 2 + 3) & ~np.int64(3)
        np.cumsum(tri[:-1], out=sys_k_off[1:])
        ne64 = ne.astype(np.int64)
        tri_pt = (ne64 * (ne64 + 1) // 2 + 3) & ~np.int64(3)
        pt_pair_off = np.zeros(P, dtype=np.int64)
        np.cumsum(tri_pt[:-1], out=pt_pair_off[1:])
        modes = {p.sc_mode for p in points}
        if modes == {_lib.SC_NONE}:
            sc_mode, sc, sc_E = _lib.SC_NONE, np.zeros(0), np.zeros(0)
        elif modes <= {_lib.SC_NONE, _lib.SC_SEPARABLE}:
            sc_mode = _lib.SC_SEPARABLE
            sc = np.concatenate([p.sc if p.sc_mode else np.zeros((len(p.T), 3)) for p in points]).reshape(-1)
            sc_E = np.concatenate([p.sc_E if p.sc_mode else np.zeros((len(p.E_grid), 3)) for p in points]).reshape(-1)
        else:
            sc_mode, sc_E = _lib.SC_DENSE, np.zeros(0)
            parts = []
            for p in points:
                if p.sc_mode == _lib.SC_DENSE:
                    parts.append(np.asarray(p.sc, dtype=np.float64).reshape(-1))
                elif p.sc_mode == _lib.SC_SEPARABLE:
                    parts.append((p.sc[:, :, None] * p.sc_E.T[None, :, :]).reshape(-1))
                else:
                    parts.append(np.zeros(len(p.T) * 3 * len(p.E_grid)))
            sc = np.concatenate(parts)
        a = dict(pt_ne=ne, pt_nt=nt, pt_e_off=pt_e_off, pt_sys_off=pt_sys_off, pt_pair_off=pt_pair_off,
                 pt_e0=np.fromiter((p.E0 for p in points), dtype=np.float64, count=P),
                 pt_t_at_tmax=np.fromiter((p.t_at_tmax for p in points), dtype=np.float64, count=P),
                 e_grid=np.concatenate([p.E_grid for p in points]),
                 sys_point=sys_point, sys_T=np.concatenate([p.T for p in points]),
                 sys_f_off=sys_f_off, sys_k_off=sys_k_off,
                 s0=np.concatenate([p.S0 for p in points]).reshape(-1),
                 sc=np.ascontiguousarray(sc, dtype=np.float64), sc_E=np.ascontiguousarray(sc_E, dtype=np.float64))
        counts = dict(n_points=P, n_systems=S, n_energy_total=int(ne.sum()), n_sys_energy_total=int(sys_ne.sum()),
                      n_pairs_total=int(tri.sum()), n_pairs_points=int(tri_pt.sum()), ne_max=int(ne.max()), nt_max=int(nt.max()),
                      sc_mode=sc_mode)
        return a, counts

    _ORDER = ("pt_ne", "pt_nt", "pt_e_off", "pt_sys_off", "pt_pair_off", "pt_e0", "pt_t_at_tmax", "e_grid", "sys_point", "sys_T",
              "sys_f_off", "sys_k_off", "s0", "sc", "sc_E")

    @staticmethod
    def point_cost_bytes(ne, nt):
        """Workspace bytes one point needs (5 K planes + pair closed forms + rate work list + cross-section table of the
        pdi cells)."""
        tri = (ne * (ne + 1) // 2 + 3) & ~3
        return nt * (5 * 8 * tri + 32 * ne) + 24 * tri + 4 * nt + 512 + 16 * 17 * 8 * ne

    MAX_POINTS_PER_CHUNK = 32768     # the kernels index points with blockIdx.y (limit 65535)

    def chunks(self, points, budget=None):
        """Greedy split of the point list so that each chunk's workspace fits the budget."""
        if budget is None:
            budget = params.workspace_bytes
            if budget is None:      # default: 40% of the device memory that is free at first use, at most 64 GiB,
                if not hasattr(self, "_auto_budget"):      # shared between the lanes that exist on the device by then
                    free, _ = torch.cuda.mem_get_info(self.device)
                    n_lanes = 1 + len((self._primary or self)._siblings)
                    self._auto_budget = int(min(64 * 1024**3, 0.4 * free) / n_lanes)
                budget = self._auto_budget
        out, cur, acc = [], [], 0
        for i, p in enumerate(points):
            c = self.point_cost_bytes(len(p.E_grid), len(p.T))
            if cur and (acc + c > budget or len(cur) >= self.MAX_POINTS_PER_CHUNK):
                out.append(cur)
                cur, acc = [], 0
            cur.append(i)
            acc += c
        if cur:
            out.append(cur)
        return out

    def run(self, points, stages=_lib.STAGE_ALL, want=("Yf", "mpdi", "mdcy", "status"), budget=None, sync=True):
        """Run the hot path for a list of PointSpecs.  Returns a dict of numpy arrays for the keys in `want`
        (any of G, F, pdi, mpdi, mdcy, Yf, status), concatenated over points in input order."""
        self._sync_params()
        res = {k: [] for k in want}
        pending = []
        self.last_h2d_bytes = self.last_d2h_bytes = 0
        for idx in self.chunks(points, budget):
            ch = self.stage([points[i] for i in idx], want)
            self.launch(ch, stages)
            pending.append(self.fetch(ch, want))
            if len(pending) > 1:                      # keep one chunk in flight while the next is assembled
                self._collect(pending.pop(0), res)
        for pnd in pending:
            self._collect(pnd, res)
        return {k: (np.concatenate(v) if v else np.zeros(0)) for k, v in res.items()}

    def submit(self, points, stages=_lib.STAGE_ALL, want=("Yf", "status"), budget=None):
        """Stage and launch `points` (split by the workspace budget) without waiting; returns pending handles for
        ``collect``.  Lets the caller prepare the next group of points on the host while the GPU works."""
        self._sync_params()
        pend = []
        for idx in self.chunks(points, budget):
            ch = self.stage([points[i] for i in idx], want)
            self.launch(ch, stages)
            pend.append(self.fetch(ch, want))
        return pend

    def collect(self, pending, want=("Yf", "status")):
        """Wait for pending handles of ``submit`` and return the concatenated outputs."""
        res = {k: [] for k in want}
        for pnd in pending:
            self._collect(pnd, res)
        return {k: (np.concatenate(v) if v else np.zeros(0)) for k, v in res.items()}

    def stage(self, pts, want=("Yf", "status"), resident=False):
        """Assemble a chunk on the host, copy it to the device (one pinned H2D) and allocate its outputs.  With
        ``resident`` the chunk gets its own device buffers (it can be launched repeatedly, e.g. by the benchmark)."""
        # heaviest points first: blocks are scheduled in index order, and a launch that ends with the longest blocks of the
        # largest points has the longest tail (measured: 1.1 ms of an 80 ms slice).  `perm[k]` = input index of sorted point k;
        # _collect() puts the outputs back into input order.
        cost = np.fromiter((len(p.T) * len(p.E_grid) ** 2 for p in pts), dtype=np.int64, count=len(pts))
        perm = np.argsort(-cost, kind="stable")
        if np.array_equal(perm, np.arange(len(pts))):
            perm = None
        else:
            pts = [pts[i] for i in perm]
        a, cnt = self.assemble(pts)
        st = self.stream
        arena = _Arena(self.device) if resident else self._arena_for_chunk()
        if not resident and arena.busy is not None:
            arena.busy.synchronize()      # the copy that last used this staging buffer must have finished
        ptrs, nbytes = arena.upload([a[k] for k in self._ORDER], st)
        self.last_h2d_bytes = getattr(self, "last_h2d_bytes", 0) + nbytes
        b = _lib.Batch()
        for k, v in cnt.items():
            setattr(b, k, v)
        for k, ptr in zip(self._ORDER, ptrs):
            setattr(b, k, ptr if a[k].nbytes else None)
        S, P, NF = cnt["n_systems"], cnt["n_points"], cnt["n_sys_energy_total"]
        dev = self.device
        with torch.cuda.stream(st):
            outs = dict(G=torch.empty(NF * 3, dtype=torch.float64, device=dev),
                        pdi=torch.empty(S * _lib.NR, dtype=torch.float64, device=dev),
                        mpdi=torch.empty(P * 81, dtype=torch.float64, device=dev),
                        mdcy=torch.empty(P * 81, dtype=torch.float64, device=dev),
                        Yf=torch.empty(P * 27, dtype=torch.float64, device=dev),
                        status=torch.zeros(P, dtype=torch.int32, device=dev))
            if "F" in want:
                outs["F"] = torch.empty(NF * 3, dtype=torch.float64, device=dev)
            need = int(self.lib.acro_workspace_bytes(self.h, C.byref(b)))
            if self._ws is None or self._ws.numel() < need:
                # growing means cudaMalloc, and that waits for everything in flight (measured: 14-37 ms in the middle of a
                # pipelined scan whose later points are larger): grow with 50 % headroom so that it happens once or twice
                free, _ = torch.cuda.mem_get_info(dev)
                have = self._ws.numel() if self._ws is not None else 0
                self._ws = None
                self._ws = torch.empty(int(min(max(need, 1.5 * need), max(need, 0.8 * (free + have)))), dtype=torch.uint8, device=dev)
        o = _lib.Out()
        for k, t in outs.items():
            setattr(o, k, t.data_ptr())
        return dict(cnt=cnt, batch=b, out=o, outs=outs, keep=(a, arena), perm=perm, ne=a["pt_ne"], nt=a["pt_nt"])

    def launch(self, ch, stages=_lib.STAGE_ALL):
        """Enqueue the stages of a staged chunk on the engine's stream (asynchronous)."""
        rc = self.lib.acro_run_batch(self.h, C.byref(ch["batch"]), C.byref(ch["out"]), self._ws.data_ptr(), self._ws.numel(),
                                     stages, self.stream.cuda_stream)
        self._check(rc, "acro_run_batch")

    def fetch(self, ch, want):
        """Enqueue the D2H copies of the wanted outputs into pinned memory; returns a pending handle."""
        host = {}
        with torch.cuda.stream(self.stream):
            for k in want:
                host[k] = torch.empty(ch["outs"][k].shape, dtype=ch["outs"][k].dtype, pin_memory=True)
                host[k].copy_(ch["outs"][k], non_blocking=True)
                self.last_d2h_bytes = getattr(self, "last_d2h_bytes", 0) + host[k].numel() * host[k].element_size()
            ev = torch.cuda.Event()
            ev.record(self.stream)
        return dict(host=host, ev=ev, cnt=ch["cnt"], keep=ch, batch=ch["batch"], perm=ch["perm"], ne=ch["ne"], nt=ch["nt"])

    def count_work(self, ch):
        """(pairs, N_a9, N_a10, N_a11) of a staged chunk (SURVEY.md 8d work census)."""
        out = (C.c_int64 * 4)()
        self._check(self.lib.acro_count_work(self.h, C.byref(ch["batch"]), C.byref(out), self.stream.cuda_stream), "acro_count_work")
        return [int(v) for v in out]

    def stage_ms(self):
        ms = (C.c_float * 4)()
        self._check(self.lib.acro_last_stage_ms(self.h, C.byref(ms)), "acro_last_stage_ms")
        return np.array(list(ms), dtype=np.float64)

    def kernel_split_ms(self):
        """Device time [ms] of the kernel builder's launches in the last acro_run_batch, in the order of
        ``kernel_split_names()``."""
        ms = (C.c_float * len(_lib.KERNEL_SPLIT_NAMES))()
        self._check(self.lib.acro_last_kernel_split_ms(self.h, C.byref(ms)), "acro_last_kernel_split_ms")
        return np.array(list(ms), dtype=np.float64)

    @staticmethod
    def kernel_split_names():
        return _lib.KERNEL_SPLIT_NAMES

    @staticmethod
    def n_kernel_splits():
        return len(_lib.KERNEL_SPLIT_NAMES)

    def _arena_for_chunk(self):
        # two arenas alternate so that chunk n+1 can be staged while chunk n is still being read by the GPU
        if not hasattr(self, "_arenas"):
            self._arenas, self._arena_i = [_Arena(self.device) for _ in range(4)], 0
        self._arena_i = (self._arena_i + 1) % len(self._arenas)
        return self._arenas[self._arena_i]

    def _collect(self, pnd, res):
        pnd["ev"].synchronize()     # this chunk's own event only: later chunks keep the GPU busy meanwhile
        cnt = pnd["cnt"]
        shapes = dict(Yf=(cnt["n_points"], 9, 3), mpdi=(cnt["n_points"], 9, 9), mdcy=(cnt["n_points"], 9, 9),
                      status=(cnt["n_points"],), pdi=(cnt["n_systems"], _lib.NR), G=(-1,), F=(-1,))
        for k, t in pnd["host"].items():
            res[k].append(self.unpermute(k, t.numpy().reshape(shapes[k]), pnd["perm"], pnd["ne"], pnd["nt"]))

    @staticmethod
    def unpermute(key, arr, perm, ne, nt):
        """Output `key` of a chunk that stage() sorted by cost, back in the caller's point order (a copy in any case)."""
        if perm is None:
            return arr.copy()
        if key in ("Yf", "mpdi", "mdcy", "status"):          # one row per point
            out = np.empty_like(arr)
            out[perm] = arr
            return out
        seg = (nt if key == "pdi" else 3 * ne.astype(np.int64) * nt).astype(np.int64)     # rows / values per sorted point
        start = np.concatenate(([0], np.cumsum(seg)))
        inv = np.argsort(perm)                                # sorted position of input point i
        return np.concatenate([arr[start[k]:start[k + 1]] for k in inv]) if len(inv) else arr.copy()

    # ------------------------------------------------------------------------------------------------------
    def kernels_of(self, point, it):
        """Dense K[5,NE,NE] (planes gg, eg, ge-, ge+, ee) of `point` at temperature index `it` (parity checks)."""
        self._sync_params()
        sub = PointSpec(point.E0, point.E_grid, point.T[it:it + 1], point.S0[it:it + 1], point.t_at_tmax)
        ch = self.stage([sub], ("G",))
        self.launch(ch, _lib.STAGE_RATES | _lib.STAGE_KERNELS)
        pnd = self.fetch(ch, ("G",))
        ne = len(point.E_grid)
        K = torch.empty(5 * ne * ne, dtype=torch.float64, device=self.device)
        with torch.cuda.stream(self.stream):
            self._check(self.lib.acro_unpack_kernels(self.h, C.byref(pnd["batch"]), self._ws.data_ptr(), 0, K.data_ptr(),
                                                     self.stream.cuda_stream), "acro_unpack_kernels")
        self.stream.synchronize()
        return K.cpu().numpy().reshape(5, ne, ne), pnd["host"]["G"].numpy().reshape(3, ne).copy()

    def total_rates(self, E, T):
        """Gamma_X(E,T) for arrays of (E,T) -> [n,3] (SpectrumGenerator._rate_x, cascade.py:721-730)."""
        self._sync_params()
        E = np.ascontiguousarray(np.broadcast_arrays(E, T)[0], dtype=np.float64).reshape(-1)
        T = np.ascontiguousarray(np.broadcast_to(T, E.shape), dtype=np.float64).reshape(-1)
        n = E.shape[0]
        dE, dT = torch.tensor(E, device=self.device), torch.tensor(T, device=self.device)
        out = torch.empty(n * 3, dtype=torch.float64, device=self.device)
        torch.cuda.current_stream(self.device).synchronize()
        self._check(self.lib.acro_total_rates(self.h, n, dE.data_ptr(), dT.data_ptr(), out.data_ptr(),
                                              self.stream.cuda_stream), "acro_total_rates")
        return out.cpu().numpy().reshape(n, 3)

    def direct_rates(self, E, T):
        """The two tabulated rates as converged direct integrals for arrays of (E,T) -> [n,2] (pair creation, inverse
        Compton): what rates.db stores before log10 (tools/make_rates_db.py)."""
        self._sync_params()
        E = np.ascontiguousarray(np.broadcast_arrays(E, T)[0], dtype=np.float64).reshape(-1)
        T = np.ascontiguousarray(np.broadcast_to(T, E.shape), dtype=np.float64).reshape(-1)
        n = E.shape[0]
        dE, dT = torch.tensor(E, device=self.device), torch.tensor(T, device=self.device)
        out = torch.empty(n * 3, dtype=torch.float64, device=self.device)
        torch.cuda.current_stream(self.device).synchronize()
        self._check(self.lib.acro_direct_rates(self.h, n, dE.data_ptr(), dT.data_ptr(), out.data_ptr(), self.stream.cuda_stream),
                    "acro_direct_rates")
        return out.cpu().numpy().reshape(n, 3)[:, :2]

    def matp(self, T_list, pdi_list, T_lo=None, t_at_tmax=None, want_Yf=False):
        """MatrixGenerator.get_matp for several points: T_list[p] (NT_p,), pdi_list[p] (NT_p,17)."""
        self._sync_params()
        P = len(T_list)
        nt = np.array([len(t) for t in T_list], dtype=np.int32)
        off = np.zeros(P, dtype=np.int32)
        np.cumsum(nt[:-1], out=off[1:])
        dev = self.device
        d_nt, d_off = torch.from_numpy(nt).to(dev), torch.from_numpy(off).to(dev)
        d_T = torch.from_numpy(np.concatenate(T_list).astype(np.float64)).to(dev)
        d_pdi = torch.from_numpy(np.concatenate([np.asarray(p, dtype=np.float64).reshape(-1, _lib.NR) for p in pdi_list])
                                 .reshape(-1)).to(dev)
        d_lo = torch.from_numpy(np.array(T_lo, dtype=np.float64)).to(dev) if T_lo is not None else None
        d_tt = torch.from_numpy(np.asarray(t_at_tmax if t_at_tmax is not None else np.zeros(P), dtype=np.float64)).to(dev)
        mpdi = torch.empty(P * 81, dtype=torch.float64, device=dev)
        mdcy = torch.empty(P * 81, dtype=torch.float64, device=dev)
        Yf = torch.empty(P * 27, dtype=torch.float64, device=dev)
        status = torch.zeros(P, dtype=torch.int32, device=dev)
        torch.cuda.current_stream(dev).synchronize()
        rc = self.lib.acro_matp_batch(self.h, P, int(nt.max()), d_nt.data_ptr(), d_off.data_ptr(), d_T.data_ptr(),
                                      d_pdi.data_ptr(), d_lo.data_ptr() if d_lo is not None else None, d_tt.data_ptr(),
                                      mpdi.data_ptr(), mdcy.data_ptr(), Yf.data_ptr(), status.data_ptr(),
                                      self.stream.cuda_stream)
        self._check(rc, "acro_matp_batch")
        self.stream.synchronize()
        out = (mpdi.cpu().numpy().reshape(P, 9, 9), mdcy.cpu().numpy().reshape(P, 9, 9))
        if want_Yf:
            return out + (Yf.cpu().numpy().reshape(P, 9, 3), status.cpu().numpy())
        return out

    def transfer(self, mpdi, mdcy, factor, t_at_tmax, want_expm=False, index=None):
        """Fast-parameter path: Yf[n,9,3] = postd . expm(factor*mpdi + mdcy) . pred . Y0 (scans.py:106-107).  With ``index``
        (int[n]) job i uses the matrix pair index[i] of mpdi/mdcy [m,9,9]: the rescalings of one buffered pair share it."""
        dev = self.device
        mpdi = np.ascontiguousarray(mpdi, dtype=np.float64).reshape(-1, 81)
        mdcy = np.ascontiguousarray(mdcy, dtype=np.float64).reshape(-1, 81)
        n = mpdi.shape[0] if index is None else len(index)
        d_mp = torch.from_numpy(mpdi.reshape(-1)).to(dev)
        d_md = torch.from_numpy(mdcy.reshape(-1)).to(dev)
        d_ix = torch.from_numpy(np.array(index, dtype=np.int32)).to(dev) if index is not None else None
        d_f = torch.from_numpy(np.array(np.broadcast_to(factor, (n,)), dtype=np.float64)).to(dev)
        d_t = torch.from_numpy(np.array(np.broadcast_to(t_at_tmax, (n,)), dtype=np.float64)).to(dev)
        Yf = torch.empty(n * 27, dtype=torch.float64, device=dev)
        ex = torch.empty(n * 81, dtype=torch.float64, device=dev) if want_expm else None
        status = torch.zeros(n, dtype=torch.int32, device=dev)
        torch.cuda.current_stream(dev).synchronize()
        rc = self.lib.acro_transfer_batch(self.h, n, d_mp.data_ptr(), d_md.data_ptr(), d_ix.data_ptr() if d_ix is not None else None,
                                          d_f.data_ptr(), d_t.data_ptr(), ex.data_ptr() if ex is not None else None, Yf.data_ptr(),
                                          status.data_ptr(), self.stream.cuda_stream)
        self._check(rc, "acro_transfer_batch")
        self.stream.synchronize()
        self.last_transfer_status = status.cpu().numpy()
        if want_expm:
            return Yf.cpu().numpy().reshape(n, 9, 3), ex.cpu().numpy().reshape(n, 9, 9)
        return Yf.cpu().numpy().reshape(n, 9, 3)
