#!/usr/bin/env python3
"""Benchmark of the hot path: scan points/s for the 2-D DecayModel scan (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W          # our arm (N>1: launched by torchrun, one rank per GPU)
    python bench.py --impl reference --gpus N --steps K --warmup W   # CPU arm: the oracle port on all host threads
    python bench.py --config {cfg1,cfg2,cfg3,cfg3b,cfg3scan,cfg4,cfg5} ...   # the other BASELINE.json configs (default cfg3)

Default workload (BASELINE.json configs[2], SURVEY.md 8d "M3"): BufferedScanner(DecayModel, mphi=ScanParameter(0,3,200),
tau=ScanParameter(3,10,200), temp0=10, n0a=1e-10, bree=0, braa=1) -- 40 000 points, all full computations.  The grid
is flattened in the reference's order (mphi outer, tau inner) and cut into 40 interleaved slices L[k::40] of 1000
points (200 mphi x 5 tau each, 160 of them below threshold); the 40 slices together ARE the full scan and every
slice has the scan's cost mix.  One step = one slice per GPU (weak scaling: step t on rank r takes slice
(t*N + r) mod 40).  `value` is device-timed with the slices resident in HBM; `e2e` goes through the public
``acropolis_b200.scans.run_points`` call (model construction, source arrays, pinned H2D, kernels, D2H) per step.

CPU arm (`cpu_baseline`, `--impl reference`): oracle/acro_oracle.c (kind "port": the reference is Python and cannot travel
to the GPU box; profiles/r02_port_vs_reference.json holds the port-to-reference factor measured in the build container)
on a bounded sample of the SAME slices: every 27th point of the slice.  A slice's tau index has period 5 along the
slice, 27 is coprime to it, so the sample cycles through all five tau values (10^3 ... 10^10 s) and steps through the mass
axis; the line prints the sample's share of points, triples and of the three thermal-integral censuses next to the
slice's so that its representativeness can be checked.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC, UNIT = "scan points/sec (DecayModel 2-D scan)", "points/s"
CPU_STRIDE = 27      # coprime to the tau period (5) of a slice; see the module docstring

# ---------------------------------------------------------------------------------------------------------------
# workloads (BASELINE.json configs)
# ---------------------------------------------------------------------------------------------------------------
WORKLOADS = {
    "cfg3": dict(kind="slices", model="DecayModel", n_grid=200, n_slices=40, NE_pd=None, NT_pd=None,
                 fixed=dict(temp0=10., n0a=1e-10, bree=0., braa=1.),
                 text="cfg3 DecayModel mphi x tau 200x200 scan (examples/scan_decay_* ranges, no fast parameter); "
                      "step = one of 40 interleaved slices = 1000 scan points per GPU (160 below threshold)"),
    "cfg5": dict(kind="slices", model="DecayModel", n_grid=50, n_slices=10, NE_pd=480, NT_pd=80,
                 fixed=dict(temp0=10., n0a=1e-10, bree=0., braa=1.),
                 text="cfg5 high-resolution sweep: DecayModel mphi x tau 50x50 scan at NE_pd=480, NT_pd=80 (4x the defaults); "
                      "step = one of 10 interleaved slices = 250 scan points per GPU"),
    "cfg1": dict(kind="point", model="DecayModel", args=(10., 1e5, 10., 1e-10, 0., 1.),
                 text="cfg1 single point DecayModel(10, 1e5, 10, 1e-10, 0, 1) (./decay 10 1e5 10 1e-10 0 1): NE=62, NT=40; "
                      "step = one run_disintegration()"),
    "cfg2": dict(kind="point", model="AnnihilationModel", args=(10., 1e-25, 0., 0., 1., 0.),
                 text="cfg2 single point AnnihilationModel(10, 1e-25, 0, 0, 1, 0): NE=98, NT=80; step = one run_disintegration()"),
    "cfg3b": dict(kind="fastscan", which="scan_decay_tau_aa",
                  text="cfg3b examples/scan_decay_tau_aa verbatim: DecayModel mphi(200) x n0a(200, fast) at tau=1e7 s: 200 full "
                       "points + 39 800 rescale-only points; step = the whole scan through BufferedScanner.perform_scan"),
    "cfg3scan": dict(kind="fastscan", which="scan_decay_full",
                     text="cfg3 as ONE call: BufferedScanner(DecayModel, mphi=ScanParameter(0,3,200), tau=ScanParameter(3,10,200), temp0=10, "
                          "n0a=1e-10, bree=0, braa=1).perform_scan() -- 40 000 full points; under torchrun the points are dealt "
                          "cost-balanced to the ranks and the rows all-gathered (strong scaling); step = the whole scan"),
    "cfg4": dict(kind="fastscan", which="scan_pwave_ee",
                 text="cfg4 examples/scan_pwave_ee verbatim: AnnihilationModel mchi(200) x b(200, fast), a=0, tempkd=1, bree=1: 200 "
                      "full + 39 800 rescale-only points, plus BenchmarkModel1-3 (ResonanceModel) points; step = the whole scan"),
}


def workload_setup(cfg):
    """Patch the resolution parameters the workload asks for (cfg5) before any grid is built."""
    import acropolis_b200.params as prm
    w = WORKLOADS[cfg]
    if w.get("NE_pd"):
        prm.NE_pd = w["NE_pd"]
    if w.get("NT_pd"):
        prm.NT_pd = w["NT_pd"]
    return w


def grid_points(w):
    n = w["n_grid"]
    mphi, tau = np.logspace(0, 3, n), np.logspace(3, 10, n)
    return [dict(mphi=float(m), tau=float(t)) for m in mphi for t in tau]   # scans.py:193-197 order


def slice_params(w, pts, k):
    return pts[(k % w["n_slices"])::w["n_slices"]]


def model_class(name):
    import acropolis_b200.models as M
    return getattr(M, name)


def build_models(w, params):
    cls = model_class(w["model"])
    return [cls(**p, **w["fixed"]) for p in params]


def cpu_reference_sample(w, pts, k, stride=CPU_STRIDE):
    """Bounded sample of slice k for the CPU arm: every `stride`-th point of the slice (same definition in both arms)."""
    return slice_params(w, pts, k)[::stride]


def sample_text(w, stride=CPU_STRIDE):
    n = len(range(0, w["n_grid"] * w["n_grid"] // w["n_slices"], stride))
    return "every %d-th point of one %d-point slice per step (%d points/step; stride coprime to the slice's tau period)" % (
        stride, w["n_grid"] * w["n_grid"] // w["n_slices"], n)


def workload_config(cfg, extra=None):
    import acropolis_b200.params as prm
    w = WORKLOADS[cfg]
    c = {"workload": w["text"], "name": cfg, "NE_pd": prm.NE_pd, "NT_pd": prm.NT_pd, "quad_order": prm.quad_order,
         "wien_from": prm.wien_from}
    if w["kind"] == "slices":
        c.update({"grid": [w["n_grid"], w["n_grid"]], "points_per_step_per_gpu": w["n_grid"] * w["n_grid"] // w["n_slices"],
                  "cpu_sample": sample_text(w),
                  "l2": "inputs larger than L2: each step streams the step's kernel planes (GBs) through HBM, no reuse between steps"})
    else:
        c["l2"] = "L2 flushed between timed iterations (256 MiB device memset)"
    if extra:
        c.update(extra)
    return c


# ---------------------------------------------------------------------------------------------------------------
# host-side work census of a list of points (numpy; the same limit tests as count_work_kernel / cascade.py:416-427,
# 516-534, 572-592): (pairs, N_a9, N_a10, N_a11) summed over the temperatures of every point
# ---------------------------------------------------------------------------------------------------------------
def host_census(specs, Ephb_T_max=200.0):
    me, me2 = 0.511, 0.511 ** 2
    tot = np.zeros(4)
    for s in specs:
        E = np.asarray(s.E_grid)
        T = np.asarray(s.T)[None, :]
        i, j = np.triu_indices(len(E))
        Ei, Ep = E[i][:, None], E[j][:, None]
        xmax = Ephb_T_max * T
        ub = Ep - me2 / (4.0 * Ep)
        with np.errstate(divide="ignore", invalid="ignore"):
            llim9 = 0.25 * me2 * Ei / (Ep * Ep - Ep * Ei)
            c9 = (~(Ep < Ei + me)) & (np.minimum(np.minimum(Ei, ub), xmax) > llim9)
            pf, qf = 0.25 * me2 / Ep - Ei, 0.25 * me2 * (Ep - Ei) / Ep
            sq = np.sqrt((0.5 * pf) ** 2 - qf)
            c10 = (Ei != Ep) & (np.minimum(np.minimum(-0.5 * pf + sq, ub), xmax) > (-0.5 * pf - sq))
            dE, rt, den = Ep - Ei, np.sqrt(Ei * Ei - me2), 4.0 * Ep * (Ep - Ei) + me2
            z1, z2 = Ep * (me2 - 2.0 * dE * (rt - Ei)) / den, Ep * (me2 + 2.0 * dE * (rt + Ei)) / den
            c11 = (~(Ep < me2 / (50.0 * T))) & (np.minimum(np.minimum(z2, Ep), xmax) > np.maximum(me2 / Ep, z1))
        tot += np.array([len(i) * T.shape[1], np.count_nonzero(c9), np.count_nonzero(c10), np.count_nonzero(c11)], dtype=np.float64)
    return tot


def census_fractions(c):
    p = max(float(c[0]), 1.0)
    return {"pairs": float(c[0]), "N_a9/pairs": float(c[1]) / p, "N_a10/pairs": float(c[2]) / p, "N_a11/pairs": float(c[3]) / p}


# ---------------------------------------------------------------------------------------------------------------
class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,timestamp")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=lambda: [self.lines.append(l) for l in self.proc.stdout], daemon=True).start()
        except OSError:
            self.proc = None

    def stop(self, window=None):
        """Summary of the samples whose time stamp lies in `window` = (t0, t1) [time.time() values]: the GPU is under load
        from the first warm-up launch to the end of the timed region.  nvidia-smi needs ~0.5 s to deliver its first sample,
        so the sampler is started well before the window and the samples are selected by their own time stamps."""
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)                      # one more sampling period: a sample taken at the end of the window arrives
        self.proc.terminate()
        time.sleep(0.05)
        import datetime
        sm, mx, reasons, n_all = [], [], set(), 0
        for l in list(self.lines):
            f = [x.strip() for x in l.split(",")]
            if len(f) < 9:
                continue
            n_all += 1
            if window is not None:
                try:
                    ts = datetime.datetime.strptime(f[8], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                except ValueError:
                    continue
                if not (window[0] - 0.05 <= ts <= window[1] + 0.1):
                    continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "samples_total": n_all, "reasons": sorted(reasons),
                "window": "first warm-up launch .. end of the timed region (GPU busy throughout)"}


def dist_setup(n_gpus):
    import torch
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        # NCCL prints its version banner to STDOUT when the communicator is created; the contract is one JSON line on
        # stdout, so file descriptor 1 points at stderr until the first collective has run.
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    else:
        torch.cuda.set_device(0)
    return rank, world, local, dist


def barrier(dist):
    import torch
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()


def _reduce(dist, v, op):
    import torch
    if dist is None:
        return v
    t = torch.tensor([v], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=getattr(dist.ReduceOp, op))
    return float(t.item())


def max_over_ranks(dist, v):
    return _reduce(dist, v, "MAX")


def sum_over_ranks(dist, v):
    return _reduce(dist, v, "SUM")


# ---------------------------------------------------------------------------------------------------------------
# CPU arm
# ---------------------------------------------------------------------------------------------------------------
def run_oracle_models(models, threads=0, eps=1e-3):
    """The oracle port (oracle/acro_oracle.c) on all host threads over a list of models -> (seconds, threads, Yf of the
    non-trivial models, triples, n models, specs).  This is the CPU restatement of scans.py's Pool over
    run_disintegration(); see oracle/ header."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    from pyoracle import Oracle
    from acropolis_b200.engine import Engine
    work = [m for m in models if not m.is_trivial()]
    specs = [m.point_spec() for m in work]
    if not specs:
        return 0.0, 1, np.zeros((0, 9, 3)), 0.0, len(models), specs
    a, cnt = Engine.assemble(specs)
    o = Oracle(models[0]._sII, eps=eps)
    t0 = time.perf_counter()
    Yf, _ = o.run_points(a, cnt, n_threads=threads)
    dt = time.perf_counter() - t0
    used = min(threads if threads > 0 else (os.cpu_count() or 1), len(specs))
    triples = float(sum(len(s.T) * len(s.E_grid) * (len(s.E_grid) + 1) / 2 for s in specs))
    return dt, used, Yf, triples, len(models), specs


def run_oracle_scan(params, threads=0, w=None):
    """Compatibility form used by tools/port_vs_reference.py: list of DecayModel parameter dicts of the cfg3 scan."""
    w = WORKLOADS["cfg3"] if w is None else w
    dt, used, Yf, triples, n, _ = run_oracle_models(build_models(w, params), threads)
    return dt, used, Yf, triples, n


def port_vs_reference_note():
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "r02_port_vs_reference.json")))
        return {"port_over_reference": d["port_over_reference"], "reference_points_per_s": d["reference"]["points_per_s"],
                "port_points_per_s": d["port"]["points_per_s"], "cores": d["reference"]["cores"], "grid": d["grid"],
                "source": "profiles/r02_port_vs_reference.json (build container, same sub-grid, both at eps=1e-3)"}
    except Exception:
        return None


def main_reference(args):
    import acropolis_b200.flags as flags
    flags.verbose = False
    if int(os.environ.get("RANK", 0)) != 0:
        return 0
    cfg = args.config
    w = workload_setup(cfg)
    tot_t, tot_p, cores = 0.0, 0, 1
    census = np.zeros(4)
    if w["kind"] == "slices":
        pts = grid_points(w)
        for step in range(args.warmup + args.steps):
            dt, cores, _, _, n, specs = run_oracle_models(build_models(w, cpu_reference_sample(w, pts, step)))
            if step >= args.warmup:
                tot_t += dt
                tot_p += n
                census += host_census(specs)
        sample = sample_text(w) + ", oracle/acro_oracle.c on %d threads" % cores
    elif w["kind"] == "point":
        for step in range(args.warmup + args.steps):
            m = model_class(w["model"])(*w["args"])
            dt, cores, _, _, n, specs = run_oracle_models([m], threads=1)
            if step >= args.warmup:
                tot_t += dt
                tot_p += n
        sample = "the point itself, oracle/acro_oracle.c on 1 thread (a single point has no parallel loop in the reference)"
    else:
        print(json.dumps({"impl": "reference", "unavailable": "no CPU arm for the fast-parameter scans (--config %s): "
                          "200 full points of the reference take hours on the host; see cfg3" % cfg}))
        return 0
    value = tot_p / tot_t
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / max(args.steps, 1), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(cfg),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                             "sample_census": census_fractions(census), "port_vs_reference": port_vs_reference_note()},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))
    return 0


# ---------------------------------------------------------------------------------------------------------------
# roofline inputs from the committed ncu summaries (tools/ncu_flops_summary.py writes profiles/r02_fp64_flops.json)
# ---------------------------------------------------------------------------------------------------------------
def load_flop_profile():
    path = os.path.join(ROOT, "profiles", "r02_fp64_flops.json")
    try:
        return json.load(open(path)), "profiles/r02_fp64_flops.json"
    except Exception:
        return None, None


def library_source_hash():
    import hashlib
    h = hashlib.sha256()
    d = os.path.join(ROOT, "acropolis_b200", "csrc")
    for f in sorted(os.listdir(d)):
        h.update(open(os.path.join(d, f), "rb").read())
    return h.hexdigest()[:16]


def flush_l2(torch, dev, buf=[None]):
    if buf[0] is None:
        buf[0] = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    buf[0].zero_()


# ---------------------------------------------------------------------------------------------------------------
# GPU arms
# ---------------------------------------------------------------------------------------------------------------
def main_gpu(args):
    import acropolis_b200.flags as flags
    flags.verbose = False
    os.environ["ACROPOLIS_B200_DIST"] = "0"      # ranks run independent units; no gather inside run_points
    w = workload_setup(args.config)
    if w["kind"] == "slices":
        return gpu_slices(args, w)
    if w["kind"] == "point":
        return gpu_point(args, w)
    return gpu_fastscan(args, w)


def gpu_slices(args, w):
    import torch
    from acropolis_b200.engine import Engine
    from acropolis_b200.scans import run_points
    from functools import partial
    import acropolis_b200.params as prm
    cfg = args.config
    rank, world, local, dist = dist_setup(args.gpus)
    pts = grid_points(w)
    W, K = args.warmup, args.steps
    per_step = w["n_grid"] * w["n_grid"] // w["n_slices"]
    model = partial(model_class(w["model"]), **w["fixed"])
    ii = build_models(w, pts[:1])[0]._sII
    eng = Engine.get(ii, local)
    launches0 = eng.launch_count()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()

    # ---- device-timed arm: slices resident in HBM --------------------------------------------------------------
    staged = []
    for step in range(W + K):
        models = build_models(w, slice_params(w, pts, step * world + rank))
        specs = [m.point_spec() for m in models if not m.is_trivial()]
        staged.append([eng.stage([specs[i] for i in idx], ("Yf", "status"), resident=True) for idx in eng.chunks(specs)])
    eng.stream.synchronize()
    census = np.zeros(4)
    for chs in staged[W:]:
        for ch in chs:
            census += np.array(eng.count_work(ch), dtype=np.float64)
    barrier(dist)
    t_load0 = time.time()
    for chs in staged[:W]:
        for ch in chs:
            eng.launch(ch)
    eng.stream.synchronize()
    barrier(dist)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = eng.launch_count()
    ev0.record(eng.stream)
    for chs in staged[W:]:
        for ch in chs:
            eng.launch(ch)
    ev1.record(eng.stream)
    ev1.synchronize()
    barrier(dist)
    t_load1 = time.time()
    dev_launches = eng.launch_count() - l0
    ms_dev = max_over_ranks(dist, ev0.elapsed_time(ev1))
    # per-stage device times: re-run the timed chunks one by one (CUDA events inside the library, on the launch stream)
    stage_ms, split_ms, k_launches = np.zeros(4), np.zeros(eng.n_kernel_splits()), 0
    for chs in staged[W:]:
        for ch in chs:
            eng.launch(ch)
            stage_ms += eng.stage_ms()
            split_ms += eng.kernel_split_ms()
            k_launches += 1
    clocks = sampler.stop((t_load0, t_load1)) if rank == 0 else None
    value = world * K * per_step / (ms_dev * 1e-3)
    ch0 = staged[-1][0]
    yf_check = eng.unpermute("Yf", ch0["outs"]["Yf"].cpu().numpy().reshape(-1, 9, 3), ch0["perm"], ch0["ne"], ch0["nt"])
    assert np.all(np.isfinite(yf_check)), "non-finite abundances in the benchmark batch"

    # ---- end-to-end arm: public API, host buffers in, host rows out ------------------------------------------------
    for step in range(min(W, 2)):     # untimed: first-use effects of the public path (pinned staging buffers, workspace growth)
        run_points(model, slice_params(w, pts, step * world + rank), devices=[local])
    barrier(dist)
    h2d = d2h = 0
    t0 = time.perf_counter()
    for step in range(W, W + K):
        rows = run_points(model, slice_params(w, pts, step * world + rank), devices=[local])
        h2d += eng.last_h2d_bytes
        d2h += eng.last_d2h_bytes
    torch.cuda.synchronize()
    t_e2e = time.perf_counter() - t0
    barrier(dist)
    t_e2e = max_over_ranks(dist, t_e2e)
    e2e_value = world * K * per_step / t_e2e
    assert rows.shape == (per_step, 29) and np.all(np.isfinite(rows))
    total_launches = sum_over_ranks(dist, eng.launch_count() - launches0)

    # ---- roofline ---------------------------------------------------------------------------------------------------
    # Dominant kernels: the K-plane builder (FP64 pipe).  achieved = EXECUTED FP64 flop per launch (thread-level DFMA x2 +
    # DMUL + DADD + DMMA x512 per warp instruction, from the committed ncu counter pass of this library version on the same
    # slices, profiles/r02_fp64_flops.json) / the kernel's launch time measured live here with CUDA events.
    fp64_peak = eng.measure_fp64_peak()
    pairs, n9, n10, n11 = census
    prof, prof_src = load_flop_profile()
    names = eng.kernel_split_names()
    per_kernel = {}
    exec_flop_total, t_total = 0.0, 0.0
    stale = None
    if prof is not None:
        stale = prof.get("library_source_hash") != library_source_hash()
        for name, ms in zip(names, split_ms):
            pk = prof["kernels"].get(name)
            if pk is None or ms <= 0:
                continue
            # the profile holds flop per (E,E',T) triple of its own slices: scale by this run's triple count
            flop = pk["flop_per_triple"] * pairs / max(k_launches, 1)
            t = ms / max(k_launches, 1) * 1e-3
            per_kernel[name] = {"executed_flop_per_launch": flop, "avg_launch_ms": 1e3 * t, "achieved_tflops": flop / t / 1e12,
                                "frac_of_fp64_peak": flop / t / 1e12 / fp64_peak, "ncu_fp64_pipe_active_pct": pk.get("fp64_pipe_active_pct"),
                                "ncu_dram_bytes_per_launch": pk.get("dram_bytes_per_launch")}
            exec_flop_total += flop
            t_total += t
    Q = 63                      # canonical node count of SURVEY.md 8(d) (the reference's modal QAGS cost)
    flop_alg = (Q * (92 * n9 + 94 * n10 + 137 * n11) + 150 * pairs) / max(k_launches, 1)
    t_kernels = stage_ms[1] / max(k_launches, 1) * 1e-3
    achieved = exec_flop_total / t_total / 1e12 if t_total > 0 else None
    traffic = sum(v["ncu_dram_bytes_per_launch"] or 0 for v in per_kernel.values()) or None
    hbm_peak = None
    try:
        hbm_peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    hbm_peak_src = "MEASURED_PEAKS.json" if hbm_peak else "fallback (B200_PROFILING.md)"
    hbm_peak = hbm_peak or 6650.0
    solve_bytes = 5 * 8 * pairs / max(k_launches, 1)       # the solve stage reads every K entry once (G, F, S0 are <1% of that)
    t_solve = stage_ms[2] / max(k_launches, 1) * 1e-3
    solve_prof = (prof or {}).get("kernels", {}).get("solve_pdi_kernel", {})
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_dev / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": workload_config(cfg), "chunks_per_step": len(staged[-1]),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d / K), "d2h_bytes_per_step": int(d2h / K),
                    "api": "acropolis_b200.scans.run_points (model construction + source arrays + pinned H2D + kernels + D2H)",
                    "ms_per_step": 1e3 * t_e2e / K},
            "gpu_launches": int(dev_launches), "gpu_launches_total_all_arms": int(total_launches),
            "roofline": {"kernel": "kernel builder, stage (i): " + " + ".join(names),
                         "bound": "fp64", "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s",
                         "frac": (achieved / fp64_peak) if achieved else None, "traffic": traffic,
                         "peak_source": "DFMA micro-benchmark in this run (MEASURED_PEAKS.json has no FP64 entry; DMMA and DFMA "
                                        "share one pipe on B200: tools/microbench/fp64_pipes.cu, profiles/r02_fp64_pipes.txt)",
                         "achieved_is": "executed FP64 flop per launch (ncu thread-level DFMA*2+DMUL+DADD, DMMA*512) / live "
                                        "CUDA-event launch time, time-weighted over the builder kernels",
                         "flop_source": prof_src, "flop_profile_stale": stale, "per_kernel": per_kernel,
                         "avg_launch_ms": 1e3 * t_kernels, "share_of_step": stage_ms[1] / stage_ms.sum(),
                         "algorithmic_flop_per_launch": flop_alg,
                         "algorithmic_tflops": flop_alg / t_kernels / 1e12,
                         "algorithmic_speedup": (flop_alg / (exec_flop_total or np.nan)) if exec_flop_total else None,
                         "algorithmic_note": "SURVEY 8d convention (Q=63 nodes per (E,E',T) integral); the production path executes "
                                             "fewer operations than that convention counts (e->gamma plane from tabulated one-variable "
                                             "functions, no quadrature; F/G once per grid node for all T; 4-10 Laguerre nodes on the "
                                             "Wien tail), hence algorithmic_speedup > 1",
                         "census": {"pairs": pairs, "N_a9": n9, "N_a10": n10, "N_a11": n11, "Q": Q,
                                    "fractions": census_fractions(census)}},
            "roofline_solve": {"kernel": "solve_pdi_kernel (stage ii+iii)", "bound": "hbm", "achieved": solve_bytes / t_solve / 1e9,
                               "peak": hbm_peak, "unit": "GB/s", "frac": solve_bytes / t_solve / 1e9 / hbm_peak,
                               "peak_source": hbm_peak_src, "traffic": solve_prof.get("dram_bytes_per_launch"),
                               "traffic_source": prof_src, "avg_launch_ms": 1e3 * t_solve},
            "stage_ms_per_step": {k: float(v / K) for k, v in zip(("rates", "kernels", "solve_pdi", "matrix"), stage_ms)},
            "kernel_ms_per_step": {n: float(v / K) for n, v in zip(names, split_ms)},
            "clocks": clocks}
    if rank == 0:
        if world == 1 and not args.no_cpu:
            params = cpu_reference_sample(w, pts, W)         # the sample of the first timed slice
            mods = build_models(w, params)
            dt, cores, Yc, triples, n, specs = run_oracle_models(mods)
            sc = host_census(specs)
            slice_triples = float(census[0]) / K
            line["cpu_baseline"] = {"value": n / dt, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": sample_text(w) + "; here: slice %d, %d points, oracle/acro_oracle.c on %d threads, %.1f s"
                                              % (W % w["n_slices"], n, cores, dt),
                                    "sample_share_of_slice": {"points": n / per_step, "triples": triples / slice_triples},
                                    "sample_census": census_fractions(sc), "slice_census": census_fractions(census),
                                    "port_vs_reference": port_vs_reference_note()}
            # parity spot check of the same points (they span tau = 1e3 ... 1e10 s): CUDA rows vs oracle
            rows_g = run_points(model, params, devices=[local])
            Yg = rows_g[:, 2:].reshape(len(params), 3, 9).transpose(0, 2, 1)
            worst, kk = 0.0, 0
            for i, m in enumerate(mods):
                if m.is_trivial():
                    continue
                big = Yc[kk] > 1e-30
                worst = max(worst, float(np.max(np.abs(Yg[i][big] / Yc[kk][big] - 1))))
                kk += 1
            # the oracle at the reference's eps = 1e-3 carries the reference's quadrature noise (pdi / T integrals: up to 1e-3 on
            # Yf where the abundances change by O(1)); every 4th sample point again with the oracle run to eps = 1e-7
            sub = [i for i in range(0, len(mods), 4) if not mods[i].is_trivial()]
            _, _, Yt, _, _, _ = run_oracle_models([mods[i] for i in sub], eps=1e-7)
            worst_t = max(float(np.max(np.abs(Yg[i][Yt[k] > 1e-30] / Yt[k][Yt[k] > 1e-30] - 1))) for k, i in enumerate(sub)) if sub else None
            line["parity_vs_oracle_on_sample"] = {"max_rel_diff_Yf": worst, "oracle_eps": 1e-3,
                                                  "max_rel_diff_Yf_vs_converged_oracle": worst_t, "converged_oracle_eps": 1e-7,
                                                  "converged_subset": "every 4th point of the sample (%d points)" % len(sub),
                                                  "tau_span_s": [min(p["tau"] for p in params), max(p["tau"] for p in params)]}
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()
    return 0


def gpu_point(args, w):
    """cfg1 / cfg2: one model point; a step is one run_disintegration()."""
    import torch
    from acropolis_b200.engine import Engine
    cfg = args.config
    rank, world, local, dist = dist_setup(args.gpus)     # N > 1: N replicas of the same point (no sharding of one point)
    W, K = args.warmup, args.steps
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    cls = model_class(w["model"])
    m = cls(*w["args"])
    eng = Engine.get(m._sII, local)
    spec = m.point_spec()
    ch = eng.stage([spec], ("Yf", "status"), resident=True)
    t_load0 = time.time()
    for _ in range(max(W, 200)):      # a single point is ~1 ms of GPU work: enough launches for nvidia-smi to see the GPU busy
        eng.launch(ch)
    eng.stream.synchronize()
    barrier(dist)
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    l0 = eng.launch_count()
    stage_ms = np.zeros(4)
    for a, b in evs:
        flush_l2(torch, eng.device)
        torch.cuda.current_stream().synchronize()
        a.record(eng.stream)
        eng.launch(ch)
        b.record(eng.stream)
        b.synchronize()
        stage_ms += eng.stage_ms()
    barrier(dist)
    clocks = sampler.stop((t_load0, time.time())) if rank == 0 else None
    dev_launches = eng.launch_count() - l0
    ms_dev = max_over_ranks(dist, sum(a.elapsed_time(b) for a, b in evs))
    value = world * K / (ms_dev * 1e-3)
    # e2e: construct the model, run_disintegration() (H2D of the point, kernels, D2H of the matrices, expm call)
    for _ in range(min(W, 2)):
        cls(*w["args"]).run_disintegration()
    barrier(dist)
    t0 = time.perf_counter()
    for _ in range(K):
        Y = cls(*w["args"]).run_disintegration()
    torch.cuda.synchronize()
    t_e2e = max_over_ranks(dist, time.perf_counter() - t0)
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms_dev / K,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(cfg), "latency_ms_device": ms_dev / K, "latency_ms_e2e": 1e3 * t_e2e / K,
            "e2e": {"value": world * K / t_e2e, "unit": UNIT, "h2d_bytes_per_step": int(eng.last_h2d_bytes),
                    "d2h_bytes_per_step": int(eng.last_d2h_bytes), "api": "Model(...).run_disintegration()"},
            "gpu_launches": int(dev_launches),
            "stage_ms_per_step": {k: float(v / K) for k, v in zip(("rates", "kernels", "solve_pdi", "matrix"), stage_ms)},
            "Y_D_mean": float(Y[2, 0]), "clocks": clocks}
    if rank == 0:
        if not args.no_cpu:
            dt, cores, Yc, triples, n, _ = run_oracle_models([m], threads=1)
            big = Yc[0] > 1e-30
            line["cpu_baseline"] = {"value": 1.0 / dt, "unit": UNIT, "cores": 1, "kind": "port",
                                    "sample": "the point itself, oracle/acro_oracle.c on 1 thread, %.2f s" % dt,
                                    "port_vs_reference": port_vs_reference_note()}
            line["parity_vs_oracle_on_sample"] = {"max_rel_diff_Yf": float(np.max(np.abs(Y[big] / Yc[0][big] - 1))), "oracle_eps": 1e-3}
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()
    return 0


def gpu_fastscan(args, w):
    """cfg3b / cfg4: the shipped example scans verbatim through BufferedScanner.perform_scan (200 full points + 39 800
    rescale-only points; under torchrun the slow axis is sharded over the ranks and the rows are all-gathered)."""
    import torch
    os.environ["ACROPOLIS_B200_DIST"] = "1"
    from acropolis_b200.models import DecayModel, AnnihilationModel
    from acropolis_b200.scans import BufferedScanner, ScanParameter, run_models
    from acropolis_b200.engine import Engine
    cfg = args.config
    rank, world, local, dist = dist_setup(args.gpus)
    W, K = args.warmup, args.steps
    if w["which"] == "scan_decay_tau_aa":      # examples/scan_decay_tau_aa:25-32
        mk = lambda: BufferedScanner(DecayModel, mphi=ScanParameter(0., 3., 200), tau=1e7, temp0=10.,
                                     n0a=ScanParameter(-14., -3., 200, fast=True), bree=0., braa=1.)
        extra = None
    elif w["which"] == "scan_decay_full":      # the benchmark scan (cfg3) through the public scan call
        mk = lambda: BufferedScanner(DecayModel, mphi=ScanParameter(0., 3., 200), tau=ScanParameter(3., 10., 200), temp0=10.,
                                     n0a=1e-10, bree=0., braa=1.)
        extra = None
    else:                                      # examples/scan_pwave_ee:25-32
        mk = lambda: BufferedScanner(AnnihilationModel, mchi=ScanParameter(0., 3., 200), a=0.,
                                     b=ScanParameter(-23., -10., 200, fast=True), tempkd=1., bree=1., braa=0.)
        from acropolis_b200.ext.benchmarks import BenchmarkModel1, BenchmarkModel2, BenchmarkModel3
        kw = dict(mchi=10., delta=1e-2, gammad=1e-3, gammav=1e-12)
        extra = lambda: run_models([B(**kw) for B in (BenchmarkModel1, BenchmarkModel2, BenchmarkModel3)], devices=[local])
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(min(W, 1)):
        mk().perform_scan(devices=[local])
    barrier(dist)
    eng = Engine.get(DecayModel(10., 1e5, 10., 1e-10, 0., 1.)._sII, local)
    l0 = eng.launch_count()
    t_load0 = time.time()
    t0 = time.perf_counter()
    for _ in range(K):
        rows = mk().perform_scan(devices=[local])
        if extra is not None and rank == 0:
            Yb = extra()
    torch.cuda.synchronize()
    barrier(dist)
    t = max_over_ranks(dist, time.perf_counter() - t0)
    clocks = sampler.stop((t_load0, time.time())) if rank == 0 else None
    npts = rows.shape[0] + (3 if extra is not None else 0)
    value = K * npts / t
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": 1e3 * t / K,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(cfg), "points_per_step": int(npts), "finite_rows": int(np.count_nonzero(np.all(np.isfinite(rows), axis=1))),
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": None, "d2h_bytes_per_step": None,
                    "api": "BufferedScanner(...).perform_scan() wall clock incl. model construction, copies and the final gather"},
            "note": "value == e2e: this workload is only defined through the public scan API (host pipeline + device stages)",
            "gpu_launches": int(eng.launch_count() - l0), "clocks": clocks}
    if rank == 0:
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="cfg3", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    a = ap.parse_args()
    sys.exit(main_reference(a) if a.impl == "reference" else main_gpu(a))
