#!/usr/bin/env python3
"""Benchmark of the hot path: scan points/s for the 2-D DecayModel scan (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W          # our arm (N>1: launched by torchrun, one rank per GPU)
    python bench.py --impl reference --gpus N --steps K --warmup W   # CPU arm: the oracle port on all host threads

Workload (BASELINE.json configs[2], SURVEY.md 8d "M3"): BufferedScanner(DecayModel, mphi=ScanParameter(0,3,200),
tau=ScanParameter(3,10,200), temp0=10, n0a=1e-10, bree=0, braa=1) -- 40 000 points, all full computations.  The grid
is flattened in the reference's order (mphi outer, tau inner) and cut into 40 interleaved slices L[k::40] of 1000
points (200 mphi x 5 tau each, 160 of them below threshold); the 40 slices together ARE the full scan and every
slice has the scan's cost mix.  One step = one slice per GPU (weak scaling: step t on rank r takes slice
(t*N + r) mod 40).  `value` is device-timed with the slices resident in HBM; `e2e` goes through the public
``acropolis_b200.scans.run_points`` call (model construction, source arrays, pinned H2D, kernels, D2H) per step.
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

N_GRID, N_SLICES = 200, 40
FIXED = dict(temp0=10., n0a=1e-10, bree=0., braa=1.)
METRIC, UNIT = "scan points/sec (DecayModel 2-D scan)", "points/s"


def grid_points():
    mphi, tau = np.logspace(0, 3, N_GRID), np.logspace(3, 10, N_GRID)
    return [dict(mphi=float(m), tau=float(t)) for m in mphi for t in tau]   # scans.py:193-197 order


def slice_params(pts, k):
    return pts[(k % N_SLICES)::N_SLICES]


def build_models(params):
    from acropolis_b200.models import DecayModel
    return [DecayModel(**p, **FIXED) for p in params]


def workload_config(extra=None):
    import acropolis_b200.params as prm
    c = {"workload": "cfg3 DecayModel mphi x tau 200x200 scan (examples/scan_decay_* ranges, no fast parameter); "
                     "step = one of 40 interleaved slices = 1000 scan points per GPU (160 below threshold)",
         "grid": [N_GRID, N_GRID], "points_per_step_per_gpu": N_GRID * N_GRID // N_SLICES, "NE_pd": prm.NE_pd,
         "NT_pd": prm.NT_pd, "quad_order": prm.quad_order, "wien_from": prm.wien_from,
         "l2": "inputs larger than L2: each step streams ~20 GB of kernel planes through HBM, no reuse between steps"}
    if extra:
        c.update(extra)
    return c


# ---------------------------------------------------------------------------------------------------------------
class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=lambda: [self.lines.append(l) for l in self.proc.stdout], daemon=True).start()
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        time.sleep(0.05)
        sm, mx, reasons = [], [], set()
        for l in list(self.lines):
            f = [x.strip() for x in l.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


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


def max_over_ranks(dist, v):
    import torch
    if dist is None:
        return v
    t = torch.tensor([v], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(dist, v):
    import torch
    if dist is None:
        return v
    t = torch.tensor([v], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


# ---------------------------------------------------------------------------------------------------------------
def cpu_reference_sample(pts, k, stride):
    """Bounded sample of slice k for the CPU arm: every `stride`-th point of the slice."""
    return slice_params(pts, k)[::stride]


def run_oracle_scan(params, threads=0):
    """The oracle port (oracle/acro_oracle.c) on all host threads over a list of scan points -> (seconds, threads,
    Yf, triples).  This is the CPU restatement of scans.py's Pool over run_disintegration(); see oracle/ header."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    from pyoracle import Oracle
    from acropolis_b200.engine import Engine
    models = build_models(params)
    work = [m for m in models if not m.is_trivial()]
    specs = [m.point_spec() for m in work]
    a, cnt = Engine.assemble(specs)
    o = Oracle(models[0]._sII)
    t0 = time.perf_counter()
    Yf, _ = o.run_points(a, cnt, n_threads=threads)
    dt = time.perf_counter() - t0
    used = min(threads if threads > 0 else (os.cpu_count() or 1), len(specs))
    triples = float(sum(len(s.T) * len(s.E_grid) * (len(s.E_grid) + 1) / 2 for s in specs))
    return dt, used, Yf, triples, len(models)


def main_reference(args):
    import acropolis_b200.flags as flags
    flags.verbose = False
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return 0
    pts = grid_points()
    stride = args.cpu_stride
    tot_t, tot_p, cores = 0.0, 0, 1
    for step in range(args.warmup + args.steps):
        params = cpu_reference_sample(pts, step, stride)
        dt, cores, _, _, n = run_oracle_scan(params)
        if step >= args.warmup:
            tot_t += dt
            tot_p += n
    value = tot_p / tot_t
    sample = "every %d-th point of one 1000-point slice per step (%d points/step)" % (stride, len(cpu_reference_sample(pts, 0, stride)))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / max(args.steps, 1), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config({"cpu_sample": sample}),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))
    return 0


def main_gpu(args):
    import torch
    import acropolis_b200.flags as flags
    flags.verbose = False
    os.environ["ACROPOLIS_B200_DIST"] = "0"      # ranks run independent slices; no gather inside run_points
    from acropolis_b200 import _lib
    from acropolis_b200.engine import Engine
    from acropolis_b200.scans import run_points
    from acropolis_b200.models import DecayModel
    from functools import partial
    import acropolis_b200.params as prm
    rank, world, local, dist = dist_setup(args.gpus)
    pts = grid_points()
    W, K = args.warmup, args.steps
    per_step = N_GRID * N_GRID // N_SLICES
    model = partial(DecayModel, **FIXED)
    ii = build_models(pts[:1])[0]._sII
    eng = Engine.get(ii, local)
    launches0 = eng.launch_count()

    # ---- device-timed arm: slices resident in HBM --------------------------------------------------------------
    staged = []
    for step in range(W + K):
        models = build_models(slice_params(pts, step * world + rank))
        specs = [m.point_spec() for m in models if not m.is_trivial()]
        staged.append([eng.stage([specs[i] for i in idx], ("Yf", "status"), resident=True) for idx in eng.chunks(specs)])
    eng.stream.synchronize()
    census = np.zeros(4)
    for chs in staged[W:]:
        for ch in chs:
            census += np.array(eng.count_work(ch), dtype=np.float64)
    for chs in staged[:W]:
        for ch in chs:
            eng.launch(ch)
    eng.stream.synchronize()
    sampler = ClockSampler(local)
    barrier(dist)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    stage_ms = np.zeros(4)
    l0 = eng.launch_count()
    ev0.record(eng.stream)
    for chs in staged[W:]:
        for ch in chs:
            eng.launch(ch)
    ev1.record(eng.stream)
    ev1.synchronize()
    barrier(dist)
    clocks = sampler.stop() if rank == 0 else None
    dev_launches = eng.launch_count() - l0
    ms_dev = max_over_ranks(dist, ev0.elapsed_time(ev1))
    # per-stage device times of the dominant kernel: re-run the timed chunks one by one (events inside the library)
    k_launches = 0
    for chs in staged[W:]:
        for ch in chs:
            eng.launch(ch)
            stage_ms += eng.stage_ms()
            k_launches += 1
    value = world * K * per_step / (ms_dev * 1e-3)
    ch0 = staged[-1][0]
    yf_check = eng.unpermute("Yf", ch0["outs"]["Yf"].cpu().numpy().reshape(-1, 9, 3), ch0["perm"], ch0["ne"], ch0["nt"])
    assert np.all(np.isfinite(yf_check)), "non-finite abundances in the benchmark batch"

    # ---- end-to-end arm: public API, host buffers in, host rows out ------------------------------------------------
    for step in range(min(W, 1)):
        run_points(model, slice_params(pts, step * world + rank), devices=[local])
    barrier(dist)
    h2d = d2h = 0
    t0 = time.perf_counter()
    for step in range(W, W + K):
        rows = run_points(model, slice_params(pts, step * world + rank), devices=[local])
        h2d += eng.last_h2d_bytes
        d2h += eng.last_d2h_bytes
    torch.cuda.synchronize()
    t_e2e = time.perf_counter() - t0
    barrier(dist)
    t_e2e = max_over_ranks(dist, t_e2e)
    e2e_value = world * K * per_step / t_e2e
    assert rows.shape == (per_step, 29) and np.all(np.isfinite(rows))
    total_launches = sum_over_ranks(dist, eng.launch_count() - launches0)

    # ---- roofline of the dominant kernel (the K-plane builder), FP64 pipe ----------------------------------------
    Q = 63                      # canonical node count of SURVEY.md 8(d) (the reference's modal QAGS cost)
    pairs, n9, n10, n11 = census
    flop_kernels = Q * (92 * n9 + 94 * n10 + 137 * n11) + 150 * pairs      # algorithmic, over the K timed steps
    fp64_peak = eng.measure_fp64_peak()
    t_kernels = stage_ms[1] * 1e-3
    achieved = flop_kernels / t_kernels / 1e12
    hbm_peak = None
    try:
        hbm_peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    hbm_peak_src = "MEASURED_PEAKS.json" if hbm_peak else "fallback (B200_PROFILING.md)"
    hbm_peak = hbm_peak or 6650.0
    solve_bytes = 5 * 8 * pairs       # the solve stage reads every K entry once (G, F, S0 are <1% of that)
    # DRAM traffic and executed-FP64 utilisation of the two kernel-builder kernels from the committed ncu captures of one
    # slice (profiles/r01_ncu_kernels_shared4_kernel.txt, r01_ncu_kernels_wien_kernel.txt): bytes per slice, i.e. per launch pair
    ncu = {"traffic_bytes_per_launch": 36.10e9 + 9.58e9, "algorithmic_bytes_per_launch": 5 * 8 * pairs / max(k_launches, 1),
           "fp64_pipe_active_pct": {"kernels_shared4_kernel": 38.6, "kernels_wien_kernel": 54.9},
           "issue_active_pct": {"kernels_shared4_kernel": 46.6, "kernels_wien_kernel": 65.8},
           "source": "profiles/r01_ncu_kernels_shared4_kernel.txt, profiles/r01_ncu_kernels_wien_kernel.txt"}
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_dev / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": workload_config({"chunks_per_step": len(staged[-1])}),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d / K), "d2h_bytes_per_step": int(d2h / K),
                    "api": "acropolis_b200.scans.run_points (model construction + source arrays + pinned H2D + kernels + D2H)",
                    "ms_per_step": 1e3 * t_e2e / K},
            "gpu_launches": int(dev_launches), "gpu_launches_total_all_arms": int(total_launches),
            "roofline": {"kernel": "kernel builder, stage (i): pair_closed_forms_kernel + kernels_shared4_kernel + kernels_wien_kernel<%d>" % prm.quad_order,
                         "bound": "fp64", "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s", "frac": achieved / fp64_peak,
                         "traffic": ncu["traffic_bytes_per_launch"],
                         "peak_source": "DFMA micro-benchmark in this run (MEASURED_PEAKS.json has no FP64 entry)",
                         "note": "achieved = ALGORITHMIC flop (SURVEY 8d: Q=63 nodes per (E,E',T) integral) / measured stage time; it "
                                 "can exceed the DFMA peak because F/G are evaluated once per (E,E',node) for all temperatures of a "
                                 "point and the Wien tail uses 4-24 Gauss-Laguerre nodes; executed-FP64 utilisation is in `ncu`",
                         "algorithmic_flop_per_launch": flop_kernels / max(k_launches, 1),
                         "avg_launch_ms": stage_ms[1] / max(k_launches, 1),
                         "census": {"pairs": pairs, "N_a9": n9, "N_a10": n10, "N_a11": n11, "Q": Q},
                         "share_of_step": stage_ms[1] / stage_ms.sum(), "ncu": ncu},
            "roofline_solve": {"kernel": "solve_pdi_kernel (stage ii+iii)", "bound": "hbm", "achieved": solve_bytes / (stage_ms[2] * 1e-3) / 1e9,
                               "peak": hbm_peak, "unit": "GB/s", "frac": solve_bytes / (stage_ms[2] * 1e-3) / 1e9 / hbm_peak,
                               "peak_source": hbm_peak_src, "traffic": 20.89e9,
                               "traffic_source": "profiles/r01_ncu_solve_pdi_kernel.txt (bytes per slice)"},
            "stage_ms_per_step": {k: float(v / K) for k, v in zip(("rates", "kernels", "solve_pdi", "matrix"), stage_ms)},
            "clocks": clocks}
    if rank == 0:
        if world == 1 and not args.no_cpu:
            params = cpu_reference_sample(pts, 0, args.cpu_stride)
            dt, cores, Yc, triples, n = run_oracle_scan(params)
            slice_triples = float(census[0]) / K
            line["cpu_baseline"] = {"value": n / dt, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": "every %d-th point of slice 0 (%d points, %.3g of a slice's (E,E',T) triples), "
                                              "oracle/acro_oracle.c on %d threads, %.1f s" % (args.cpu_stride, n, triples / slice_triples, cores, dt)}
            # parity spot check of the same points: CUDA rows vs oracle
            rows_g = run_points(model, params, devices=[local])
            Yg = rows_g[:, 2:].reshape(len(params), 3, 9).transpose(0, 2, 1)
            mods = build_models(params)
            worst, kk = 0.0, 0
            for i, m in enumerate(mods):
                if m.is_trivial():
                    continue
                big = Yc[kk] > 1e-30
                worst = max(worst, float(np.max(np.abs(Yg[i][big] / Yc[kk][big] - 1))))
                kk += 1
            line["parity_vs_oracle_on_sample"] = worst
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
    ap.add_argument("--cpu-stride", type=int, default=25, help="CPU arm: every n-th point of a slice")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    a = ap.parse_args()
    sys.exit(main_reference(a) if a.impl == "reference" else main_gpu(a))
