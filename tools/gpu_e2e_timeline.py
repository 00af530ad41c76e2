"""Where does the wall time of scans.run_points go?  Runs it on bench slices with CUDA events around every group's launch.
usage: gpu_e2e_timeline.py"""
import os, sys, time, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import acropolis_b200.flags as flags; flags.verbose = False
from acropolis_b200.models import DecayModel
from acropolis_b200.engine import Engine
from acropolis_b200.scans import run_points
import bench
from functools import partial
W = bench.WORKLOADS["cfg3"]
model = partial(DecayModel, **W["fixed"])
pts = bench.grid_points(W)
slice_params = lambda p, k: bench.slice_params(W, p, k)
evs = []
orig_launch = Engine.launch
def launch(self, ch, stages=7 | 8):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(self.stream); orig_launch(self, ch, stages); b.record(self.stream)
    evs.append((a, b, ch["cnt"]["n_points"], time.perf_counter()))     # with lanes > 1 the groups overlap: the gaps are not idle time
Engine.launch = launch
import itertools
CFG = [(40, 5, 2)]     # (first_group, growth, lanes): the defaults of run_points; add tuples to compare
for (fg, gr, ln), rep in itertools.product(CFG, range(4)):
    par = slice_params(pts, rep)
    torch.cuda.synchronize()
    del evs[:]
    t0 = time.perf_counter()
    z = torch.cuda.Event(enable_timing=True); z.record(torch.cuda.current_stream())
    rows = run_points(model, par, devices=[0], first_group=fg, growth=gr, lanes=ln)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    if rep >= 2:
        print("first_group %d growth %d lanes %d:" % (fg, gr, ln), "call %.1f ms wall; groups (points, host launch time since start [ms], GPU ms, idle gap before [ms]):" % (1e3 * (t1 - t0)))
        prev = None
        for a, b, n, th in evs:
            gap = prev.elapsed_time(a) if prev is not None else float("nan")
            print("   %4d  %7.2f  %7.2f  %6.2f" % (n, 1e3 * (th - t0), a.elapsed_time(b), gap))
            prev = b
        print("   GPU span first launch -> last end: %.2f ms; sum of groups %.2f ms" % (evs[0][0].elapsed_time(evs[-1][1]), sum(a.elapsed_time(b) for a, b, _, _ in evs)))
