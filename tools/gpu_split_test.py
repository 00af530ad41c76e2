import os, sys, numpy as np, torch
sys.path.insert(0, "/root/repo")
import acropolis_b200.flags as flags; flags.verbose = False
from acropolis_b200 import _lib
from acropolis_b200.engine import Engine
from acropolis_b200.models import DecayModel
from bench import grid_points, slice_params, build_models
models = [m for m in build_models(slice_params(grid_points(), 1)) if not m.is_trivial()]
specs = DecayModel.point_specs(models)
eng = Engine.get(models[0]._sII)
def run(chs, reps=4):
    for _ in range(2):
        for ch in chs: eng.launch(ch)
    eng.stream.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(eng.stream)
    for _ in range(reps):
        for ch in chs: eng.launch(ch)
    b.record(eng.stream); b.synchronize()
    per = []
    for ch in chs:
        eng.launch(ch); per.append(eng.stage_ms().round(2).tolist())
    return a.elapsed_time(b) / reps, per
allc = [eng.stage(specs, ("Yf",), resident=True)]
t, per = run(allc); print("one launch of %d points: %.2f ms" % (len(specs), t), per)
for n1 in (384, 200, 600):
    chs = [eng.stage(specs[:n1], ("Yf",), resident=True), eng.stage(specs[n1:], ("Yf",), resident=True)]
    t, per = run(chs); print("split %d + %d: %.2f ms" % (n1, len(specs) - n1, t), per)
