"""A/B of kernel-builder builds/knobs on a GPU box: time + worst relative difference to the plain GL-96 planes."""
import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import acropolis_b200.flags as flags; flags.verbose = False
import acropolis_b200.params as params
from acropolis_b200 import _lib
from acropolis_b200.engine import Engine
from bench import grid_points, slice_params, build_models
models = [m for m in build_models(slice_params(grid_points(), 0)[::5]) if not m.is_trivial()]
specs = [m.point_spec() for m in models]
eng = Engine.get(models[0]._sII)
def planes(mode, order):
    params.kernel_mode, params.quad_order = mode, order
    eng._sync_params()
    ch = eng.stage(specs, ("Yf",), resident=True)
    for rep in range(3):
        eng.launch(ch, _lib.STAGE_RATES | _lib.STAGE_KERNELS); ms = eng.stage_ms()
    n = ch["cnt"]["n_pairs_total"]
    return eng._ws[:5 * n * 8].view(torch.float64).clone().view(5, n), ms[1]
os.environ["ACRO_HALF_RANGE"] = "0"
ref, t = planes(0, 96)
print("fast Q=96 reference: %.1f ms" % t)
for hr in sys.argv[1:] or ["0"]:
    os.environ["ACRO_HALF_RANGE"] = hr
    K, t = planes(0, 64)
    big = ref.abs() > 1e-150
    rel = torch.where(big, (K / torch.where(big, ref, torch.ones_like(ref)) - 1).abs(), torch.zeros_like(ref))
    print("%s half_range=%s fast Q=64: %.1f ms  worst rel vs fast Q=96 per plane %s" % (
        os.environ.get("ACROPOLIS_B200_LIB", "default").split('/')[-1], hr, t, ["%.1e" % v for v in rel.max(dim=1).values.tolist()]), flush=True)
