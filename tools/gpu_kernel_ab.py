"""A/B on a GPU box: production (fast) vs plain Gauss-Legendre evaluation of the K planes on a strided subset of one
benchmark slice -- stage times and the worst relative difference per plane (entries > 1e-150)."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import acropolis_b200.flags as flags
flags.verbose = False
import acropolis_b200.params as params
from acropolis_b200 import _lib
from acropolis_b200.engine import Engine
from bench import grid_points, slice_params, build_models

stride = int(sys.argv[1]) if len(sys.argv) > 1 else 5
models = [m for m in build_models(slice_params(grid_points(), 0)[::stride]) if not m.is_trivial()]
specs = [m.point_spec() for m in models]
eng = Engine.get(models[0]._sII)
res = {}
for mode, order in ((1, 64), (1, 96), (0, 64), (0, 32), (0, 96)):
    params.kernel_mode, params.quad_order = mode, order
    eng._sync_params()
    ch = eng.stage(specs, ("Yf",), resident=True)
    for rep in range(2):
        eng.launch(ch, _lib.STAGE_RATES | _lib.STAGE_KERNELS)
        ms = eng.stage_ms()
    npairs = ch["cnt"]["n_pairs_total"]
    K = eng._ws[:5 * npairs * 8].view(torch.float64).clone().view(5, npairs)
    eng.launch(ch, _lib.STAGE_ALL)
    eng.stream.synchronize()
    Yf = ch["outs"]["Yf"].clone()
    res[(mode, order)] = (K, Yf, ms[1])
    print("mode=%d Q=%d  kernels stage %.1f ms for %d pairs (%d points)" % (mode, order, ms[1], npairs, len(specs)), flush=True)
ref = res[(1, 96)][0]
for key, (K, Yf, ms) in res.items():
    big = ref.abs() > 1e-150
    rel = torch.where(big, (K / torch.where(big, ref, torch.ones_like(ref)) - 1).abs(), torch.zeros_like(ref))
    supp = ((K != 0) != (ref != 0)).sum().item()
    yref = res[(1, 96)][1]
    ybig = yref.abs() > 1e-30
    yrel = ((Yf[ybig] / yref[ybig]) - 1).abs().max().item()
    print("mode=%d Q=%d vs plain Q=96: worst rel per plane" % key, ["%.1e" % v for v in rel.max(dim=1).values.tolist()],
          "support mismatches", supp, " Yf rel %.1e" % yrel, " speedup vs plain64 %.2fx" % (res[(1, 64)][2] / ms))

# --- locate the worst entries of the gamma->e plane between plain Q=64 and plain Q=96
a, cnt = eng.assemble(specs)
K64, K96 = res[(1, 64)][0][3], res[(1, 96)][0][3]
big = K96.abs() > 1e-150
rel = torch.where(big, (K64 / torch.where(big, K96, torch.ones_like(K96)) - 1).abs(), torch.zeros_like(K96))
vals, idx = torch.topk(rel, 12)
print("entries with rel > 1e-6: %d of %d" % ((rel > 1e-6).sum().item(), big.sum().item()))
for v, t in zip(vals.tolist(), idx.tolist()):
    s = int(np.searchsorted(a["sys_k_off"], t, side="right") - 1)
    loc = t - int(a["sys_k_off"][s])
    pt = int(a["sys_point"][s]); ne = int(a["pt_ne"][pt])
    i = 0
    while (i + 1) * ne - (i + 1) * i // 2 <= loc:
        i += 1
    j = i + loc - (i * ne - i * (i - 1) // 2)
    E = a["e_grid"][a["pt_e_off"][pt] + i]; Ep = a["e_grid"][a["pt_e_off"][pt] + j]; T = a["sys_T"][s]
    print("rel %.2e  E=%.17g Ep=%.17g T=%.17g  K64=%.6e K96=%.6e fast64=%.6e" % (v, E, Ep, T, K64[t].item(), K96[t].item(), res[(0, 64)][0][3][t].item()))
