"""A/B of kernel-builder modes on a GPU box: time + worst relative difference to the plain GL-96 planes.
usage: gpu_ab_variants.py [stride]"""
import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import acropolis_b200.flags as flags; flags.verbose = False
import acropolis_b200.params as params
from acropolis_b200 import _lib
from acropolis_b200.engine import Engine
from bench import grid_points, slice_params, build_models
stride = int(sys.argv[1]) if len(sys.argv) > 1 else 5
models = [m for m in build_models(slice_params(grid_points(), 0)[::stride]) if not m.is_trivial()]
specs = [m.point_spec() for m in models]
eng = Engine.get(models[0]._sII)
def planes(mode, order):
    params.kernel_mode, params.quad_order = mode, order
    eng._sync_params()
    ch = eng.stage(specs, ("Yf",), resident=True)
    for rep in range(3):
        eng.launch(ch, _lib.STAGE_RATES | _lib.STAGE_KERNELS); ms = eng.stage_ms()
    if mode == 0:
        print("   split shared/wien [ms]:", eng.kernel_split_ms().round(2).tolist(), flush=True)
    n = ch["cnt"]["n_pairs_total"]
    K = eng._ws[:5 * n * 8].view(torch.float64).clone().view(5, n)
    eng.launch(ch, _lib.STAGE_ALL); eng.stream.synchronize()
    return K, ms[1], ch["outs"]["Yf"].clone()
ref, t, yref = planes(1, 96)
print("plain GL-96 reference: %.1f ms (%d points)" % (t, len(specs)), flush=True)
for mode, order, name in ((2, 64, "per-T fast"), (0, 64, "T-shared + fast")):
    K, t, Y = planes(mode, order)
    big = ref.abs() > 1e-150
    rel = torch.where(big, (K / torch.where(big, ref, torch.ones_like(ref)) - 1).abs(), torch.zeros_like(ref))
    ybig = yref.abs() > 1e-30
    print("%-16s Q=%d: %.1f ms  worst rel vs plain GL-96 per plane %s  nan %d  support mismatch %d  Yf rel %.1e" % (
        name, order, t, ["%.1e" % v for v in rel.max(dim=1).values.tolist()], int(torch.isnan(K).sum()),
        int(((K != 0) != (ref != 0)).sum()), float(((Y[ybig] / yref[ybig]) - 1).abs().max())), flush=True)
    if rel.max() > 2e-8:
        a, cnt = eng.assemble(specs)
        for pl in (1, 3, 4):
            v, idx = torch.topk(rel[pl], 3)
            for vv, t_ in zip(v.tolist(), idx.tolist()):
                s = int(np.searchsorted(a["sys_k_off"], t_, side="right") - 1)
                loc = t_ - int(a["sys_k_off"][s]); pt = int(a["sys_point"][s]); ne = int(a["pt_ne"][pt])
                i = 0
                while (i + 1) * ne - (i + 1) * i // 2 <= loc: i += 1
                j = i + loc - (i * ne - i * (i - 1) // 2)
                E = a["e_grid"][a["pt_e_off"][pt] + i]; Ep = a["e_grid"][a["pt_e_off"][pt] + j]; T = a["sys_T"][s]
                print("   plane %d rel %.2e E=%.6g Ep=%.6g T=%.4g (Tmin %.3g Tmax %.3g) got %.6e ref %.6e" % (
                    pl, vv, E, Ep, T, a["sys_T"][a["pt_sys_off"][pt]], a["sys_T"][a["pt_sys_off"][pt] + a["pt_nt"][pt] - 1], K[pl][t_].item(), ref[pl][t_].item()))
