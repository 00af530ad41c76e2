"""Diagnostic run on a GPU box: prints the parity errors of every stage against the golden fixtures."""
import os, sys, time, glob
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from acropolis_b200 import _lib
from acropolis_b200.input import shared_input_interface
from acropolis_b200.engine import Engine, PointSpec
import acropolis_b200.flags as flags
flags.verbose = False
G = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def rel(a, b, floor=1e-150):
    a, b = np.asarray(a), np.asarray(b)
    m = np.abs(b) > floor
    if not m.any():
        return 0.0
    return float(np.max(np.abs(a[m] / b[m] - 1)))


def spec_from(z):
    SC = z["SC"]
    mode = _lib.SC_DENSE if np.any(SC != 0) else _lib.SC_NONE
    return PointSpec(float(z["E0"]), z["E_grid"], z["T"], z["S0"], float(z["t_at_Tmax"]), mode, SC if mode else None, None)


ii = shared_input_interface()
eng = Engine.get(ii)
print("fp64 peak TFLOP/s:", eng.measure_fp64_peak())
names = sys.argv[1:] or ["cfg1_decay", "cfg1_decay_tight", "decay_ee", "decay_ee_tight", "decay_nemin", "decay_mixed",
                         "decay_lowT", "decay_highT", "decay_big", "cfg2_annih", "annih_pwave", "resonance_b3",
                         "decay_hires_ne60_nt50"]
for name in names:
    z = np.load(os.path.join(G, name + ".npz"))
    sp = spec_from(z)
    NT, NE = len(sp.T), len(sp.E_grid)
    t0 = time.time()
    out = eng.run([sp], want=("G", "F", "pdi", "mpdi", "mdcy", "Yf", "status"))
    torch.cuda.synchronize()
    dt = time.time() - t0
    Gd = out["G"].reshape(NT, 3, NE)
    Fd = out["F"].reshape(NT, 3, NE)
    line = "%-24s NE=%3d NT=%3d %.3fs st=%d  G %.1e  F %.1e  pdi %.1e  mpdi %.1e  mdcy %.1e  Yf %.1e" % (
        name, NE, NT, dt, out["status"][0], rel(Gd, z["G"]), rel(Fd, z["F"]), rel(out["pdi"], z["pdi"], 1e-100),
        rel(out["mpdi"][0], z["mpdi"], 1e-100), rel(out["mdcy"][0], z["mdcy"], 1e-100), rel(out["Yf"][0], z["Yf"], 1e-30))
    print(line, " stage_ms", np.round(eng.stage_ms(), 3))
    for it in z["k_at"]:
        K, Gk = eng.kernels_of(sp, int(it))
        Kr = z["K_it%d" % it]
        print("    K at it=%d:" % it, " ".join("%s %.1e" % (n, rel(K[k], Kr[k])) for k, n in enumerate(["gg", "eg", "ge-", "ge+", "ee"])),
              " nnz mismatch", int(((K != 0) != (Kr != 0)).sum()))
    # per-reaction pdi error detail for the worst
    r = np.abs(out["pdi"] / z["pdi"] - 1) * (z["pdi"] > 1e-100)
    print("    pdi worst per rid:", np.array2string(r.max(axis=0), precision=1, max_line_width=250))
# expm via transfer
z = np.load(os.path.join(G, "cfg1_decay.npz"))
Y, ex = eng.transfer(z["mpdi"][None], z["mdcy"][None], 1.0, float(z["t_at_Tmax"]), want_expm=True)
print("expm vs scipy:", rel(ex[0], z["expm"], 1e-30), " Yf via transfer:", rel(Y[0], z["Yf"], 1e-30))
print("launches:", eng.launch_count())
