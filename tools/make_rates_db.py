#!/usr/bin/env python3
"""Regenerate (and optionally extend) the rate table rates_db.npz on the GPU -- the B200 replacement of the reference's
stale tools/create_db.py (:34-55; it calls methods that no longer exist) and tools/compress_ratedb.py (:12-19).

The table holds log10 of the thermal pair-creation rate (gamma gamma_th -> e+ e-, cascade.py:336-358) and of the e+- inverse
Compton rate (cascade.py:469-485) on a (T, E) grid, log-spaced: Tnum x Enum = 750 x 450 over log10 T in [-6,-1], log10 E in
[0,3] (params.py:54-59), zeros replaced by approx_zero = 1e-200.  Here every entry is one warp's converged 2-D integral
(direct_rates_kernel through acro_direct_rates); the whole table takes seconds.

    python tools/make_rates_db.py --check                       # compare with the shipped table, print the statistics
    python tools/make_rates_db.py --out my_rates_db.npz [--tlog -7 0 --tnum 1050]   # write a (wider) table

Run on a machine with a GPU (gpurun)."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--out")
    ap.add_argument("--tlog", type=float, nargs=2, default=None)
    ap.add_argument("--elog", type=float, nargs=2, default=None)
    ap.add_argument("--tnum", type=int, default=None)
    ap.add_argument("--enum", type=int, default=None)
    ap.add_argument("--report", default=None, help="write the comparison statistics to this JSON file")
    a = ap.parse_args()
    import acropolis_b200.flags as flags
    import acropolis_b200.params as params
    flags.verbose = False
    from acropolis_b200.engine import Engine
    from acropolis_b200.input import shared_input_interface
    from acropolis_b200.db import import_data_from_db
    eng = Engine.get(shared_input_interface())
    Tl = a.tlog or (params.Tmin_log, params.Tmax_log)
    El = a.elog or (params.Emin_log, params.Emax_log)
    Tn, En = a.tnum or params.Tnum, a.enum or params.Enum
    Tg, Eg = np.logspace(Tl[0], Tl[1], Tn), np.logspace(El[0], El[1], En)
    TT, EE = np.meshgrid(Tg, Eg, indexing="ij")          # row index jT * Enum + jE (db.py:95-98)
    t0 = time.perf_counter()
    r = eng.direct_rates(EE.reshape(-1), TT.reshape(-1))
    dt = time.perf_counter() - t0
    with np.errstate(divide="ignore"):
        tab = np.log10(np.maximum(r, params.approx_zero))
    print("%d x %d entries in %.2f s" % (Tn, En, dt))
    if a.out:
        np.savez_compressed(a.out, log10_rates=tab.reshape(Tn, En, 2), Emin_log=El[0], Emax_log=El[1], Tmin_log=Tl[0], Tmax_log=Tl[1])
        print("written", a.out)
    if a.check:
        ref = import_data_from_db()
        assert ref.shape == tab.shape, "--check needs the shipped table's grid"
        stats = {"entries": int(ref.shape[0]), "seconds": dt}
        for c, name in enumerate(("pair_creation", "inverse_compton")):
            live = (ref[:, c] > -150) & (tab[:, c] > -150)
            d = np.abs(10.0 ** (tab[live, c] - ref[live, c]) - 1.0)
            stats[name] = {"entries_above_1e-150": int(live.sum()), "max_rel_diff": float(d.max()), "median_rel_diff": float(np.median(d)),
                           "p99_rel_diff": float(np.quantile(d, 0.99)),
                           "floored_in_both": int(np.count_nonzero((ref[:, c] <= -150) & (tab[:, c] <= -150))),
                           "floored_in_one_only": int(np.count_nonzero((ref[:, c] <= -150) != (tab[:, c] <= -150)))}
        print(json.dumps(stats, indent=1))
        if a.report:
            json.dump(stats, open(a.report, "w"), indent=1)


if __name__ == "__main__":
    main()
