#!/usr/bin/env python3
"""Port-to-reference factor on the cfg3 sub-grid (VERDICT r1 item 1 / SURVEY.md 8d "CPU baseline timing").

The reference (hep-mh/acropolis v1.3.1, pure Python) cannot travel to the GPU box, so bench.py's CPU arm times the C
restatement oracle/acro_oracle.c there (kind "port").  This script evidences the chain port -> reference: it runs, in the
BUILD container, on the same N x N sub-grid with the same log ranges as the 200x200 scan,

  (a) the UNMODIFIED reference: BufferedScanner(DecayModel, ...).perform_scan(cores=-1)   (scans.py:179-245)
  (b) the oracle port on all host threads (the function bench.py --impl reference calls)

and writes wall times, points/s, the factor and the row-wise agreement to profiles/r02_port_vs_reference.json.

    NUMBA_CACHE_DIR=/tmp/numba_cache python tools/port_vs_reference.py [--n 6]
"""
import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("ACROPOLIS_REFERENCE", "/root/reference")

REF_SNIPPET = r"""
import sys, time, json, os
sys.path.insert(0, %(ref)r)
import numpy as np
import acropolis.flags as flags
flags.verbose = False
from acropolis.models import DecayModel
from acropolis.scans import BufferedScanner, ScanParameter
DecayModel(10., 1e5, 10., 1e-10, 0., 1.).run_disintegration()      # warm the numba cache (SURVEY 8d)
t0 = time.time()
rows = BufferedScanner(DecayModel, mphi=ScanParameter(0, 3, %(n)d), tau=ScanParameter(3, 10, %(n)d), temp0=10., n0a=1e-10,
                       bree=0., braa=1.).perform_scan(cores=-1)
dt = time.time() - t0
np.save(%(out)r, rows)
print(json.dumps({"wall_s": dt, "cores": os.cpu_count()}))
"""


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=6)
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r02_port_vs_reference.json"))
    a = ap.parse_args()
    n = a.n
    tmp = "/tmp/pvr_rows_%d.npy" % n
    env = dict(os.environ)
    env.setdefault("NUMBA_CACHE_DIR", "/tmp/numba_cache")
    out = subprocess.check_output([sys.executable, "-c", REF_SNIPPET % dict(ref=REF, n=n, out=tmp)], env=env, text=True)
    ref = json.loads(out.strip().splitlines()[-1])
    rows_ref = np.load(tmp)

    sys.path.insert(0, ROOT)
    import acropolis_b200.flags as flags
    flags.verbose = False
    import bench
    mphi, tau = np.logspace(0, 3, n), np.logspace(3, 10, n)
    params = [dict(mphi=float(m), tau=float(t)) for m in mphi for t in tau]
    dt, cores, Yf, triples, npts = bench.run_oracle_scan(params)
    # agreement of the non-trivial rows (reference rows are [mphi, tau, 9 mean, 9 high, 9 low])
    mods = bench.build_models(params)
    worst, k = 0.0, 0
    for i, m in enumerate(mods):
        if m.is_trivial():
            continue
        Yr = rows_ref[i, 2:].reshape(3, 9).T
        big = Yr > 1e-30
        worst = max(worst, float(np.max(np.abs(Yf[k][big] / Yr[big] - 1))))
        k += 1
    res = {"grid": [n, n], "ranges": "mphi 10^0..10^3 MeV, tau 10^3..10^10 s (cfg3), temp0=10, n0a=1e-10, bree=0, braa=1",
           "points": n * n, "trivial_points": int(sum(m.is_trivial() for m in mods)), "triples": triples,
           "reference": {"impl": "hep-mh/acropolis v1.3.1 BufferedScanner.perform_scan(cores=-1), eps=1e-3",
                         "wall_s": ref["wall_s"], "cores": ref["cores"], "points_per_s": n * n / ref["wall_s"]},
           "port": {"impl": "oracle/acro_oracle.c (run_oracle_scan of bench.py), eps=1e-3", "wall_s": dt, "cores": cores,
                    "points_per_s": npts / dt},
           "port_over_reference": (npts / dt) / (n * n / ref["wall_s"]),
           "max_rel_diff_Yf_port_vs_reference": worst,
           "where": "build container (no GPU), %d host cores" % (os.cpu_count() or 0),
           "when": time.strftime("%Y-%m-%dT%H:%M:%SZ", time.gmtime())}
    with open(a.out, "w") as f:
        json.dump(res, f, indent=1)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
