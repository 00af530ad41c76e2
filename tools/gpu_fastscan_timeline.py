#!/usr/bin/env python3
"""Where the time of a fast-parameter scan goes (examples/scan_decay_tau_aa: 200 full points + 39 800 rescale-only points):
host model construction, the full points, the batched transfer (expm) call, row assembly.  Run on the GPU box."""
import cProfile
import pstats
import sys
import os
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np                                  # noqa: E402
import acropolis_b200.flags as flags                # noqa: E402
flags.verbose = False
from acropolis_b200.models import DecayModel        # noqa: E402
from acropolis_b200.scans import BufferedScanner, ScanParameter   # noqa: E402
from acropolis_b200.engine import Engine            # noqa: E402

mk = lambda: BufferedScanner(DecayModel, mphi=ScanParameter(0., 3., 200), tau=1e7, temp0=10.,
                             n0a=ScanParameter(-14., -3., 200, fast=True), bree=0., braa=1.)
mk().perform_scan()
t0 = time.perf_counter()
rows = mk().perform_scan()
print("scan wall %.1f ms, %d rows" % (1e3 * (time.perf_counter() - t0), len(rows)))
pr = cProfile.Profile()
pr.enable()
mk().perform_scan()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
# the transfer call alone
eng = Engine.get(DecayModel(10., 1e5, 10., 1e-10, 0., 1.)._sII)
m = DecayModel(100., 1e7, 10., 1e-10, 0., 1.)
m.run_disintegration()
mp, md = m.get_matp_buffer()
n = 39800
f = np.logspace(0, 11, n)
for rep in range(3):
    t0 = time.perf_counter()
    Y = eng.transfer(np.repeat(mp[None], n, 0), np.repeat(md[None], n, 0), f, m._t_at_tmax())
    print("transfer of %d matrices: %.2f ms (incl. H2D/D2H)" % (n, 1e3 * (time.perf_counter() - t0)))
