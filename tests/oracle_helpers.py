"""Shared helpers of the test-suite: oracle access and golden-fixture -> PointSpec conversion."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))

_ORACLE = {}


def get_oracle(eps=1e-3):
    from pyoracle import Oracle
    import acropolis_b200.flags as flags
    flags.verbose = False
    if eps not in _ORACLE:
        _ORACLE[eps] = Oracle(eps=eps)
    return _ORACLE[eps]


def spec_from_golden(z, t_slice=None):
    from acropolis_b200.engine import PointSpec
    SC = z["SC"]
    T, S0 = z["T"], z["S0"]
    if t_slice is not None:
        T, S0, SC = T[t_slice], S0[t_slice], SC[t_slice]
    mode = 2 if np.any(SC != 0) else 0
    return PointSpec(float(z["E0"]), z["E_grid"], T, S0, float(z["t_at_Tmax"]), mode, SC if mode else None, None)
