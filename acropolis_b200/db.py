"""Rate-table access: same entry points as the reference's ``acropolis/db.py`` (lines 21-54).

The table itself (750 x 450 x 2 log10 rates) lives in ``data/rates_db.npz`` (converted from the reference's
``rates.db.gz`` by tools/convert_reference_data.py) and is uploaded once per device handle; the bilinear
interpolation (db.py:92-124) runs on the device (csrc/acro_physics.cuh: interp_rate_db).  There is no host
interpolation routine on purpose: the product path has no CPU fallback."""
from os import path
from time import time

import numpy as np

from acropolis_b200.pprint import print_info
import acropolis_b200.flags as flags
from acropolis_b200.params import Emin_log, Emax_log, Enum
from acropolis_b200.params import Tmin_log, Tmax_log, Tnum

_DB_CACHE = {}


def import_data_from_db():
    """float64[Tnum*Enum, 2] (row jT*Enum+jE; col 0 photon pair creation, col 1 electron IC) or None."""
    pkg_dir, _ = path.split(__file__)
    db_file = path.join(pkg_dir, "data", "rates_db.npz")
    if not flags.usedb or not path.exists(db_file):
        return None
    if db_file not in _DB_CACHE:
        start_time = time()
        with np.load(db_file) as z:
            tab = np.ascontiguousarray(z["log10_rates"], dtype=np.float64)
        assert tab.shape == (Tnum, Enum, 2), tab.shape
        _DB_CACHE[db_file] = tab.reshape(Tnum * Enum, 2)
        print_info("Read rate database in {:.1f}ms.".format(1e3 * (time() - start_time)),
                   "acropolis_b200.db.import_data_from_db", verbose_level=1)
    return _DB_CACHE[db_file]


def in_rate_db(E_log, T_log):
    return (Emin_log <= E_log <= Emax_log) and (Tmin_log <= T_log <= Tmax_log)
