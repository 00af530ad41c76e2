#!/usr/bin/env python3
"""Convert the reference's INPUT DATA (not code) into the containers this package ships.

* ``acropolis/data/rates.db.gz``  (gzip + numpy-1.x pickle of float64[Tnum*Enum, 2] log10 rates,
  row index jT*Enum + jE, col 0 = photon pair creation, col 1 = electron inverse Compton;
  reference db.py:21-47,95-110)  ->  ``acropolis_b200/data/rates_db.npz`` holding the same numbers as a
  plain float64[Tnum, Enum, 2] array plus the grid bounds from params.py:54-59.
* ``acropolis/data/sm.tar.gz`` (cosmo_file.dat 7374x5, abundance_file.dat 9x3, param_file.dat; reference
  input.py:40-54)  ->  ``acropolis_b200/data/sm.tar.gz``, re-packed with the same three members so that
  ``InputInterface(path)`` keeps accepting the reference's on-disk format, user files included.

Run once in the build container:  python tools/convert_reference_data.py
"""
import gzip
import io
import os
import pickle
import sys
import tarfile

import numpy as np

REF = os.environ.get("ACROPOLIS_REFERENCE", "/root/reference")
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "acropolis_b200", "data")


def main():
    os.makedirs(DST, exist_ok=True)
    with gzip.open(os.path.join(REF, "acropolis", "data", "rates.db.gz"), "rb") as f:
        raw = np.asarray(pickle.load(f), dtype=np.float64)
    Emin_log, Emax_log, Tmin_log, Tmax_log, num_pd = 0, 3, -6, -1, 150
    Enum, Tnum = (Emax_log - Emin_log) * num_pd, (Tmax_log - Tmin_log) * num_pd
    assert raw.shape == (Tnum * Enum, 2), raw.shape
    np.savez_compressed(os.path.join(DST, "rates_db.npz"), log10_rates=raw.reshape(Tnum, Enum, 2),
                        Emin_log=Emin_log, Emax_log=Emax_log, Tmin_log=Tmin_log, Tmax_log=Tmax_log)
    src = tarfile.open(os.path.join(REF, "acropolis", "data", "sm.tar.gz"), "r:gz")
    with tarfile.open(os.path.join(DST, "sm.tar.gz"), "w:gz") as dst:
        for name in ("cosmo_file.dat", "abundance_file.dat", "param_file.dat"):
            data = src.extractfile(name).read()
            ti = tarfile.TarInfo(name)
            ti.size = len(data)
            dst.addfile(ti, io.BytesIO(data))
    print("wrote", os.listdir(DST))


if __name__ == "__main__":
    sys.exit(main())
