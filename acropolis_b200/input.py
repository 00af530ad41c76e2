"""Cosmological input data: same public surface as the reference's ``acropolis/input.py`` (lines 20-278) --
``locate_sm_file``, ``InputData``, ``InputInterface`` with ``temperature/time/dTdt/hubble_rate/scale_factor/
bbn_abundances/parameter`` -- plus array-valued interpolation for the batched engine and a process-wide cache so
that a scan does not re-read the tarball for every model instance (the reference does, models.py:213).

The scalar code path keeps the reference's arithmetic (log10-log10 linear interpolation between the bracketing
rows, input.py:205-226) operation for operation, because the integer grid sizes NT/NE downstream depend on the
last bit of ``temperature(tau)`` ("bit-level grid conventions")."""
from math import log10
from os import path
import tarfile
from time import monotonic

import numpy as np
from scipy.integrate import cumulative_simpson

from acropolis_b200.params import hbar, NY, NC
from acropolis_b200.pprint import print_error


def locate_data_file(filename):
    pkg_dir, _ = path.split(__file__)
    return path.join(pkg_dir, "data", filename)


_SM_FILE = None


def locate_sm_file():
    global _SM_FILE
    if _SM_FILE is None:
        _SM_FILE = locate_data_file("sm.tar.gz")
    return _SM_FILE


class InputData(object):
    """Raw (cosmo, abundance, param) arrays (reference input.py:71-89)."""

    def __init__(self, cosmo_data, abund_data, param_data):
        self._sCosmoData = cosmo_data
        self._sAbundData = abund_data
        self._sParamData = param_data

    def get_cosmo_data(self):
        return self._sCosmoData

    def get_abund_data(self):
        return self._sAbundData

    def get_param_data(self):
        return self._sParamData


def _read_params(fobj):
    return np.genfromtxt(fobj, delimiter="=", dtype=None, encoding=None)


def input_data_from_dir(dirname):
    return InputData(np.genfromtxt(dirname + "/cosmo_file.dat"), np.genfromtxt(dirname + "/abundance_file.dat"),
                     _read_params(dirname + "/param_file.dat"))


def input_data_from_file(filename):
    with tarfile.open(filename, "r:gz") as tf:
        tc = {m.name: tf.extractfile(m) for m in tf.getmembers()}
        return InputData(np.genfromtxt(tc["cosmo_file.dat"]), np.genfromtxt(tc["abundance_file.dat"]),
                         _read_params(tc["param_file.dat"]))


_II_CACHE = {}
_II_RECENT = {}    # filename -> (InputInterface, time of the last stat): a scan constructs thousands of models per second


def shared_input_interface(filename=None):
    """One InputInterface per input file and process (keyed on path, mtime and size; the file is re-examined at most
    once per second)."""
    filename = locate_sm_file() if filename is None else filename
    now = monotonic()
    hit = _II_RECENT.get(filename)
    if hit is not None and now - hit[1] < 1.0:
        return hit[0]
    st = (path.abspath(filename), path.getmtime(filename), path.getsize(filename))
    ii = _II_CACHE.get(st)
    if ii is None:
        ii = _II_CACHE[st] = InputInterface(filename)
    _II_RECENT[filename] = (ii, now)
    return ii


class InputInterface(object):

    def __init__(self, input_data, type="file"):
        if type == "file":
            if not isinstance(input_data, str):
                raise ValueError("If 'type == file', 'input_data' must be a string")
            input_data = input_data_from_file(input_data)
        elif type == "dir":
            if not isinstance(input_data, str):
                raise ValueError("If 'type == dir', 'input_data' must be be a string")
            input_data = input_data_from_dir(input_data)
        elif type == "raw":
            if not isinstance(input_data, InputData):
                raise ValueError("If 'type == raw', 'input_data' must be an instance of InputData")
        else:
            raise ValueError("Unknown type for input data: Only 'file', 'dir' and 'raw' are supported")

        cosmo = np.asarray(input_data.get_cosmo_data(), dtype=np.float64)
        # scale factor a(t) = exp( int H dt ), cumulative Simpson in log t (reference input.py:125-126, utils.py:72-73)
        t_nat = cosmo[:, 0] / hbar
        sf = np.exp(cumulative_simpson(t_nat * cosmo[:, 4], x=np.log(t_nat), initial=0.))
        self._sCosmoData = np.column_stack([cosmo, sf])
        self._sCosmoDataLog = np.log10(np.abs(self._sCosmoData))
        self._sCosmoDataShp = self._sCosmoData.shape
        abund = np.asarray(input_data.get_abund_data(), dtype=np.float64)
        self._sAbundData = abund.reshape((NY, abund.size // NY))
        pd = input_data.get_param_data()
        self._sParamData = dict(pd.reshape(pd.size))
        self._sTempRange = None
        self._sSearch = {}      # per column: (ascending copy for searchsorted, descending flag)
        self._sMemo = {}        # scalar interpolation results: the points of a grid scan repeat their arguments
        self._check_data()

    def _check_data(self):
        if "eta" not in self._sParamData:
            print_error("The mandatory variable 'eta' could not be found in 'param_file.dat'",
                        "acropolis_b200.input.InputInterface::_check_data")
        if self._sAbundData.shape[0] != NY:
            print_error("The content of 'abundance_file.dat' does not have the required shape.",
                        "acropolis_b200.input.InputInterface::_check_data")
        if self._sCosmoDataShp[1] < NC:
            print_error("The content of 'cosmo_file.dat' does not have the required shape.",
                        "acropolis_b200.input.InputInterface::_check_data")

    # 1. COSMO_DATA ###########################################################

    def _bracket(self, x, v, xc=None):
        """Index ix with v between x[ix] and x[ix+1]; x monotone (either direction); v scalar or array."""
        n = x.shape[0]
        key = xc
        ent = self._sSearch.get(key) if key is not None else None
        if ent is None:
            asc = bool(x[0] < x[1])
            ent = (np.ascontiguousarray(x if asc else x[::-1]), asc)
            if key is not None:
                self._sSearch[key] = ent
        xs, asc = ent
        if asc:
            return np.searchsorted(xs, v, side="right") - 1
        return n - np.searchsorted(xs, v, side="left") - 1

    def _interp_cosmo_data(self, val, xc, yc):
        scalar = isinstance(val, (float, int)) or np.ndim(val) == 0
        if scalar:
            key = (xc, yc, float(val))
            hit = self._sMemo.get(key)
            if hit is not None:
                return hit
        x = self._sCosmoDataLog[:, xc]
        y = self._sCosmoDataLog[:, yc]
        if scalar:
            val_log = log10(val)
            ix = int(self._bracket(x, val_log, xc))
            n = x.shape[0]
            if ix < 0 or ix > n - 2 or not (x[ix] <= val_log <= x[ix + 1] or x[ix] >= val_log >= x[ix + 1]):
                raise ValueError("Interpolation error: Index out of range")
            m = (y[ix + 1] - y[ix]) / (x[ix + 1] - x[ix])
            b = y[ix] - m * x[ix]
            if len(self._sMemo) >= 65536:
                self._sMemo.clear()
            res = self._sMemo[key] = 10.**(m * val_log + b)
            return res
        val_log = np.log10(np.asarray(val, dtype=np.float64))
        ix = self._bracket(x, val_log, xc)
        n = x.shape[0]
        if ix.size and (ix.min() < 0 or ix.max() > n - 2):
            raise ValueError("Interpolation error: Index out of range")
        m = (y[ix + 1] - y[ix]) / (x[ix + 1] - x[ix])
        b = y[ix] - m * x[ix]
        return np.power(10., m * val_log + b)

    def temperature(self, t):
        return self._interp_cosmo_data(t, 0, 1)

    def time(self, T):
        return self._interp_cosmo_data(T, 1, 0)

    def dTdt(self, T):
        return -self._interp_cosmo_data(T, 1, 2)

    def neutrino_temperature(self, T):
        return self._interp_cosmo_data(T, 1, 3)

    def hubble_rate(self, T):
        return self._interp_cosmo_data(T, 1, 4)

    def scale_factor(self, T):
        return self._interp_cosmo_data(T, 1, -1)

    def cosmo_column(self, yc, val, xc=1):
        return self._interp_cosmo_data(val, xc, yc)

    def temperature_range(self):
        if self._sTempRange is None:   # the reference rescans 7374 rows per call (input.py:257-258)
            self._sTempRange = (float(np.min(self._sCosmoData[:, 1])), float(np.max(self._sCosmoData[:, 1])))
        return self._sTempRange

    # device tables for the engine (log10 T and log10 |dT/dt| columns as stored)
    def cosmo_log_tables(self):
        return (np.ascontiguousarray(self._sCosmoDataLog[:, 1]), np.ascontiguousarray(self._sCosmoDataLog[:, 2]))

    # 2. ABUNDANCE_DATA #######################################################

    def bbn_abundances(self):
        return self._sAbundData

    def bbn_abundances_0(self):
        return self._sAbundData[:, 0]

    # 3. PARAM_DATA ###########################################################

    def parameter(self, key):
        return self._sParamData[key]

    def param_keys(self):
        return self._sParamData.keys()
