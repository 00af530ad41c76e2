"""``LogInterp`` with the reference's semantics (acropolis/utils.py:9-68): log-log linear interpolation on a grid
that is equidistant in log space, ValueError outside the range unless ``fill_value`` is given.  Host-side utility
for users; the engine evaluates the same interpolant on the device (solve_pdi_kernel / matrix_kernel)."""
from math import log

import numpy as np
from scipy.integrate import cumulative_simpson


class LogInterp(object):

    def __init__(self, x_grid, y_grid, base=np.e, fill_value=None):
        self._sLogBase = log(base)
        self._sBase = base
        self._sFillValue = fill_value
        self._sXLog = np.log(x_grid) / self._sLogBase
        self._sYLog = np.log(y_grid) / self._sLogBase
        xdiff = np.diff(self._sXLog)
        if not np.all(xdiff >= 0):
            raise ValueError("The values in x_grid need to be in ascending order.")
        if not np.allclose(xdiff, xdiff[0]):
            raise ValueError("The values in x_grid need to be equidistant in log space.")
        self._sN = len(self._sXLog)

    def __call__(self, x):
        x_log = np.log(x) / self._sLogBase
        lo, hi = self._sXLog[0], self._sXLog[-1]
        inside = (lo <= x_log) & (x_log <= hi)
        if not np.all(inside) and self._sFillValue is None:
            raise ValueError("The given value does not lie within the interpolation range.")
        ix = np.clip(((x_log - lo) * (self._sN - 1) / (hi - lo)).astype(int) if np.ndim(x_log) else
                     int((x_log - lo) * (self._sN - 1) / (hi - lo)) if inside else 0, 0, self._sN - 2)
        m = (self._sYLog[ix + 1] - self._sYLog[ix]) / (self._sXLog[ix + 1] - self._sXLog[ix])
        b = self._sYLog[ix + 1] - m * self._sXLog[ix + 1]
        val = np.power(self._sBase, m * x_log + b)
        return np.where(inside, val, self._sFillValue) if self._sFillValue is not None else val


def cumsimp(x_grid, y_grid):
    return cumulative_simpson(x_grid * y_grid, x=np.log(x_grid), initial=0.)
