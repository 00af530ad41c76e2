"""``SpectrumGenerator`` with the reference's interface (acropolis/cascade.py:703-812): ``get_spectrum(E0, S0f, SCf,
T, allX=False)`` and ``rate_photon(E, T)``.  The rate vector G, the kernel planes K and the back-substitution all
run on the GPU (csrc/acro_kernels.cu); this class only evaluates the caller's source callables on the grid.

``get_spectra`` is the batched form (many temperatures in one launch set) that ``NuclearReactor`` uses."""
import numpy as np

from acropolis_b200 import _lib
from acropolis_b200.engine import Engine, PointSpec, energy_grid
import acropolis_b200.params as params


class SpectrumGenerator(object):

    def __init__(self, ii, device=None):
        self._sII = ii
        self._sNX = 3
        self._sEngine = Engine.get(ii, device)

    def _rate_x(self, X, E, T):
        return float(self._sEngine.total_rates([E], [T])[0, X])

    def rate_photon(self, E, T):
        return self._rate_x(0, E, T)

    def _spec(self, E0, S0f, SCf, Ts):
        E_grid = energy_grid(E0)
        S0 = np.array([[float(S0X(T)) for S0X in S0f] for T in Ts])
        SC = np.array([[[float(SCX(E, T)) for E in E_grid] for SCX in SCf] for T in Ts])
        mode = _lib.SC_DENSE if np.any(SC != 0.) else _lib.SC_NONE
        return PointSpec(E0, E_grid, np.asarray(Ts, dtype=np.float64), S0, 0.0, mode, SC if mode else None, None)

    def get_spectra(self, E0, S0f, SCf, Ts):
        """Spectra for several temperatures at once -> float64[NT, 4, NE] (rows: E, F_photon, F_electron, F_positron)."""
        spec = self._spec(E0, S0f, SCf, Ts)
        out = self._sEngine.run([spec], stages=_lib.STAGE_RATES | _lib.STAGE_KERNELS | _lib.STAGE_SOLVE, want=("F",))
        NE = len(spec.E_grid)
        F = out["F"].reshape(len(spec.T), 3, NE)
        sol = np.empty((len(spec.T), 4, NE))
        sol[:, 0, :] = spec.E_grid
        sol[:, 1:, :] = F
        return sol

    def get_spectrum(self, E0, S0f, SCf, T, allX=False):
        sol = self.get_spectra(E0, S0f, SCf, [T])[0]
        return sol[0:2, :] if not allX else sol

    def get_universal_spectrum(self, E0, S0f, SCf, T, offset=0.):
        """The universal photon spectrum of astro-ph/0211258 (reference cascade.py:774-812) -> float64[2, NE] (rows E,
        F_photon): SN K0 (EX/E)^{3/2} / Gamma_gamma below EX = me^2/(80T), SN K0 (EX/E)^2 / Gamma_gamma up to (1+offset) EC
        with EC = me^2/(22T), floored at approx_zero above.  The total photon rates come from the device (closed forms +
        rates.db interpolation / direct integrals); with ``flags.universal`` the batched path applies the same formula inside
        solve_pdi_kernel (offset = 5e-2, nucl.py:419-421)."""
        from math import log
        E_grid = energy_grid(E0)
        EC, EX = params.me2 / (22. * T), params.me2 / (80. * T)
        K0 = E0 / ((EX**2.) * (2. + log(EC / EX)))
        SN = sum(float(S0X(T)) for S0X in S0f)
        rate = self._sEngine.total_rates(E_grid, np.full(len(E_grid), T))[:, 0]
        F = np.zeros(len(E_grid))
        lo = E_grid < EX
        mid = (~lo) & (E_grid <= (1. + offset) * EC)
        F[lo] = SN * K0 * (EX / E_grid[lo])**1.5 / rate[lo]
        F[mid] = SN * K0 * (EX / E_grid[mid])**2.0 / rate[mid]
        F[F < params.approx_zero] = params.approx_zero
        return np.array([E_grid, F])

    def get_kernels(self, E0, T):
        """(E_grid, G[3,NE], K[5,NE,NE]) at one temperature; K planes ordered gg, eg, ge-, ge+, ee, i.e. the reference's
        K[X,Xp] for (X,Xp) = (0,0), (0,1)=(0,2), (1,0), (2,0), (1,1)=(2,2) (cascade.py:751-755)."""
        E_grid = energy_grid(E0)
        spec = PointSpec(E0, E_grid, np.array([T]), np.zeros((1, 3)), 0.0)
        K, G = self._sEngine.kernels_of(spec, 0)
        return E_grid, G, K
