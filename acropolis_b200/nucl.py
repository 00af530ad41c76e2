"""``NuclearReactor`` and ``MatrixGenerator`` with the reference's interface (acropolis/nucl.py:186-512, 515-659).
Reaction tables are restated here for API users; the numbers the engine uses live in csrc/acro_physics.cuh."""
import numpy as np

from acropolis_b200 import _lib
from acropolis_b200.engine import Engine, PointSpec, energy_grid, temperature_grid
from acropolis_b200.params import tau_n, tau_t, NY
import acropolis_b200.params as params

# nuclide identifiers (nucl.py:32-42)
_nuclei = {"n": 0, "p": 1, "d": 2, "t": 3, "He3": 4, "He4": 5, "Li6": 6, "Li7": 7, "Be7": 8}

# photodisintegration reactions and thresholds in MeV (nucl.py:49-88)
_reactions = {1: "d+a>n+p", 2: "t+a>n+d", 3: "t+a>n+p+n", 4: "He3+a>p+d", 5: "He3+a>n+p+p", 6: "He4+a>p+t",
              7: "He4+a>n+He3", 8: "He4+a>d+d", 9: "He4+a>n+p+d", 10: "Li6+a>n+p+He4", 11: "Li6+a>X",
              12: "Li7+a>t+He4", 13: "Li7+a>n+Li6", 14: "Li7+a>n+n+p+He4", 15: "Be7+a>He3+He4", 16: "Be7+a>p+Li6",
              17: "Be7+a>p+p+n+He4"}
_eth = {1: 2.224573, 2: 6.257248, 3: 8.481821, 4: 5.493485, 5: 7.718058, 6: 19.813852, 7: 20.577615, 8: 23.846527,
        9: 26.071100, 10: 3.698892, 11: 15.794685, 12: 2.467032, 13: 7.249962, 14: 10.948850, 15: 1.586627,
        16: 5.605794, 17: 9.304680}
_decays = {1: "n>p", 2: "t>He3"}
_tau = {1: tau_n, 2: tau_t}
_nnuc, _nrec, _ndec = len(_nuclei), len(_reactions), len(_decays)
_lrid, _ldid = list(_reactions.keys()), list(_decays.keys())


class NuclearReactor(object):

    def __init__(self, s0, sc, temp_rg, e0, ii, device=None):
        self._sII = ii
        self._sE0 = e0
        self._sS0 = s0
        self._sSC = sc
        self._sTrg = temp_rg
        self._sEngine = Engine.get(ii, device)

    def _spec(self, Tr):
        E_grid = energy_grid(self._sE0)
        S0 = np.array([[float(f(T)) for f in self._sS0] for T in Tr])
        SC = np.array([[[float(f(E, T)) for E in E_grid] for f in self._sSC] for T in Tr])
        mode = _lib.SC_DENSE if np.any(SC != 0.) else _lib.SC_NONE
        return PointSpec(self._sE0, E_grid, Tr, S0, 0.0, mode, SC if mode else None, None)

    def get_pdi_grids(self):
        """(Tr[NT], {rid: float64[NT]}) as nucl.py:470-512, all temperatures in one batch."""
        Tr = temperature_grid(*self._sTrg)
        out = self._sEngine.run([self._spec(Tr)], stages=_lib.STAGE_RATES | _lib.STAGE_KERNELS | _lib.STAGE_SOLVE,
                                want=("pdi",))
        pdi = out["pdi"]
        return (Tr, {rid: pdi[:, rid - 1].copy() for rid in _lrid})


class MatrixGenerator(object):

    def __init__(self, temp, pdi_grids, ii, device=None):
        self._sII = ii
        self._sTemp = np.asarray(temp, dtype=np.float64)
        self._sPdiGrids = pdi_grids
        (self._sTmin, self._sTmax) = self._sTemp[0], self._sTemp[-1]
        self._sPdi = np.column_stack([np.asarray(pdi_grids[rid], dtype=np.float64) for rid in _lrid])
        self._sEngine = Engine.get(ii, device)

    def get_matp(self, T):
        """(mpdi, mdcy): the rate matrix integrated from Tmax down to T (nucl.py:583-623)."""
        mpdi, mdcy = self._sEngine.matp([self._sTemp], [self._sPdi], T_lo=[T])
        return (mpdi[0], mdcy[0])

    def get_all_matp(self):
        NT = len(self._sTemp)
        mpdi, mdcy = self._sEngine.matp([self._sTemp] * NT, [self._sPdi] * NT, T_lo=self._sTemp)
        return self._sTemp, (mpdi, mdcy)

    def get_final_matp(self):
        return self.get_matp(self._sTmin)
