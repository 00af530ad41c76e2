"""Model API: same classes, constructor signatures and hooks as the reference's ``acropolis/models.py``
(``AbstractModel`` :24-206, ``DecayModel`` :209-287, ``AnnihilationModel`` :290-389).

What differs is underneath ``_pdi_matrix()`` (reference :112-130): instead of NuclearReactor -> SpectrumGenerator ->
QUADPACK per temperature, the model describes itself as one ``PointSpec`` (grids + source arrays) and the batched
CUDA engine computes spectra, photodisintegration rates, the T-integral and the matrix exponential.  The source-term
hooks stay host Python; the built-in models evaluate them vectorised over the temperature grid, user subclasses
that override scalar hooks are honoured through a (slow, host-side) per-call loop that produces the same arrays."""
from abc import ABC, abstractmethod
from math import pi, exp, log, log10

import numpy as np

from acropolis_b200.input import InputInterface, locate_sm_file, shared_input_interface
from acropolis_b200.params import zeta3
from acropolis_b200.params import hbar, c_si, me2, alpha, tau_t
from acropolis_b200.params import NY
import acropolis_b200.params as params
import acropolis_b200.flags as flags
from acropolis_b200.pprint import print_info, print_warning
from acropolis_b200 import _lib
from acropolis_b200.engine import Engine, PointSpec, energy_grid, temperature_grid


def _fsr_shape(E, EX):
    """E-dependence of the final-state-radiation photon source (reference models.py:276-287 / :378-389):
    (1/EX)(alpha/pi)(1+(1-x)^2)/x log((1-x)/y) for x = E/EX <= 1-y, y = me^2/(4 EX^2); else 0."""
    E = np.asarray(E, dtype=np.float64)
    x = E / EX
    y = me2 / (4. * EX**2.)
    ok = ~(1. - y < x)
    xs = np.where(ok, x, 0.5)
    return np.where(ok, (1. / EX) * (alpha / pi) * (1. + (1. - xs)**2.) / xs * np.log((1. - xs) / y), 0.)


_FSR_MEMO = {}


def _fsr_columns(E_grid, EX):
    """[NE,3] energy shape of the final-state-radiation source (photon column only); shared by the points of a scan."""
    key = (float(EX), len(E_grid), float(E_grid[0]))
    v = _FSR_MEMO.get(key)
    if v is None:
        if len(_FSR_MEMO) >= 8192:
            _FSR_MEMO.clear()
        ze = np.zeros(len(E_grid))
        v = _FSR_MEMO[key] = np.column_stack([_fsr_shape(E_grid, EX), ze, ze])
        v.setflags(write=False)
    return v


def _fold_neutron(mat):
    """mat with the free-neutron decay folded out: column n := 0, row p += row n, row n := 0 (see _neutron_foldable)."""
    m = np.array(mat, dtype=np.float64)
    m[:, 0] = 0.
    m[1, :] += m[0, :]
    m[0, :] = 0.
    return m


def _neutron_foldable(model, mpdi, mdcy):
    """True if exp(mpdi+mdcy) may be replaced by exp of the folded matrix inside post-decay x expm x pre-decay without
    changing the product: neutrons are no target (their column is pure decay into protons, the proton column is empty),
    the pre-decay empties the neutron entry and the post-decay adds the remaining neutrons to the protons -- so nothing
    depends on the neutron lifetime, while its rate x time (up to 1e9) is what forces ~29 squarings in expm and costs 8
    digits of nucleon-number conservation.  Same fold as apply_transfer() in csrc/acro_kernels.cu, so that
    run_disintegration() and the batched scan path give the same abundances."""
    if not _hooks_unchanged(model, AbstractModel, ("_pred_matrix", "_postd_matrix")):
        return False
    a = np.asarray(mpdi) + np.asarray(mdcy)
    return bool(a[0, 0] + a[1, 0] == 0. and not a[2:, 0].any() and not a[:, 1].any())


class AbstractModel(ABC):

    def __init__(self, e0, ii):
        self._sII = ii
        self._sE0 = e0
        self._sTrg = self._temperature_range()
        (self._sS0, self._sSc) = self.get_source_terms()
        self._sMatpBuffer = None     # (mpdi, mdcy), reused across the fast scan parameter

    # --------------------------------------------------------------------------------------------------
    def _warn_ranges(self):
        if not flags.universal and int(self._sE0) > 1e3:
            print_warning("Injection energy > 1 GeV. Results cannot be trusted.",
                          "acropolis_b200.models.AbstractMode.run_disintegration")
        cf = self._sII.temperature_range()
        if not (cf[0] <= self._sTrg[0] <= self._sTrg[1] <= cf[1]):
            print_warning("Temperature range not covered by input data. Results cannot be trusted.",
                          "acropolis_b200.models.AbstractMode.run_disintegration")

    def is_trivial(self):
        """E0 below every threshold: no computation (reference models.py:65-71)."""
        return self._sE0 <= params.Emin

    def run_disintegration(self):
        self._warn_ranges()
        if self.is_trivial():
            print_info("Injection energy is below all thresholds. No calculation required.",
                       "acropolis_b200.models.AbstractModel.run_disintegration", verbose_level=1)
            return self._squeeze_decays(self._sII.bbn_abundances())
        pred_mat = self._pred_matrix()
        pdi_mat = self._pdi_matrix()
        postd_mat = self._postd_matrix()
        transf_mat = postd_mat.dot(pdi_mat.dot(pred_mat))
        return np.column_stack(list(transf_mat.dot(Y0i) for Y0i in self._sII.bbn_abundances().transpose()))

    def get_source_terms(self):
        s0 = [self._source_photon_0, self._source_electron_0, self._source_positron_0]
        sc = [self._source_photon_c, self._source_electron_c, self._source_positron_c]
        return (s0, sc)

    # --------------------------------------------------------------------------------------------------
    def _engine(self):
        return Engine.get(self._sII)

    def _t_at_tmax(self):
        return self._sII.time(max(self._sTrg))

    def _source_arrays(self, T, E_grid):
        """(S0[NT,3], sc_mode, sc, sc_E) from the scalar hooks (cascade.py:763-764).  Generic slow path."""
        S0 = np.array([[float(f(Ti)) for f in self._sS0] for Ti in T])
        cls = type(self)
        default_c = (cls._source_photon_c is AbstractModel._source_photon_c and
                     cls._source_electron_c is AbstractModel._source_electron_c and
                     cls._source_positron_c is AbstractModel._source_positron_c)
        if default_c:
            return S0, _lib.SC_NONE, None, None
        SC = np.array([[[float(f(E, Ti)) for E in E_grid] for f in self._sSc] for Ti in T])
        return S0, _lib.SC_DENSE, SC, None

    def point_spec(self):
        """The model as one row of the batched engine's input."""
        E_grid = energy_grid(self._sE0)
        T = temperature_grid(*self._sTrg)
        S0, mode, sc, sc_E = self._source_arrays(T, E_grid)
        return PointSpec(self._sE0, E_grid, T, S0, self._t_at_tmax(), mode, sc, sc_E)

    @classmethod
    def point_specs(cls, models):
        """PointSpecs of many models of this class.  Built-in models vectorise the source arrays over the whole group
        (one table interpolation for all temperatures of all models); the default is one by one."""
        return [m.point_spec() for m in models]

    def _pdi_matrix(self):
        eng = self._engine()
        if self._sMatpBuffer is None:
            out = eng.run([self.point_spec()], want=("mpdi", "mdcy", "status"))
            if int(out["status"][0]) & _lib.ST_NONFINITE:    # range warnings were printed by _warn_ranges()
                print_warning("Non-finite abundances for this parameter point.",
                              "acropolis_b200.models.AbstractModel._pdi_matrix")
            self._sMatpBuffer = (out["mpdi"][0], out["mdcy"][0])
        mpdi, mdcy = self._sMatpBuffer
        if _neutron_foldable(self, mpdi, mdcy):
            mpdi, mdcy = _fold_neutron(mpdi), _fold_neutron(mdcy)
        _, ex = eng.transfer(mpdi, mdcy, 1.0, self._t_at_tmax(), want_expm=True)
        return ex[0]

    def _pred_matrix(self):
        dmat = np.identity(NY)
        dmat[0, 0], dmat[1, 0] = 0., 1.                     # n > p
        expf = exp(-self._t_at_tmax() / tau_t)              # t > He3
        dmat[3, 3], dmat[4, 3] = expf, 1. - expf
        return dmat

    def _postd_matrix(self):
        dmat = np.identity(NY)
        dmat[0, 0], dmat[1, 0] = 0., 1.   # n   > p
        dmat[3, 3], dmat[4, 3] = 0., 1.   # t   > He3
        dmat[8, 8], dmat[7, 8] = 0., 1.   # Be7 > Li7
        return dmat

    def _squeeze_decays(self, Yf):
        dmat = self._postd_matrix()
        return np.column_stack(list(dmat.dot(Yi) for Yi in Yf.transpose()))

    def get_matp_buffer(self):
        return self._sMatpBuffer

    def set_matp_buffer(self, matp):
        self._sMatpBuffer = matp

    # ABSTRACT METHODS ##############################################

    @abstractmethod
    def _temperature_range(self):
        pass

    @abstractmethod
    def _source_photon_0(self, T):
        pass

    @abstractmethod
    def _source_electron_0(self, T):
        pass

    def _source_positron_0(self, T):
        return self._source_electron_0(T)

    def _source_photon_c(self, E, T):
        return 0.

    def _source_electron_c(self, E, T):
        return 0.

    def _source_positron_c(self, E, T):
        return self._source_electron_c(E, T)


def _hooks_unchanged(obj, base, names):
    cls = type(obj)
    return all(getattr(cls, n) is getattr(base, n) for n in names)


_HOOKS = ("_source_photon_0", "_source_electron_0", "_source_positron_0",
          "_source_photon_c", "_source_electron_c", "_source_positron_c")


class DecayModel(AbstractModel):

    def __init__(self, mphi, tau, temp0, n0a, bree, braa):
        self._sII = shared_input_interface(locate_sm_file())
        self._sMphi = mphi            # mass of the decaying particle [MeV]
        self._sTau = tau              # lifetime [s]
        self._sE0 = self._sMphi / 2.  # injection energy [MeV]
        self._sN0a = n0a              # number density relative to photons at T = temp0
        self._sT0 = temp0
        self._st0 = self._sII.time(self._sT0)
        self._sBRee = bree
        self._sBRaa = braa
        super(DecayModel, self).__init__(self._sE0, self._sII)

    def _number_density(self, T):
        sf_ratio = self._sII.scale_factor(self._sT0) / self._sII.scale_factor(T)
        delta_t = self._sII.time(T) - self._st0
        n_gamma = (2. * zeta3) * (self._sT0**3.) / (pi**2.)
        return self._sN0a * n_gamma * sf_ratio**3. * np.exp(-delta_t / self._sTau)

    def _temperature_range(self):
        mag = 2.
        Td = self._sII.temperature(self._sTau)
        Td_ofm = log10(Td)
        Tmin = 10.**(Td_ofm - 3. * mag / 4.)
        Tmax = 10.**(Td_ofm + 1. * mag / 4.)
        return (Tmin, Tmax)

    def _source_photon_0(self, T):
        return self._sBRaa * 2. * self._number_density(T) * (hbar / self._sTau)

    def _source_electron_0(self, T):
        return self._sBRee * self._number_density(T) * (hbar / self._sTau)

    def _source_photon_c(self, E, T):
        EX = self._sE0
        x = E / EX
        y = me2 / (4. * EX**2.)
        if 1. - y < x:
            return 0.
        _sp = self._source_electron_0(T)
        return (_sp / EX) * (alpha / pi) * (1. + (1. - x)**2.) / x * log((1. - x) / y)

    def _source_arrays(self, T, E_grid):
        if not _hooks_unchanged(self, DecayModel, _HOOKS + ("_number_density",)):
            return super(DecayModel, self)._source_arrays(T, E_grid)
        nd = self._number_density(T) * (hbar / self._sTau)
        S0 = np.column_stack([self._sBRaa * 2. * nd, self._sBRee * nd, self._sBRee * nd])
        if self._sBRee == 0.:
            return S0, _lib.SC_NONE, None, None
        sc = np.column_stack([self._sBRee * nd, np.zeros_like(nd), np.zeros_like(nd)])
        sc_E = np.column_stack([_fsr_shape(E_grid, self._sE0), np.zeros(len(E_grid)), np.zeros(len(E_grid))])
        return S0, _lib.SC_SEPARABLE, sc, sc_E


    @classmethod
    def point_specs(cls, models):
        if not models or any(type(m) is not DecayModel for m in models) or len({id(m._sII) for m in models}) != 1:
            return [m.point_spec() for m in models]
        ii = models[0]._sII
        Ts = [temperature_grid(*m._sTrg) for m in models]
        nt = np.array([len(t) for t in Ts])
        Tall = np.concatenate(Ts)
        rep = lambda vals: np.repeat(np.array(vals, dtype=np.float64), nt)
        sf0 = rep([ii.scale_factor(m._sT0) for m in models]) if len({m._sT0 for m in models}) > 1 else \
            float(ii.scale_factor(models[0]._sT0))
        sf_ratio = sf0 / ii.scale_factor(Tall)
        delta_t = ii.time(Tall) - rep([m._st0 for m in models])
        n_gamma = (2. * zeta3) * (rep([m._sT0 for m in models])**3.) / (pi**2.)
        tau = rep([m._sTau for m in models])
        nd_all = rep([m._sN0a for m in models]) * n_gamma * sf_ratio**3. * np.exp(-delta_t / tau) * (hbar / tau)
        t_tmax = [ii.time(max(m._sTrg)) for m in models]      # scalar path (memoised): same last bit as point_spec()
        # [sum NT, 3] source arrays of all points at once (same products as _source_arrays), sliced per point below
        braa2, bree = rep([m._sBRaa * 2. for m in models]), rep([m._sBRee for m in models])
        S0_all = np.empty((len(nd_all), 3))
        S0_all[:, 0] = braa2 * nd_all
        S0_all[:, 1] = S0_all[:, 2] = bree * nd_all
        sc_all = np.zeros((len(nd_all), 3))
        sc_all[:, 0] = S0_all[:, 1]
        specs, off = [], 0
        for k, m in enumerate(models):
            sl = slice(off, off + nt[k])
            off += nt[k]
            E_grid = energy_grid(m._sE0)
            if m._sBRee == 0.:
                specs.append(PointSpec.trusted(float(m._sE0), E_grid, Ts[k], S0_all[sl], float(t_tmax[k])))
            else:
                specs.append(PointSpec.trusted(float(m._sE0), E_grid, Ts[k], S0_all[sl], float(t_tmax[k]), _lib.SC_SEPARABLE,
                                               sc_all[sl], _fsr_columns(E_grid, m._sE0)))
        return specs


class AnnihilationModel(AbstractModel):

    def __init__(self, mchi, a, b, tempkd, bree, braa, omegah2=0.12):
        self._sII = shared_input_interface(locate_sm_file())
        self._sMchi = mchi          # dark-matter mass [MeV]
        self._sSwave = a            # s-wave part of <sigma v> [cm^3/s]
        self._sPwave = b            # p-wave part [cm^3/s]
        self._sTkd = tempkd         # kinetic decoupling temperature [MeV]
        self._sE0 = self._sMchi     # injection energy [MeV]
        self._sBRee = bree
        self._sBRaa = braa
        self._sOmgh2 = omegah2
        super(AnnihilationModel, self).__init__(self._sE0, self._sII)

    def _number_density(self, T):
        rho_d0 = 8.095894680377574e-35 * self._sOmgh2   # DM density today [MeV^4]
        T0 = 2.72548 * 8.6173324e-11                    # CMB temperature today [MeV]
        Tmin = self._sII.temperature_range()[0] * (1 + 1e-6)
        sf0 = self._sII.scale_factor(Tmin) * Tmin / T0
        sf_ratio = sf0 / self._sII.scale_factor(T)
        return rho_d0 * sf_ratio**3. / self._sMchi

    def _dm_temperature(self, T):
        if np.ndim(T) == 0:
            if T >= self._sTkd:
                return T
            sf_ratio = self._sII.scale_factor(self._sTkd) / self._sII.scale_factor(T)
            return self._sTkd * sf_ratio**2.
        T = np.asarray(T, dtype=np.float64)
        below = T < self._sTkd
        if not np.any(below):
            return T
        sf_ratio = self._sII.scale_factor(self._sTkd) / self._sII.scale_factor(T)
        return np.where(below, self._sTkd * sf_ratio**2., T)

    def _sigma_v(self, T):
        swave_nu = self._sSwave / ((hbar**2.) * (c_si**3.))
        pwave_nu = self._sPwave / ((hbar**2.) * (c_si**3.))
        v2 = 6. * self._dm_temperature(T) / self._sMchi
        return swave_nu + pwave_nu * v2

    def _temperature_range(self):
        mag = 4.
        Tmax = me2 / (22. * .5 * params.Emin)
        Tmin = 10.**(log10(Tmax) - mag)
        return (Tmin, Tmax)

    def _source_photon_0(self, T):
        return self._sBRaa * (self._number_density(T)**2.) * self._sigma_v(T)

    def _source_electron_0(self, T):
        return self._sBRee * .5 * (self._number_density(T)**2.) * self._sigma_v(T)

    def _source_photon_c(self, E, T):
        EX = self._sE0
        x = E / EX
        y = me2 / (4. * EX**2.)
        if 1. - y < x:
            return 0.
        _sp = self._source_electron_0(T)
        return (_sp / EX) * (alpha / pi) * (1. + (1. - x)**2.) / x * log((1. - x) / y)

    def _source_arrays(self, T, E_grid):
        if not _hooks_unchanged(self, AnnihilationModel, _HOOKS + ("_number_density",)):
            return super(AnnihilationModel, self)._source_arrays(T, E_grid)
        base = (self._number_density(T)**2.) * self._sigma_v_array(T)
        S0 = np.column_stack([self._sBRaa * base, self._sBRee * .5 * base, self._sBRee * .5 * base])
        if self._sBRee == 0.:
            return S0, _lib.SC_NONE, None, None
        sc = np.column_stack([self._sBRee * .5 * base, np.zeros_like(base), np.zeros_like(base)])
        sc_E = np.column_stack([_fsr_shape(E_grid, self._sE0), np.zeros(len(E_grid)), np.zeros(len(E_grid))])
        return S0, _lib.SC_SEPARABLE, sc, sc_E

    @classmethod
    def point_specs(cls, models):
        """Group form for plain AnnihilationModels on one temperature grid (all of them have Tmax = me^2/(11 Emin) and four
        decades): the table interpolations are done once for the group, the rest is a handful of 80-element array
        operations per model, in the order of _number_density / _dm_temperature / _sigma_v so that the arrays are bit-equal
        to the per-model ones (tests/test_host.py)."""
        if (not models or any(type(m) is not AnnihilationModel for m in models) or len({id(m._sII) for m in models}) != 1
                or len({m._sTrg for m in models}) != 1):
            return [m.point_spec() for m in models]
        ii = models[0]._sII
        T = temperature_grid(*models[0]._sTrg)
        sfT = ii.scale_factor(T)
        T0 = 2.72548 * 8.6173324e-11
        Tmin = ii.temperature_range()[0] * (1 + 1e-6)
        sf0 = ii.scale_factor(Tmin) * Tmin / T0
        cube = (sf0 / sfT)**3.
        t_tmax = float(ii.time(max(models[0]._sTrg)))
        nu = (hbar**2.) * (c_si**3.)
        z = np.zeros_like(T)
        specs = []
        for m in models:
            nd = 8.095894680377574e-35 * m._sOmgh2 * cube / m._sMchi
            below = T < m._sTkd
            Tdm = np.where(below, m._sTkd * (ii.scale_factor(m._sTkd) / sfT)**2., T) if np.any(below) else T
            base = (nd**2.) * (m._sSwave / nu + m._sPwave / nu * (6. * Tdm / m._sMchi))
            E_grid = energy_grid(m._sE0)
            half = m._sBRee * .5 * base
            S0 = np.empty((len(T), 3))
            S0[:, 0] = m._sBRaa * base
            S0[:, 1] = S0[:, 2] = half
            if m._sBRee == 0.:
                specs.append(PointSpec.trusted(float(m._sE0), E_grid, T, S0, t_tmax))
            else:
                sc = np.column_stack([half, z, z])
                specs.append(PointSpec.trusted(float(m._sE0), E_grid, T, S0, t_tmax, _lib.SC_SEPARABLE, sc,
                                               _fsr_columns(E_grid, m._sE0)))
        return specs

    def _sigma_v_array(self, T):
        """<sigma v> on an array of temperatures; subclasses with a scalar-only _sigma_v are looped."""
        try:
            v = self._sigma_v(T)
            v = np.asarray(v, dtype=np.float64)
            if v.shape == np.shape(T):
                return v
            if v.ndim == 0:
                return np.full(np.shape(T), float(v))
        except (TypeError, ValueError):
            pass
        return np.array([float(self._sigma_v(float(Ti))) for Ti in T])
