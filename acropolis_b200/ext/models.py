"""``ResonanceModel`` and ``estimate_tempkd_ee`` with the reference's interface (acropolis/ext/models.py:30-325).

The model only changes the annihilation cross-section ``_sigma_v(T)`` (Breit-Wigner thermal average) and the kinetic
decoupling temperature; everything below ``_pdi_matrix`` is the shared CUDA path.  The reference evaluates the thermal
average with three adaptive ``quad`` calls PER temperature and per source-term call (~NE+20 times per temperature,
SURVEY.md 3.4); here it is one vectorised fixed-order evaluation for the whole temperature grid:

* around the resonance (|u - u_R| <= eps u_R) the substitution u = u_R + w tan(theta) absorbs the Lorentzian, GL-64 in theta;
* below / above it the integration variable is s = log|log(u/u_R)| (the Lorentzian tail ~ 1/log(u/u_R)^2 becomes a
  power law), GL-96 each,

with the reference's limits (u_max = 200 m^2/x, the window edges u_R (1 -+ eps), eps = params.eps)."""
from math import sqrt, exp, log, pi

import numpy as np
from numpy.polynomial.legendre import leggauss
from scipy.optimize import brentq
from scipy.special import kn

from acropolis_b200.models import AnnihilationModel
from acropolis_b200.params import me, pi2, zeta3
import acropolis_b200.params as params
from acropolis_b200.pprint import print_error

_GL = {n: leggauss(n) for n in (64, 96, 128)}


def _gl(f, a, b, n):
    """Gauss-Legendre of a vectorised f over [a, b]; a, b may be arrays (broadcast along the last axis of the nodes)."""
    x, w = _GL[n]
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    h, m = 0.5 * (b - a), 0.5 * (b + a)
    return h * np.sum(w * f(m[..., None] + h[..., None] * x), axis=-1)


def estimate_tempkd_ee(mchi, delta, gammad, gammav, nd, S, ii, sigma_ee):
    """Kinetic-decoupling temperature from chi e -> chi e scattering: root of log(<sigma v> n_e / (H N_col)) in log T
    (reference ext/models.py:30-133; the reference uses scipy root(lm) from log(mchi))."""
    fac = 200.
    (Tmin, Tmax) = ii.temperature_range()
    Tmin *= (1 + 1e-6)
    Tmax *= (1 - 1e-6)

    def _k1(y):
        return sqrt(pi / 2) * (y**0.5 + (3. / 8.) * y**1.5 - (15. / 128.) * y**2.5 + (105. / 1024.) * y**3.5)

    def _k2(y):
        return sqrt(pi / 2) * (y**0.5 + (15. / 8.) * y**1.5 + (105. / 128.) * y**2.5 - (315. / 1024.) * y**3.5)

    def _kernel(logz, T):
        z = np.exp(logz)
        s = (z + me + mchi)**2.
        sqrt_s = np.sqrt(s)
        sigma = sigma_ee(s=s, mchi=mchi, delta=delta, gammad=gammad, gammav=gammav)
        xmax = 400
        ye, ychi, ys = T / me, T / mchi, T / sqrt_s
        if me / T < xmax < mchi / T:
            bessel = np.exp(-(z + me) / T) * _k1(ys) / _k2(ychi) / kn(2, me / T)
        elif me / T > xmax:
            bessel = np.exp(-z / T) * _k1(ys) / (_k2(ychi) * _k2(ye))
        else:
            bessel = kn(1, sqrt_s / T) / (kn(2, me / T) * kn(2, mchi / T))
        return z * (2. * sqrt_s) * sigma * (s - (me + mchi)**2.) * (s - (me - mchi)**2.) * bessel / sqrt_s

    def _sigma_v_ee(T):
        return _gl(lambda lz: _kernel(lz, T), log(T / fac), log(fac * T), 128) / (8. * me**2. * mchi**2. * T)

    def _nee(T):
        xe = me / T
        # int_xe^fac y sqrt(y^2 - xe^2)/(e^y + 1) dy with y = xe + t^2 (removes the square-root end point)
        tmax = sqrt(max(fac - xe, 0.))
        f = _gl(lambda t: (xe + t * t) * np.sqrt(np.maximum((xe + t * t)**2. - xe**2., 0.)) / (np.exp(xe + t * t) + 1.) * 2. * t,
                0., tmax, 128) if xe < fac else 0.
        nee_1 = 4. * f * (T**3.) / (2. * pi2)
        Y = ii.bbn_abundances_0()
        nee_2 = (Y[1] + 2. * Y[5]) * ii.parameter('eta') * (2. * zeta3 / pi2) * (T**3.)
        return max(nee_1, nee_2)

    def _root(logT):
        T = float(np.clip(exp(logT), Tmin, Tmax))
        Ncol = max(1., mchi / T)
        return log(_sigma_v_ee(T) * _nee(T) / ii.hubble_rate(T) / Ncol)

    def _root_grid(logT):
        """_root for an array of log T: the same expressions with T as a column and the Bessel-function regimes of _kernel as
        masks.  Only used to LOCATE the sign changes (400 scalar calls with a 128-node integral each were 23 of the 25 ms of a
        ResonanceModel constructor); the root itself is refined with the scalar function below."""
        T = np.clip(np.exp(logT), Tmin, Tmax)
        out = np.empty(len(T))
        xe_all, xchi_all = me / T, mchi / T
        regimes = (xe_all < 400) & (400 < xchi_all), xe_all > 400
        regimes = regimes + (~(regimes[0] | regimes[1]),)
        for which, mask in enumerate(regimes):
            if not np.any(mask):
                continue
            Tc = T[mask][:, None]

            def kern(lz):
                z = np.exp(lz)
                s = (z + me + mchi)**2.
                sqrt_s = np.sqrt(s)
                sigma = sigma_ee(s=s, mchi=mchi, delta=delta, gammad=gammad, gammav=gammav)
                ye, ychi, ys = Tc / me, Tc / mchi, Tc / sqrt_s
                if which == 0:
                    bessel = np.exp(-(z + me) / Tc) * _k1(ys) / _k2(ychi) / kn(2, me / Tc)
                elif which == 1:
                    bessel = np.exp(-z / Tc) * _k1(ys) / (_k2(ychi) * _k2(ye))
                else:
                    bessel = kn(1, sqrt_s / Tc) / (kn(2, me / Tc) * kn(2, mchi / Tc))
                return z * (2. * sqrt_s) * sigma * (s - (me + mchi)**2.) * (s - (me - mchi)**2.) * bessel / sqrt_s

            Tm = T[mask]
            sv = _gl(kern, np.log(Tm / fac), np.log(fac * Tm), 128) / (8. * me**2. * mchi**2. * Tm)
            xe = (me / Tm)[:, None]
            tmax = np.sqrt(np.maximum(fac - me / Tm, 0.))
            with np.errstate(over="ignore"):      # rows with me/T >= fac are discarded by the where() below
                f = _gl(lambda t: (xe + t * t) * np.sqrt(np.maximum((xe + t * t)**2. - xe**2., 0.)) / (np.exp(xe + t * t) + 1.) * 2. * t,
                        np.zeros_like(tmax), tmax, 128)
            nee_1 = 4. * np.where(me / Tm < fac, f, 0.) * (Tm**3.) / (2. * pi2)
            Y = ii.bbn_abundances_0()
            nee_2 = (Y[1] + 2. * Y[5]) * ii.parameter('eta') * (2. * zeta3 / pi2) * (Tm**3.)
            with np.errstate(divide="ignore", invalid="ignore"):
                out[mask] = np.log(sv * np.maximum(nee_1, nee_2) / ii.hubble_rate(Tm) / np.maximum(1., mchi / Tm))
        return out

    x0 = log(mchi)
    grid = np.linspace(log(Tmin), log(Tmax), 400)
    vals = _root_grid(grid)
    sc = np.where(np.sign(vals[:-1]) * np.sign(vals[1:]) < 0)[0]
    tempkd = np.nan
    if len(sc):
        k = sc[np.argmin(np.abs(grid[sc] - x0))]
        tempkd = exp(brentq(_root, grid[k], grid[k + 1], xtol=1e-13, rtol=1e-13))
    return np.nanmin([tempkd, mchi / 20.])


class ResonanceModel(AnnihilationModel):

    def __init__(self, mchi, delta, gammad, gammav, nd, tempkd, S=1, omegah2=0.12):
        super(ResonanceModel, self).__init__(mchi, None, None, None, 1, 0, omegah2)
        if callable(tempkd):
            self._sTkd = tempkd(mchi=mchi, delta=delta, gammad=gammad, gammav=gammav, nd=nd, S=S, ii=self._sII)
        else:
            self._sTkd = tempkd
        self._sDelta = delta
        self._sGammad = gammad
        self._sGammav = gammav
        self._sNd = nd
        self._sNv = 1
        self._sS = S
        self._sMR = self._sMchi * (2. + self._sDelta)
        self._sPR = self._sMchi * sqrt(self._sDelta)
        self._sWidth = self._decay_width(self._sPR)
        self._check_input_parameters()

    def _check_nwa(self, eps=0.1):
        return (self._sWidth / self._sMR) < eps and self._sNd != 0

    def _check_input_parameters(self):
        if self._sDelta > 1:
            print_error("The mass splitting must be < 1. The calculation cannot be trusted.",
                        "acropolis_b200.ext.models.ResonanceModel._check_input_parameters")
        if self._sGammad >= 1 or self._sGammav >= 1:
            print_error("The couplings must be small (< 1). The calculation cannot be trusted.",
                        "acropolis_b200.ext.models.ResonanceModel._check_input_parameters")
        if self._sNd not in [0, 1]:
            print_error("Currently only s-wave annihilations with 'nd = 0' and p-wave annihilations with 'nd = 1' are supported.",
                        "acropolis_b200.ext.models.ResonanceModel._check_input_parameters")

    def _decay_width_d(self, p):
        return self._sGammad * self._sMR * (p / self._sMchi)**(2. * self._sNd + 1.)

    def _decay_width_v(self, p):
        return self._sGammav * self._sMR * (p / self._sMchi)**(2. * self._sNv + 1.)

    def _decay_width(self, p):
        return self._decay_width_d(p) + self._decay_width_v(p)

    def _sigma_v_res(self, T):
        x = self._sMchi / T
        return 8. * self._sS * (pi * x)**1.5 * self._sGammad * self._sGammav * self._sMR**2. * self._sDelta**(self._sNd + .5) \
            * np.exp(-self._sDelta * x) / self._sWidth / self._sMchi**3.

    def _sigma_v_non_res(self, T):
        x = self._sMchi / T
        gamma = {0: sqrt(pi) / 2., 1: 3. * sqrt(pi) / 4.}[self._sNd]
        return 4. * self._sS * sqrt(pi) * x**(-self._sNd) * self._sGammad * self._sGammav * self._sMR**2. * gamma \
            / (self._sMchi**4. * self._sDelta**2.)

    def _sigma_v_full(self, T):
        """Thermally averaged Breit-Wigner cross-section (reference ext/models.py:272-318), scalar or array T."""
        scalar = np.ndim(T) == 0
        T = np.atleast_1d(np.asarray(T, dtype=np.float64))
        x = self._sMchi / T
        uR, m2 = self._sPR**2., self._sMchi**2.
        eps = params.eps
        pref = 4. * x * np.sqrt(x * pi) * self._sS * self._sGammad * self._sGammav * (self._sMR / self._sMchi)**2.
        xk = x[:, None]

        def kern(u):     # integrand in du (the reference integrates kern*u in d log u)
            width = self._decay_width(np.sqrt(u))
            return np.exp(-u * xk / m2) * (u / m2)**(self._sNd + .5) / ((u - uR)**2. + (self._sMchi * width / 2.)**2.)

        umax = 200. * m2 / x
        uR_l = np.minimum(uR * (1. - eps), umax)
        uR_h = uR * (1. + eps)
        on = uR_l < umax                                  # the resonance lies inside the thermal range
        # below the resonance: t = log(u_R/u) in [log(u_R/uR_l), 60], s = log t
        t_lo = np.log(uR / uR_l)

        def below(s):
            t = np.exp(s)
            u = uR * np.exp(-t)
            return kern(u) * u * t
        I1 = _gl(below, np.log(t_lo), np.full_like(t_lo, log(60.)), 96)
        # around the resonance: u = u_R + w tan(theta)
        w = 0.5 * self._sMchi * self._sWidth
        th_l, th_h = np.arctan((uR_l - uR) / w), np.arctan((uR_h - uR) / w)

        def around(th):
            u = uR + w * np.tan(th)
            return kern(u) * w / np.cos(th)**2.
        I2 = np.where(on, _gl(around, th_l, np.full_like(th_l, th_h), 64), 0.)
        # above the resonance: t = log(u/u_R) in [log(1+eps), log(umax/u_R)]
        t_hi = np.log(np.maximum(umax / uR, 1. + 2. * eps))

        def above(s):
            t = np.exp(s)
            u = uR * np.exp(t)
            return kern(u) * u * t
        I3 = np.where(on, _gl(above, np.full_like(t_hi, log(np.log(1. + eps))), np.log(t_hi), 96), 0.)
        out = pref * (I1 + I2 + I3)
        return float(out[0]) if scalar else out

    def _sigma_v(self, T):
        return self._sigma_v_full(self._dm_temperature(T))
