"""The three benchmark resonance models of arXiv:2409.14900 with the reference's names (acropolis/ext/benchmarks.py:13-52):
chi e -> chi e cross-sections for the kinetic-decoupling estimate and ``BenchmarkModel1..3``."""
from functools import partial

from acropolis_b200.params import me2, pi
from acropolis_b200.ext.models import ResonanceModel, estimate_tempkd_ee


def _pref(s, mchi, delta, gammad, gammav):
    return 16. * pi * gammad * gammav / mchi**4. / (2. + delta)**4. / s**2.


def sigma_ee_b1(s, mchi, delta, gammad, gammav):
    mx2 = mchi**2.
    return _pref(s, mchi, delta, gammad, gammav) * mx2 * (s * s - 2. * s * (mx2 - 3. * me2) + (me2 - mx2)**2.)


def sigma_ee_b2(s, mchi, delta, gammad, gammav):
    mx2, mx4, me4 = mchi**2., mchi**4., me2 * me2
    return .25 * _pref(s, mchi, delta, gammad, gammav) * (4. * (s**3.) - 10. * (s**2.) * (mx2 + me2)
                                                         + s * (9. * mx4 + 22. * mx2 * me2 + 9. * me4)
                                                         - 4. * (mx4 - me4) * (mx2 - me2) + (mx2 - me2)**4. / s)


def sigma_ee_b3(s, mchi, delta, gammad, gammav):
    mx2, mx4, me4 = mchi**2., mchi**4., me2 * me2
    return 4.5 * _pref(s, mchi, delta, gammad, gammav) * (mx4 * (me2 + s) - 2. * mx2 * (me2 - s)**2. + (me4 - s**2.) * (me2 - s))


BenchmarkModel1 = partial(ResonanceModel, nd=0., S=1. / 2., tempkd=partial(estimate_tempkd_ee, sigma_ee=sigma_ee_b1))
BenchmarkModel2 = partial(ResonanceModel, nd=0., S=3. / 4., tempkd=partial(estimate_tempkd_ee, sigma_ee=sigma_ee_b2))
BenchmarkModel3 = partial(ResonanceModel, nd=1., S=3. / 2., tempkd=partial(estimate_tempkd_ee, sigma_ee=sigma_ee_b3))
