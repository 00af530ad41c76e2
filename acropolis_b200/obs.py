"""Observed light-element abundances the scan results are compared with (reference acropolis/obs.py:1-45): PDG sets by year,
each a dict {"Yp", "DH", "HeD", "LiH"} -> AbundanceObservation(mean, err).  Changes between the sets: the D/H error
(0.035e-5 in 2020, 0.025e-5 in 2021/2022, 0.029e-5 from 2023 on); ``pdg`` is the latest."""


class AbundanceObservation(object):

    def __init__(self, mean, err):
        self.mean = mean
        self.err = err

    def __repr__(self):
        return "AbundanceObservation(mean=%r, err=%r)" % (self.mean, self.err)


def _pdg_set(dh_err):
    return {
        "Yp": AbundanceObservation(2.45e-1, 0.03e-1),
        "DH": AbundanceObservation(2.547e-5, dh_err),
        "HeD": AbundanceObservation(8.3e-1, 1.5e-1),
        "LiH": AbundanceObservation(1.6e-10, 0.3e-10),
    }


pdg2020 = _pdg_set(0.035e-5)
pdg2021 = _pdg_set(0.025e-5)
pdg2022 = _pdg_set(0.025e-5)
pdg2023 = _pdg_set(0.029e-5)
pdg2024 = _pdg_set(0.029e-5)

pdg = pdg2024
