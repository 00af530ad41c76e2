"""Run-time flags, same names as the reference's ``acropolis/flags.py`` (lines 8-41)."""
usedb = True       # interpolate the two tabulated rates from rates.db (device resident)
verbose = True
debug = False
universal = False  # universal photon spectrum instead of the cascade solution (cascade.py:774-812, nucl.py:416-422)


def set_universal(value):
    global universal
    universal = value
