"""Run-time flags, same names as the reference's ``acropolis/flags.py`` (lines 8-41)."""
usedb = True       # interpolate the two tabulated rates from rates.db (device resident)
verbose = True
debug = False
universal = False  # universal-spectrum shortcut (cascade.py:774-812): not on the CUDA path


def set_universal(value):
    global universal
    universal = value
