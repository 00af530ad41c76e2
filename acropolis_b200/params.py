"""Physical constants and algorithm parameters, same names and values as the reference's
``acropolis/params.py`` (v1.3.1, lines 15-128).  Unlike the reference -- whose consumer modules import these *by
value*, so that changing them needs a ``sed`` on the file (tools/scan_pd.sh:26-31) -- the engine reads
``NE_pd``, ``NT_pd`` etc. from this module at call time: ``acropolis_b200.params.NE_pd = 150`` takes effect."""
from math import pi

# PHYSICAL CONSTANTS (params.py:15-39)
alpha = 1. / 137.036          # fine-structure constant
me = 0.511                    # electron mass [MeV]
me2 = me**2.
re = alpha / me               # classical electron radius [1/MeV]
GN = 6.70861e-45              # gravitational constant [1/MeV^2]
hbar = 6.582119514e-22        # [MeV s]
c_si = 2.99792458e10          # [cm/s]
tau_n = 8.794e2               # neutron lifetime [s]
tau_t = 5.605e8               # tritium lifetime [s]

# MATHEMATICAL CONSTANTS (params.py:44-49); zeta(3) to double precision (scipy.special.zeta(3.))
zeta3 = 1.2020569031595942
pi2 = pi**2.

# rates.db grid (params.py:54-59)
Emin_log, Tmin_log = 0, -6
Emax_log, Tmax_log = 3, -1
num_pd = 150
Enum = (Emax_log - Emin_log) * num_pd
Tnum = (Tmax_log - Tmin_log) * num_pd

# ALGORITHM-SPECIFIC PARAMETERS (params.py:78-128)
NY = 9                 # nuclides
NC = 5                 # mandatory columns of cosmo_file.dat
Emin = 1.5             # minimum energy of the spectra [MeV]
approx_zero = 1e-200
eps = 1e-3             # the reference's quadrature tolerance; the CUDA path uses fixed-order rules instead
Ephb_T_max = 200.
E_EC_max = 10.
NE_pd = 120            # energy-grid points per decade
NE_min = 10
NT_pd = 20             # temperature-grid points per decade

# B200-path parameters (no reference counterpart)
quad_order = 64        # Gauss-Legendre nodes per thermal kernel integral (16, 32, 64, 96)
pdi_cell_order = 8     # GL nodes per energy-grid cell of the pdi-rate integrals (8, 16)
kernel_mode = 0        # 0: fast evaluation of the thermal kernel integrals (default); 1: plain Gauss-Legendre cross-check
wien_from = 7.38905609893065   # x_a/T from which a thermal integral is done per temperature on the Wien tail (e .. e^2):
                               # e^2 keeps every kernel entry within 1e-7 of the converged integral, e within 1e-9 (1.4x slower)
workspace_bytes = None  # HBM budget [bytes] for the kernel planes of one chunk of points; None = 40% of free HBM (<= 64 GiB)
