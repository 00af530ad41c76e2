"""Consumer of the scan rows (reference acropolis/plots.py): from the ``[N, 2 + 3*NY]`` table that ``BufferedScanner``
returns / ``np.savetxt`` stores, the deviation of Yp, D/H, He3/D and Li7/H from the observed values in units of the combined
error (plots.py:40-101), the 95 % C.L. exclusion masks, and the exclusion plot (plots.py:264-398).

The numbers (``scan_deviations`` and below) are plain numpy and are what tests/ check against the reference's
``_get_deviations`` on reference scan rows; matplotlib is only imported inside the drawing functions, so the module loads
(and the numbers are available) on a machine without it."""
from math import log10, floor, ceil
import warnings

import numpy as np

from acropolis_b200.obs import pdg2022
from acropolis_b200.pprint import print_info
from acropolis_b200.params import NY

# number of standard deviations beyond which a point counts as excluded (95 % C.L., plots.py:35)
_95cl = 1.95996

_plot_number = 0


# numbers ###########################################################################################################

def _get_abundance(data, i):
    """(mean, error estimate) of nuclide i over all rows: the last 3*NY columns are mean | high | low blocks, the error is
    the smaller of |mean - high| and |mean - low| (plots.py:40-52)."""
    first = data.shape[1] - 3 * NY + i
    mean = data[:, first]
    return mean, np.minimum(np.abs(mean - data[:, first + NY]), np.abs(mean - data[:, first + 2 * NY]))


def _sum_in_quadrature(data, i, j):
    mi, ei = _get_abundance(data, i)
    mj, ej = _get_abundance(data, j)
    return mi + mj, np.sqrt(ei**2. + ej**2.)


def _ratio(ma, ea, mb, eb):
    r = ma / mb
    return r, r * np.sqrt((ea / ma)**2. + (eb / mb)**2.)


def _get_deviations(data, obs):
    """(Yp, DH, HeD, LiH) deviations per row (plots.py:55-101).  H = n + p, mass 7 = Li7 + Be7, mass 3 = t + He3 (what is left
    after the late decays), errors added in quadrature and propagated to the ratios.  He3/D only bounds from above: rows where
    D/H or He3/D fall below the observational error are set to +10 (excluded), undefined D/H (no hydrogen) to -10."""
    mH, eH = _sum_in_quadrature(data, 0, 1)
    m7, e7 = _sum_in_quadrature(data, 7, 8)
    m3, e3 = _sum_in_quadrature(data, 3, 4)
    mD, eD = _get_abundance(data, 2)
    mHe4, eHe4 = _get_abundance(data, 5)

    def sigmas(m, e, o):
        return (m - o.mean) / np.sqrt(o.err**2. + e**2.)

    with warnings.catch_warnings(), np.errstate(all="ignore"):
        warnings.simplefilter("ignore", RuntimeWarning)
        mDH, eDH = _ratio(mD, eD, mH, eH)
        mHeD, eHeD = _ratio(m3, e3, mD, eD)
        mLiH, eLiH = _ratio(m7, e7, mH, eH)
        Yp = sigmas(4. * mHe4, 4. * eHe4, obs['Yp'])
        DH = sigmas(mDH, eDH, obs['DH'])
        HeD = sigmas(mHeD, eHeD, obs['HeD'])
        LiH = sigmas(mLiH, eLiH, obs['LiH'])
        HeD[mDH < obs['DH'].err] = 10
        HeD[mHeD < obs['HeD'].err] = 10
    DH[np.isnan(DH)] = -10
    return Yp, DH, HeD, LiH


def _grid_shape(x):
    """(Nx, Ny) of a scan table whose first parameter is the slow one (scans.py row order; plots.py:272-277)."""
    Ny = int((x == x[0]).sum())
    return len(x) // Ny, Ny


def _fix_helium(HeD):
    """Close 'holes' in the He3/D exclusion region (plots.py:293-303): going down a column (fixed second parameter, growing
    first parameter), once a point exceeds the limit every later point below it is set to +10."""
    for col in HeD.T:
        seen = False
        for i, el in enumerate(col):
            if seen and el < _95cl:
                col[i] = 10
            if not seen and el > _95cl:
                seen = True
    return HeD


def scan_deviations(data, obs=pdg2022, fix_helium=False, xc=0, yc=1):
    """The arrays ``plot_scan_results`` draws, without drawing them: dict with x, y, Yp, DH, HeD, LiH reshaped to (Nx, Ny),
    ``overall`` = max(|DH|, |Yp|, HeD) (plots.py:305-307) and ``excluded`` = overall > 1.95996."""
    if isinstance(data, str):
        data = np.loadtxt(data)
    x, y = data[:, xc], data[:, yc]
    shape = _grid_shape(x)
    Yp, DH, HeD, LiH = (a.reshape(shape) for a in _get_deviations(data, obs))
    if fix_helium:
        HeD = _fix_helium(HeD.copy())
    overall = np.maximum(np.maximum(np.abs(DH), np.abs(Yp)), HeD)
    return dict(x=x.reshape(shape), y=y.reshape(shape), Yp=Yp, DH=DH, HeD=HeD, LiH=LiH, overall=overall,
                excluded=overall > _95cl)


# labels ############################################################################################################

_tex_data = {
    # DecayModel
    'mphi': (r'm_\phi', r'\mathrm{MeV}'),
    'tau': (r'\tau_\phi', r'\mathrm{s}'),
    'temp0': (r'T_0', r'\mathrm{MeV}'),
    'n0a': (r'(n_\phi/n_\gamma)|_{T=T_0}', r''),
    # AnnihilationModel
    'braa': (r'\text{BR}_{\gamma\gamma} = 1-\text{BR}_{e^+e^-}', r''),
    'mchi': (r'm_\chi', r'\mathrm{MeV}'),
    'a': (r'a', r'\mathrm{cm^3/s}'),
    'b': (r'b', r'\mathrm{cm^3/s}'),
    'tempkd': (r'T_\text{kd}', r'\mathrm{MeV}'),
}


def add_tex_data(key, tex, unit):
    _tex_data[key] = (tex, unit)


def _value_string(val):
    # floats are printed as powers of ten (the exponent truncated, as in plots.py:139-149), everything else verbatim
    return r'10^' + str(int(log10(val))) if isinstance(val, float) else str(val)


def tex_title(**kwargs):
    """'$m_\\phi=10^1\\,\\mathrm{MeV},\\;...$' from parameter keywords (plots.py:129-167); units are dropped for a value 0."""
    parts = []
    for key, val in kwargs.items():
        tex, unit = _tex_data[key]
        parts.append(tex + '=' + _value_string(val) + ('\\,' + unit if val != 0 else r''))
    return r'$' + r',\;'.join(parts) + r'$'


def tex_label(key):
    if key not in _tex_data:
        return ''
    tex, unit = _tex_data[key]
    return r'$' + tex + (r'\;[' + unit + r']' if unit != r'' else r'') + r'$'


def tex_labels(key_x, key_y):
    return (tex_label(key_x), tex_label(key_y))


# drawing (needs matplotlib) ########################################################################################

def _pyplot():
    try:
        import matplotlib.pyplot as plt
    except ImportError as e:   # absent in the build image: the numbers above do not need it
        raise ImportError("acropolis_b200.plots: drawing needs matplotlib (scan_deviations() does not)") from e
    return plt


_style_set = False


def init_figure():
    global _style_set
    plt = _pyplot()
    if not _style_set:
        plt.rc('text', usetex=True)
        plt.rc('font', family='serif', size=14)
        plt.rcParams['text.latex.preamble'] = r'\usepackage{amsmath}\usepackage{mathpazo}'
        _style_set = True
    fig = plt.figure(figsize=(4.8, 4.4), dpi=150, edgecolor='white')
    ax = fig.add_subplot(1, 1, 1)
    ax.tick_params(axis='both', which='both', labelsize=11, direction='in', width=0.5)
    ax.xaxis.set_ticks_position('both')
    ax.yaxis.set_ticks_position('both')
    for side in ('top', 'bottom', 'left', 'right'):
        ax.spines[side].set_linewidth(0.5)
    return fig, ax


def _decade_ticks(vmin, vmax):
    """Major ticks at the decades that enclose [vmin, vmax] (rounded away from zero in log10), minor ticks at 2..9 times
    each, labels 10^k (plots.py:205-247)."""
    away = lambda v: ceil(v) if v >= 0 else floor(v)
    lo, hi = away(log10(vmin)), away(log10(vmax))
    major = np.linspace(lo, hi, abs(hi - lo) + 1)
    minor = [log10(i * 10**j) for i in range(1, 10) for j in major]
    labels = [r'$10^{' + f'{int(k)}' + '}$' for k in major]
    return lo, hi, major, minor, labels


def _set_tick_labels(ax, x, y):
    from matplotlib.ticker import FixedLocator, FixedFormatter
    for axis, setlim, v in ((ax.xaxis, ax.set_xlim, x), (ax.yaxis, ax.set_ylim, y)):
        lo, hi, major, minor, labels = _decade_ticks(np.min(v), np.max(v))
        axis.set_major_locator(FixedLocator(major))
        axis.set_minor_locator(FixedLocator(minor))
        axis.set_major_formatter(FixedFormatter(labels))
        setlim(lo, hi)


def save_figure(output_file=None, show_fig=False):
    global _plot_number
    plt = _pyplot()
    if output_file is None:
        output_file = 'acropolis_plot_{}.pdf'.format(_plot_number)
        _plot_number += 1
    plt.savefig(output_file)
    if show_fig:
        plt.show()
    print_info("Figure has been saved as '{}'".format(output_file), "acropolis.plot.save_figure")


def plot_scan_results(data, output_file=None, contour_file=None, title='', labels=('', ''), fix_helium=False, show_fig=False,
                      obs=pdg2022, xc=0, yc=1):
    """Exclusion plot of a scan table or file (plots.py:264-398): filled regions for D/H (low grey, high red), Yp (low blue,
    high red) and He3/D (high green, upper limit only), their 95 % C.L. lines and the overall line in black; the overall
    line's vertices go to ``contour_file`` when it is a single path."""
    plt = _pyplot()
    d = scan_deviations(data, obs, fix_helium, xc, yc)
    lx, ly = np.log10(d["x"]), np.log10(d["y"])
    fig, ax = init_figure()
    _set_tick_labels(ax, d["x"], d["y"])
    cut = 1e10
    ax.contourf(lx, ly, d["DH"], levels=[-cut, -_95cl, _95cl, cut], colors=['0.6', 'white', 'tomato'], alpha=0.2)
    ax.contourf(lx, ly, d["Yp"], levels=[-cut, -_95cl, _95cl, cut], colors=['dodgerblue', 'white', 'lightcoral'], alpha=0.2)
    ax.contourf(lx, ly, d["HeD"], levels=[_95cl, cut], colors=['mediumseagreen'], alpha=0.2)
    for arr, level, colour in ((d["DH"], -_95cl, '0.6'), (d["DH"], _95cl, 'tomato'), (d["Yp"], -_95cl, 'dodgerblue'),
                               (d["HeD"], _95cl, 'mediumseagreen')):
        ax.contour(lx, ly, arr, levels=[level], colors=colour, linestyles='-')
    cs = ax.contour(lx, ly, d["overall"], levels=[_95cl], colors='black', linestyles='-')
    ax.set_title(title, fontsize=11)
    ax.set_xlabel(labels[0])
    ax.set_ylabel(labels[1])
    plt.tight_layout()
    if contour_file is not None:
        paths = cs.get_paths() if hasattr(cs, "get_paths") else [p for c in cs.collections for p in c.get_paths()]
        if len(paths) == 1:
            np.savetxt(contour_file, paths[0].vertices)
            print_info("Overall exclusion line has been saved as '{}'".format(contour_file), "acropolis.plot.plot_scan_results")
        else:
            print_info("Could not extract unique contour. No data has been saved.", "acropolis.plot.plot_scan_results")
    if output_file is not None:
        save_figure(output_file)
    if show_fig:
        plt.show()
    return fig, ax
