"""acropolis_b200 -- B200-native hot path of ACROPOLIS (hep-mh/acropolis v1.3.1) behind the reference's Python API.

    import acropolis_b200 as acropolis            # or: acropolis_b200.install_as("acropolis")
    from acropolis_b200.models import DecayModel
    Yf = DecayModel(10., 1e5, 10., 1e-10, 0., 1.).run_disintegration()

Submodules mirror the reference: params, flags, input, db, utils, cascade, nucl, models, scans, ext.models, pprint.
The compute path needs the in-tree CUDA library (``__graft_entry__.build()``); there is no CPU fallback."""
import sys as _sys

from acropolis_b200.info import version as __version__  # noqa: F401


def install_as(name="acropolis"):
    """Alias this package (and its submodules) under another top-level name, so that scripts written against the
    reference (``from acropolis.models import DecayModel``) run unchanged."""
    import importlib
    pkg = _sys.modules[__name__]
    _sys.modules[name] = pkg
    for sub in ("params", "flags", "info", "pprint", "input", "db", "utils", "cascade", "nucl", "models", "scans",
                "obs", "plots", "ext", "ext.models", "ext.benchmarks"):
        try:
            _sys.modules[name + "." + sub] = importlib.import_module(__name__ + "." + sub)
        except ImportError:
            pass
    return pkg
