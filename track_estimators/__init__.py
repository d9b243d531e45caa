"""``track_estimators``: the reference's import name, served by the B200-native implementation.

``pip install`` of this repository provides the reference's public surface for the UKF / URTSS path
under its own names, so that code written against NOC-OI/ship-track-estimators runs unchanged on
the GPU::

    from track_estimators.kalman_filters.unscented import UnscentedKalmanFilter
    from track_estimators.kalman_filters.non_linear_process import geodetic_dynamics
    from track_estimators.ship_track import ShipTrack
    from track_estimators.utils import generate_dts, smooth

Every module below IS the corresponding ``ship_track_estimators_b200`` module (one object, two
names).  The Gaussian-process estimator of the reference (``track_estimators.gaussian_processes``)
is outside the scope of this package and is not provided.
"""
import importlib
import sys

import ship_track_estimators_b200 as _impl

__version__ = _impl.__version__

_MODULES = (
    "constants", "utils", "ship_track", "performance_metrics",
    "kalman_filters", "kalman_filters.kalman_filter", "kalman_filters.non_linear_process", "kalman_filters.unscented",
    "cli", "cli.argument_parser", "cli.json_loader", "cli.main_cli",
)
for _name in _MODULES:
    _mod = importlib.import_module(f"ship_track_estimators_b200.{_name}")
    sys.modules[f"{__name__}.{_name}"] = _mod
    if "." not in _name:
        globals()[_name] = _mod
del _name, _mod


def __getattr__(name):
    if name == "gaussian_processes":
        raise ImportError("track_estimators.gaussian_processes is not part of the B200-native package "
                          "(only the UKF / URTSS path of the reference is implemented)")
    raise AttributeError(f"module {__name__!r} has no attribute {name!r}")
