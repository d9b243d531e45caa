"""Host-side helpers that shape the inputs of the UKF/URTSS hot path.

Drop-in names for reference ``src/track_estimators/utils.py``.  Nothing here runs on the
GPU: these are O(nobs) preprocessing steps (SURVEY.md section 8, row a10 / "next" row N1).
All functions accept scalars or numpy arrays (the reference is scalar-only).
"""
from __future__ import annotations

from typing import List, Union

import numpy as np

from .constants import EARTH_RADIUS

_SAME_POINT_TOL = 1e-8  # reference utils.py:32, 64


def _wgs84_inverse(lat1, lon1, lat2, lon2):
    """WGS84 inverse geodesic through the optional third-party ``geographiclib``."""
    try:
        from geographiclib.geodesic import Geodesic
    except ImportError as exc:  # pragma: no cover - depends on the environment
        raise ImportError(
            "geographiclib is not installed; build ShipTrack with "
            "calc_distance_func=haversine_formula, calc_heading_func=heading instead"
        ) from exc
    return Geodesic.WGS84.Inverse(lat1, lon1, lat2, lon2)


def geographiclib_distance(lon1: float, lat1: float, lon2: float, lat2: float) -> float:
    """Ellipsoidal distance in km (reference utils.py:9-39)."""
    if abs(lat1 - lat2) < _SAME_POINT_TOL and abs(lon1 - lon2) < _SAME_POINT_TOL:
        return 0.0
    return _wgs84_inverse(lat1, lon1, lat2, lon2)["s12"] * 1e-3


def geographiclib_heading(lon1: float, lat1: float, lon2: float, lat2: float) -> float:
    """Initial azimuth in [0, 360) degrees (reference utils.py:42-72)."""
    if abs(lat1 - lat2) < _SAME_POINT_TOL and abs(lon1 - lon2) < _SAME_POINT_TOL:
        return 0.0
    return (_wgs84_inverse(lat1, lon1, lat2, lon2)["azi1"] + 360) % 360


def haversine_formula(lon1, lat1, lon2, lat2):
    """Great-circle distance in km on the sphere of radius EARTH_RADIUS (reference utils.py:75-113).

    The reference evaluates the ``atan2`` form of the central angle; so does this.
    """
    lam1, phi1, lam2, phi2 = (np.radians(v) for v in (lon1, lat1, lon2, lat2))
    a = np.sin((phi2 - phi1) / 2.0) ** 2 + np.cos(phi1) * np.cos(phi2) * np.sin((lam2 - lam1) / 2.0) ** 2
    return 2 * np.arctan2(np.sqrt(a), np.sqrt(1 - a)) * EARTH_RADIUS


def heading(lon1, lat1, lon2, lat2):
    """Initial great-circle bearing in [0, 360) degrees (reference utils.py:116-147)."""
    lam1, phi1, lam2, phi2 = (np.radians(v) for v in (lon1, lat1, lon2, lat2))
    dlam = lam2 - lam1
    east = np.sin(dlam) * np.cos(phi2)
    north = np.cos(phi1) * np.sin(phi2) - np.sin(phi1) * np.cos(phi2) * np.cos(dlam)
    return (np.degrees(np.arctan2(east, north)) + 360) % 360


def smooth(y: np.ndarray, box_pts: int) -> np.ndarray:
    """Centred moving average of width ``box_pts`` with zero padding (reference utils.py:150-172)."""
    return np.convolve(y, np.full(box_pts, 1.0 / box_pts), mode="same")


def generate_dts(dts: Union[np.ndarray, List[Union[int, float]]], substeps: int) -> np.ndarray:
    """Split every inter-observation gap into ``substeps`` equal steps (reference utils.py:175-199).

    ``dts[j] / substeps`` repeated ``substeps`` times; the filter's step grid.
    """
    dts = np.asarray(dts, dtype=np.float64)
    if substeps <= 0 or dts.size == 0:
        return np.asarray([], dtype=np.float64)
    return np.repeat(dts / substeps, substeps)
