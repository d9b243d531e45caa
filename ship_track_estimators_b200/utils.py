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


WGS84_A = 6378137.0                 # semi-major axis, m
WGS84_F = 1.0 / 298.257223563       # flattening


def vincenty_inverse(lat1: float, lon1: float, lat2: float, lon2: float, tol: float = 1e-15, max_iter: int = 200):
    """WGS84 inverse geodesic by Vincenty's formulae: ``(s12 in metres, azi1 in degrees (-180, 180])``.

    Used when ``geographiclib`` is not installed.  It agrees with geographiclib to ~1e-12 relative on
    ship-scale legs (the reference's one exact vector: 8e-13 in distance, 2e-11 degrees in azimuth);
    nearly antipodal points, where the iteration does not converge, raise ``ValueError``."""
    import math

    b = (1.0 - WGS84_F) * WGS84_A
    U1 = math.atan((1.0 - WGS84_F) * math.tan(math.radians(lat1)))
    U2 = math.atan((1.0 - WGS84_F) * math.tan(math.radians(lat2)))
    L = math.radians(lon2 - lon1)
    sU1, cU1, sU2, cU2 = math.sin(U1), math.cos(U1), math.sin(U2), math.cos(U2)

    def at(lam):
        sl, cl = math.sin(lam), math.cos(lam)
        ty, tx = cU2 * sl, cU1 * sU2 - sU1 * cU2 * cl
        ss, cs = math.hypot(ty, tx), sU1 * sU2 + cU1 * cU2 * cl
        sa = cU1 * cU2 * sl / ss
        c2a = 1.0 - sa * sa
        c2sm = cs - 2.0 * sU1 * sU2 / c2a if c2a != 0.0 else 0.0
        return ty, tx, ss, cs, math.atan2(ss, cs), sa, c2a, c2sm

    lam = L
    for _ in range(max_iter):
        ty, tx, ss, cs, sig, sa, c2a, c2sm = at(lam)
        C = WGS84_F / 16.0 * c2a * (4.0 + WGS84_F * (4.0 - 3.0 * c2a))
        nxt = L + (1.0 - C) * WGS84_F * sa * (sig + C * ss * (c2sm + C * cs * (-1.0 + 2.0 * c2sm * c2sm)))
        done = abs(nxt - lam) < tol
        lam = nxt
        if done:
            break
    else:
        raise ValueError("Vincenty's inverse iteration did not converge (nearly antipodal points)")
    ty, tx, ss, cs, sig, sa, c2a, c2sm = at(lam)
    u2 = c2a * (WGS84_A * WGS84_A - b * b) / (b * b)
    A = 1.0 + u2 / 16384.0 * (4096.0 + u2 * (-768.0 + u2 * (320.0 - 175.0 * u2)))
    B = u2 / 1024.0 * (256.0 + u2 * (-128.0 + u2 * (74.0 - 47.0 * u2)))
    dsig = B * ss * (c2sm + B / 4.0 * (cs * (-1.0 + 2.0 * c2sm * c2sm)
                                      - B / 6.0 * c2sm * (-3.0 + 4.0 * ss * ss) * (-3.0 + 4.0 * c2sm * c2sm)))
    return b * A * (sig - dsig), math.degrees(math.atan2(ty, tx))


def _wgs84_inverse(lat1, lon1, lat2, lon2):
    """WGS84 inverse geodesic: the third-party ``geographiclib`` when it is installed (the reference's
    own route, utils.py:36, 68), this module's Vincenty restatement otherwise."""
    try:
        from geographiclib.geodesic import Geodesic
    except ImportError:
        s12, azi1 = vincenty_inverse(float(lat1), float(lon1), float(lat2), float(lon2))
        return {"s12": s12, "azi1": azi1}
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
