"""Runs the UNMODIFIED reference (NOC-OI/ship-track-estimators) on in-memory tracks.

TEST / BENCH INFRASTRUCTURE ONLY (``bench.py --impl reference`` and its ``cpu_baseline`` leg): the
shipped package never imports this module.  The reference is the pip-installed copy under
``baseline/_ref`` (``python -m pip install --no-index --no-build-isolation --no-deps --target
baseline/_ref <copy of /root/reference>``; git-ignored, it travels to the GPU box with the working
tree) or, in the build container, ``/root/reference/src``.  Its one missing import,
``geographiclib``, is satisfied by ``oracle/stubs`` (never called: tracks are built from arrays).

What is patched, and why, exactly as SURVEY.md section 8(c) prescribes:
  * ``np.random.normal`` returns zeros while a track runs (the reference adds unseeded noise,
    ``unscented.py:198-202, 232-236, 320-323``; the workload is defined with zero noise);
  * for the gated configuration only, a subclass re-enables the robustification call the reference
    leaves commented out (``unscented.py:228``) and silences its ``print``.
"""
from __future__ import annotations

import builtins
import os
import sys
from types import SimpleNamespace

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
CANDIDATES = (os.path.join(REPO, "baseline", "_ref"), "/root/reference/src")

_ref = None


def locate():
    """Directory holding an importable ``track_estimators`` package, or None."""
    for path in CANDIDATES:
        if os.path.isdir(os.path.join(path, "track_estimators", "kalman_filters")):
            return path
    return None


def load():
    """-> namespace(UnscentedKalmanFilter, GatedUKF, geodetic_dynamics, generate_dts, path); raises ImportError."""
    global _ref
    if _ref is not None:
        return _ref
    path = locate()
    if path is None:
        raise ImportError("the reference is not installed (baseline/_ref) and /root/reference/src does not exist")
    for p in (os.path.join(HERE, "stubs"), path):
        if p not in sys.path:
            sys.path.insert(0, p)
    from track_estimators.kalman_filters.non_linear_process import geodetic_dynamics
    from track_estimators.kalman_filters.unscented import UnscentedKalmanFilter
    from track_estimators.utils import generate_dts

    class GatedUKF(UnscentedKalmanFilter):
        """The reference with its robustification line re-enabled (unscented.py:228-229)."""

        def update(self, z):
            z = z.reshape(-1, 1)
            keep_print, keep_R = builtins.print, self.R
            builtins.print = lambda *a, **k: None
            try:
                self.R = self.check_robustness(z, self.P, self.R)
            finally:
                builtins.print = keep_print
            try:
                super().update(z)
            finally:
                self.R = keep_R

    _ref = SimpleNamespace(UnscentedKalmanFilter=UnscentedKalmanFilter, GatedUKF=GatedUKF, geodetic_dynamics=geodetic_dynamics,
                           generate_dts=generate_dts, path=path)
    return _ref


def run_track(z, dts, substeps, H, Q, R, P, sog_rate, cog_rate, smoother=True, gating=False):
    """One track through the reference's own ``run`` / ``run_rts_smoother`` with zero noise.
    ``z`` (4, nobs) rows lon, lat, sog, cog; ``dts`` (nobs-1,) hours.  -> dict of means/covs[, _s]."""
    ref = load()
    st = SimpleNamespace(dts=np.asarray(dts, dtype=float), z=np.asarray(z, dtype=float), sog_rate=np.array(sog_rate, dtype=float),
                         cog_rate=np.array(cog_rate, dtype=float), sog=np.asarray(z[2], dtype=float), cog=np.asarray(z[3], dtype=float))
    dt_array = np.asarray(ref.generate_dts(st.dts, substeps), dtype=float)
    cls = ref.GatedUKF if gating else ref.UnscentedKalmanFilter
    ukf = cls(H=np.array(H, dtype=float), Q=np.array(Q, dtype=float), R=np.array(R, dtype=float), P=np.array(P, dtype=float),
              x0=st.z[:, 0].reshape(-1, 1).copy(), non_linear_process=ref.geodetic_dynamics)
    orig = np.random.normal
    np.random.normal = lambda loc=0.0, scale=1.0, size=None: np.zeros(size)
    try:
        means, covs = ukf.run(len(dt_array), dt_array, st)
        out = dict(means=np.asarray(means).reshape(len(dt_array) + 1, -1), covs=np.asarray(covs).reshape(len(dt_array) + 1, 4, 4))
        if smoother:
            ms, cs = ukf.run_rts_smoother(st)
            out.update(means_s=np.asarray(ms).reshape(len(dt_array) + 1, -1), covs_s=np.asarray(cs).reshape(len(dt_array) + 1, 4, 4))
    finally:
        np.random.normal = orig
    return out
