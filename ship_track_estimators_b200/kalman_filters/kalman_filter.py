"""Kalman filter base class: the time loop and the smoother driver of the reference API.

Drop-in for reference ``src/track_estimators/kalman_filters/kalman_filter.py``.  The loops
themselves (``run`` :36-117, ``run_rts_smoother`` :119-137) execute on the GPU as one kernel launch
each (``ste_ukf_forward_f64`` / ``ste_urtss_backward_f64``); this class keeps the reference's
attributes (``means``, ``covariances``, ``time``, ``dt``, ``nsteps``, ``c``...) and return shapes.
"""
from __future__ import annotations

from typing import List, Tuple, Union

import numpy as np


class KalmanFilterBase:
    """State containers shared by the filters (reference ``kalman_filter.py:20-34``)."""

    def __init__(self, *args, **kwargs):
        self.time = 0
        self.c = None  # control input vector; computed for parity, never used (kalman_filter.py:92)
        self.means: list = []
        self.covariances: list = []
        self.means_smoothed: list = []
        self.covariances_smoothed: list = []
        self.dt = None
        self.nsteps = None

    # The concrete filter supplies the batched device implementation of both loops.
    def _run_device(self, nsteps, dt, ship_track):
        raise NotImplementedError("Predict not implemented.")

    def run(
        self,
        nsteps: int,
        dt: Union[int, float, List[Union[int, float]], np.ndarray],
        ship_track,
        *args,
        **kwargs,
    ) -> Tuple[np.ndarray, np.ndarray]:
        """Filter ``ship_track`` over ``nsteps`` steps of length ``dt`` (scalar or per-step array).

        Returns ``(means (N+1, n), covariances (N+1, n, n))``; row 0 is the prior stored before the
        initial update (reference ``kalman_filter.py:76-81``).  As in the reference, results
        accumulate in ``self.means`` / ``self.covariances`` across calls.
        """
        if isinstance(dt, (list, np.ndarray)):
            assert len(dt) == nsteps, "dt must be the same length as nsteps"
        else:
            dt = np.ones(nsteps) * dt
        self.dt = dt
        self.nsteps = nsteps
        self._run_device(nsteps, np.asarray(dt, dtype=np.float64), ship_track)
        return np.asarray(self.means).squeeze(), np.asarray(self.covariances).squeeze()

    def run_rts_smoother(self, ship_track) -> Tuple[np.ndarray, np.ndarray]:
        """Rauch-Tung-Striebel smoothing of the stored filtered states (reference ``:119-137``)."""
        x, P = self.rts_step(np.asarray(self.means), np.asarray(self.covariances), ship_track)
        return x.squeeze(), P.squeeze()

    def predict(self, *args, **kwargs):
        """Predict the state."""
        raise NotImplementedError("Predict not implemented.")

    def update(self, *args, **kwargs):
        """Update the state."""
        raise NotImplementedError("Update not implemented.")
