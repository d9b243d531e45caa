"""Named measurement models for ``UnscentedKalmanFilter(measurement_model=...)`` / ``BatchedUKF``.

The reference applies an arbitrary callable to the measurement vector before every update
(``z = self.measurement_model(z)``, ``unscented.py:221-225``; no caller ever sets one).  A Python
callable cannot run inside the batched kernels, so the CUDA path accepts the element-wise models
below: each has a host form (the callable itself, for one ``(n, 1)`` vector) and a batched form
(``apply_rows``: torch operations on the ``[max_obs][T]`` observation rows on the device).
"""
from __future__ import annotations

import numpy as np


def identity(z):
    """``z`` unchanged."""
    return z


def wrap_course(z):
    """Course over ground (row 3) taken modulo 360 degrees; the other rows unchanged."""
    z = np.array(z, dtype=np.float64, copy=True).reshape(-1, 1)
    if z.shape[0] > 3:
        z[3, 0] = z[3, 0] % 360.0
    return z


def position_only(z):
    """Speed and course rows (2 and up) zeroed: a position fix carries no information about them."""
    z = np.array(z, dtype=np.float64, copy=True).reshape(-1, 1)
    z[2:, 0] = 0.0
    return z


MEASUREMENT_MODELS = {"identity": identity, "wrap_course": wrap_course, "position_only": position_only}


def apply_rows(model, rows):
    """Batched form: ``rows`` = the four ``[max_obs][T]`` observation rows (``None`` = absent) on the device."""
    import torch

    if model is None or model is identity:
        return list(rows)
    if model is wrap_course:
        return [rows[0], rows[1], rows[2], None if rows[3] is None else torch.remainder(rows[3], 360.0)]
    if model is position_only:
        return [rows[0], rows[1]] + [None if r is None else torch.zeros_like(r) for r in rows[2:]]
    raise NotImplementedError(f"measurement_model must be one of ship_track_estimators_b200.measurement_models.{sorted(MEASUREMENT_MODELS)}")
