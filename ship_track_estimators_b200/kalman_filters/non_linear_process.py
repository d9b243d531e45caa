"""Process model of the UKF: great-circle propagation of ``[lon, lat, SOG, COG]``.

Drop-in for reference ``src/track_estimators/kalman_filters/non_linear_process.py:6-85``.  The
function object doubles as the *model selector* of the CUDA path: ``UnscentedKalmanFilter``
accepts exactly this callable (identity check) and runs the fused device implementation
(``csrc/ste_math.cuh: geodetic_step``).  Called directly it evaluates the same device function
through ``ste_geodetic_f64`` - there is no host implementation.
"""
from __future__ import annotations

import numpy as np


def _device_process(model, n, x, c, dt, sog_rate, cog_rate):
    """One evaluation of a device process model (``ste_process_f64``) for a state or an ``(n, T)`` batch."""
    import torch

    from .. import _native as nat

    if c is not None and np.size(c) != 0:
        raise NotImplementedError("a control vector c is never used by the reference; only c=None is supported")
    xa = np.asarray(x, dtype=np.float64)
    if xa.shape[0] > n:
        raise NotImplementedError(f"this process model is defined for up to {n} state components")
    batched = xa.ndim == 2
    cols = xa if batched else xa.reshape(xa.shape[0], 1)
    T = cols.shape[1]
    full = np.zeros((n, T))
    full[: cols.shape[0]] = cols
    lib = nat.load()
    dev = torch.device("cuda")
    xin = torch.from_numpy(full).to(dev)
    xout = torch.empty_like(xin)

    def vec(v):
        return torch.from_numpy(np.broadcast_to(np.asarray(v, dtype=np.float64), (T,)).copy()).to(dev)

    dtv, srv, crv = vec(dt), vec(sog_rate), vec(cog_rate)
    nat.check(lib.ste_process_f64(model, n, T, T, nat.ptr(xin), nat.ptr(dtv), nat.ptr(srv), nat.ptr(crv), nat.ptr(xout), nat.current_stream()))
    out = xout.cpu().numpy()[: cols.shape[0]]
    return out if batched else out[:, 0]


def geodetic_dynamics(x, c, dt, sog_rate=0.0, cog_rate=0.0):
    """Propagate one state (or a ``(4, T)`` batch of states) by ``dt`` hours on the sphere.

    Parameters follow the reference: ``x = [lon deg, lat deg, SOG km/h, COG deg]``; ``c`` is the
    (unused) control vector, which the reference only ever passes as ``None``
    (``kalman_filter.py:92``, ``unscented.py:308``).
    """
    from .. import _native as nat

    return _device_process(nat.STE_MODEL_GEODETIC, 4, x, c, dt, sog_rate, cog_rate)


def geodetic_dynamics_turn(x, c, dt, sog_rate=0.0, cog_rate=0.0):
    """Five-state variant ``x = [lon, lat, SOG, COG, COG rate]``: the same great-circle step with the
    turn rate taken from the state, where it persists (``cog_rate`` is accepted and ignored, so the
    reference's ``run`` loop, which passes it, works unchanged).  With the reference's default
    weights (``W0 = 1 - n/3`` in (-1, 1), ``unscented.py:125-129``) n = 5 is the largest state its
    class can run; this is the model behind the dimension-generic kernels (``STE_MODEL_GEODETIC_TURN``)."""
    from .. import _native as nat

    return _device_process(nat.STE_MODEL_GEODETIC_TURN, 5, x, c, dt, sog_rate, cog_rate)


#: process callables the CUDA path can run -> (device model id, state dimension)
DEVICE_MODELS = {geodetic_dynamics: (0, 4), geodetic_dynamics_turn: (1, 5)}
