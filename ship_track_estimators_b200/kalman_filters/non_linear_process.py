"""Process model of the UKF: great-circle propagation of ``[lon, lat, SOG, COG]``.

Drop-in for reference ``src/track_estimators/kalman_filters/non_linear_process.py:6-85``.  The
function object doubles as the *model selector* of the CUDA path: ``UnscentedKalmanFilter``
accepts exactly this callable (identity check) and runs the fused device implementation
(``csrc/ste_math.cuh: geodetic_step``).  Called directly it evaluates the same device function
through ``ste_geodetic_f64`` - there is no host implementation.
"""
from __future__ import annotations

import numpy as np


def geodetic_dynamics(x, c, dt, sog_rate=0.0, cog_rate=0.0):
    """Propagate one state (or a ``(4, T)`` batch of states) by ``dt`` hours on the sphere.

    Parameters follow the reference: ``x = [lon deg, lat deg, SOG km/h, COG deg]``; ``c`` is the
    (unused) control vector, which the reference only ever passes as ``None``
    (``kalman_filter.py:92``, ``unscented.py:308``).
    """
    import torch

    from .. import _native as nat

    if c is not None and np.size(c) != 0:
        raise NotImplementedError("a control vector c is never used by the reference; only c=None is supported")
    xa = np.asarray(x, dtype=np.float64)
    n = xa.shape[0]
    if n > 4:
        raise NotImplementedError("geodetic_dynamics is defined for up to 4 state components")
    batched = xa.ndim == 2
    cols = xa if batched else xa.reshape(n, 1)
    T = cols.shape[1]
    full = np.zeros((4, T))
    full[:n] = cols
    lib = nat.load()
    dev = torch.device("cuda")
    xin = torch.from_numpy(full).to(dev)
    xout = torch.empty_like(xin)

    def vec(v):
        return torch.from_numpy(np.broadcast_to(np.asarray(v, dtype=np.float64), (T,)).copy()).to(dev)

    dtv, srv, crv = vec(dt), vec(sog_rate), vec(cog_rate)
    nat.check(lib.ste_geodetic_f64(T, T, nat.ptr(xin), nat.ptr(dtv), nat.ptr(srv), nat.ptr(crv), nat.ptr(xout), nat.current_stream()))
    out = xout.cpu().numpy()[:n]
    return out if batched else out[:, 0]
