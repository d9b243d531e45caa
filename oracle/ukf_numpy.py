"""CPU oracle (numpy/scipy restatement) of the reference UKF + URTSS hot path.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference``
legs may import it.  The shipped package (``ship_track_estimators_b200``) never
does: its arithmetic lives in the CUDA library and fails loudly without it.

Parity status: PINNED.  ``tests/golden/make_golden.py`` runs the unmodified
reference (imported from ``/root/reference/src`` with a stub for the missing
``geographiclib`` module) and stores its outputs under ``tests/golden/*.npz``;
``tests/test_oracle_golden.py`` checks every function below against those
fixtures.  The reference's own unit tests pin nothing on this path beyond the
sigma-point / weight identities (reference ``tests/test_unscented_kf.py:24-87``),
which are re-stated in ``tests/test_oracle_golden.py`` as well.

Third-party arithmetic the reference delegates to (not under /root/reference):
``scipy.linalg.sqrtm`` (scipy>=1.11.1 unpinned, 1.18.1 here; call site
``unscented.py:97``), ``numpy.linalg.pinv`` (numpy>=1.25.2 unpinned, 2.3.5 here;
call sites ``unscented.py:243, 333, 422, 475``).  This module calls the same
library routines so that it is as close to bit-faithful as a restatement can be;
``oracle/ukf_oracle.c`` restates them from their published definitions
(principal square root / Moore-Penrose inverse through a symmetric
eigen-decomposition).

All citations ``file:line`` are relative to ``/root/reference/src/track_estimators/``.
State layout: ``x = [lon deg, lat deg, SOG km/h, COG deg]``; time in hours.
"""
from __future__ import annotations

import warnings

import numpy as np
import scipy.linalg

EARTH_RADIUS_KM = 6378.137  # constants.py:1
CHI_ALPHA = 50.0  # kalman_filters/unscented.py:357


# --------------------------------------------------------------------------- #
# noise sources                                                               #
# --------------------------------------------------------------------------- #
class ZeroNoise:
    """Every draw is exactly zero (the deterministic parity mode)."""

    def draw(self, kind, scale):
        return np.zeros(len(scale))


class TapeNoise:
    """Replays pre-drawn *unit* normals, multiplied by ``scale`` like
    ``np.random.normal(scale=..., size=n)`` does (unscented.py:198-200, 232-234,
    320-322).  One queue per call site so the GPU can index them by step:

    ``upd``  (n_updates, n)  initial update first, then one per assimilated obs
    ``pred`` (N, n)          one per predict
    ``bwd``  (N, n)          indexed by backward step number (N-1 ... 0)
    """

    def __init__(self, pred=None, upd=None, bwd=None):
        self.t = {"pred": pred, "upd": upd, "bwd": bwd}
        self.i = {"pred": 0, "upd": 0}
        self.bwd_step = None

    def draw(self, kind, scale):
        tape = self.t[kind]
        if tape is None:
            return np.zeros(len(scale))
        if kind == "bwd":
            return tape[self.bwd_step] * scale
        k = self.i[kind]
        self.i[kind] = k + 1
        return tape[k] * scale


# --------------------------------------------------------------------------- #
# building blocks                                                             #
# --------------------------------------------------------------------------- #
def generate_dts(dts, substeps):
    """utils.py:175-199 - every inter-observation gap split in ``substeps`` equal parts."""
    dts = np.asarray(dts, dtype=np.float64)
    return np.repeat(dts / substeps, substeps)


def ut_weights(n, weight0=None):
    """kalman_filters/unscented.py:109-142 - returns (W0, Wi)."""
    if weight0 is None:
        weight0 = 1 - n / 3.0
    assert -1.0 < weight0 < 1.0
    return weight0, (1 - weight0) / (2 * n)


def sigma_points(x, P, w0):
    """kalman_filters/unscented.py:76-107 - columns X0=x, Xi = x +/- sqrtm(n/(1-W0) P)[:, i].

    A complex root (indefinite P) loses its imaginary part exactly as numpy's
    assignment into a float array does in the reference (:104-105).
    """
    n = x.shape[0]
    root = scipy.linalg.sqrtm((n / (1 - w0)) * P)
    if np.iscomplexobj(root):
        root = root.real
    X = np.empty((n, 2 * n + 1))
    X[:, 0] = x
    for i in range(n):
        X[:, 1 + i] = x + root[:, i]
        X[:, 1 + n + i] = x - root[:, i]
    return X


def geodetic_dynamics(x, dt, sog_rate=0.0, cog_rate=0.0):
    """kalman_filters/non_linear_process.py:6-85 - great-circle step on a sphere."""
    lon = np.radians(x[0])
    lat = np.radians(x[1])
    u = x[2]
    alpha = np.radians(x[3])
    delta = u * dt / EARTH_RADIUS_KM
    sd, cd = np.sin(delta), np.cos(delta)
    east = sd * np.sin(alpha)
    north = np.cos(lat) * cd - np.sin(lat) * sd * np.cos(alpha)
    new_lon = np.degrees(lon + np.arctan2(east, north))
    new_lat = np.degrees(np.arcsin(np.sin(lat) * cd + np.cos(lat) * sd * np.cos(alpha)))
    return np.array([new_lon, new_lat, u + sog_rate * dt, np.degrees(alpha) + cog_rate * dt])


def _weighted_outer(A, B, w0, wi):
    """sum_i W_i A[:, i] B[:, i]^T, i.e. ``A @ diag(W) @ B.T`` (unscented.py:205-207, 324-330)."""
    w = np.full(A.shape[1], wi)
    w[0] = w0
    return (A * w) @ B.T


def predict(x, P, Q, dt, sog_rate, cog_rate, noise):
    """kalman_filters/unscented.py:144-207.  Returns (x, P, X_prior, X_propagated)."""
    n = x.shape[0]
    w0, wi = ut_weights(n)
    X = sigma_points(x, P, w0)
    Y = np.empty_like(X)
    for j in range(X.shape[1]):
        Y[:, j] = geodetic_dynamics(X[:, j], dt, sog_rate, cog_rate)
    w = np.full(X.shape[1], wi)
    w[0] = w0
    mean = (Y * w).sum(axis=1)  # :195
    mean = mean + noise.draw("pred", np.sqrt(np.diag(Q)))  # :198-202
    dev = Y - mean[:, None]  # :205 (about the *noisy* mean)
    return mean, _weighted_outer(dev, dev, w0, wi) + Q, X, Y


def wrap180(a):
    """(a + 180) % 360 - 180 with Python's floored modulo (unscented.py:250, 340)."""
    return (a + 180.0) % 360.0 - 180.0


def update(x, P, H, R, z, noise):
    """kalman_filters/unscented.py:209-265 - linear update, pinv gain, Joseph form."""
    n = x.shape[0]
    z = z + noise.draw("upd", np.sqrt(np.diag(R)))  # :232-236
    S = H @ (P @ H.T) + R  # :240
    K = (P @ H.T) @ np.linalg.pinv(S)  # :243
    y = z - H @ x  # :247
    y[3] = wrap180(y[3])  # :250
    x = x + K @ y  # :254
    x[3] = x[3] % 360.0  # :257
    A = np.eye(n) - K @ H
    P = (A @ P) @ A.T + (K @ R) @ K.T  # :260-265
    return x, P


def criterion_index(x, z, P, H, R):
    """kalman_filters/unscented.py:389-426 - |(z-x)^T pinv(HPH^T+R) (z-x)| (note z - x, no wrap)."""
    y = z - x
    Sinv = np.linalg.pinv(H @ (P @ H.T) + R)
    return float(abs(y @ Sinv @ y))


def check_robustness(x, z, P, H, R, max_iter=10_000):
    """kalman_filters/unscented.py:353-387 with zero measurement noise.

    Returns (R_scaled, iterations, lambda).  ``R`` compounds: R <- R * lambda with the
    running lambda (:380, 510); lambda <- lambda + (gamma - chi) / (y^T S^+ R S^+ y) (:470-481).
    """
    lam = 1.0
    gamma = criterion_index(x, z, P, H, R)
    it = 0
    while gamma > CHI_ALPHA and it < max_iter:
        y = z - x
        Sinv = np.linalg.pinv(H @ (P @ H.T) + R)
        lam = lam + (gamma - CHI_ALPHA) / float(y @ (Sinv @ R @ Sinv) @ y)
        R = R * lam
        gamma = criterion_index(x, z, P, H, R)
        it += 1
    return R, it, lam


def update_mask(dt_array, dts, t0=0.0):
    """kalman_filters/kalman_filter.py:73, 98, 101 - which steps end exactly on an observation
    time.  ``self.time += dt`` is a sequential fp64 sum, as is ``np.cumsum``; membership is exact
    float equality."""
    obs_times = np.cumsum(np.asarray(dts, dtype=np.float64))
    t = np.empty(len(dt_array))
    acc = t0
    for i, d in enumerate(dt_array):
        acc = acc + d
        t[i] = acc
    return np.isin(t, obs_times)


# --------------------------------------------------------------------------- #
# the two hot loops                                                           #
# --------------------------------------------------------------------------- #
def run_forward(x0, P0, H, Q, R, dt_array, dts, z, sog_rate, cog_rate, noise=None, gating=False):
    """kalman_filters/kalman_filter.py:36-117.

    Returns dict(means (N+1, n), covs (N+1, n, n), mask (N,), gate_iters, gate_lambda).
    Index 0 holds the *prior* (x0, P0) stored before the initial update (:76-81).
    """
    noise = noise or ZeroNoise()
    x = np.asarray(x0, dtype=np.float64).reshape(-1).copy()
    P = np.asarray(P0, dtype=np.float64).copy()
    H = np.asarray(H, dtype=np.float64)
    Q = np.asarray(Q, dtype=np.float64)
    R = np.asarray(R, dtype=np.float64)
    dt_array = np.asarray(dt_array, dtype=np.float64)
    mask = update_mask(dt_array, dts)
    means, covs = [x.copy()], [P.copy()]
    gate_iters, gate_lambda = [], []

    def assimilate(x, P, zcol):
        Ruse = R
        if gating:
            Ruse, it, lam = check_robustness(x, zcol, P, H, R)
            gate_iters.append(it)
            gate_lambda.append(lam)
        return update(x, P, H, Ruse, zcol.copy(), noise)

    ui = 0
    x, P = assimilate(x, P, z[:, ui])  # :81
    for step, dt in enumerate(dt_array):
        x, P, _, _ = predict(x, P, Q, dt, sog_rate[ui], cog_rate[ui], noise)  # :90-95
        if mask[step]:  # :101
            ui += 1
            x, P = assimilate(x, P, z[:, ui])
        means.append(x.copy())
        covs.append(P.copy())
    return dict(
        means=np.asarray(means),
        covs=np.asarray(covs),
        mask=mask,
        gate_iters=np.asarray(gate_iters, dtype=np.int32),
        gate_lambda=np.asarray(gate_lambda),
    )


def rts_rate_index(nstates, n_dts):
    """kalman_filters/unscented.py:287-292 - ``np.repeat(rate, int(nstates / len(dts)))``:
    backward step ``s`` reads ``rate[s // rep]``."""
    return int(nstates / n_dts)


def run_smoother(means, covs, Q, dt_array, n_dts, sog_rate, cog_rate, noise=None):
    """kalman_filters/unscented.py:267-351 (URTSS backward pass), quirks kept:
    P_bwd is taken about the filtered mean of ``step`` (:324-325), the gain uses pinv (:333),
    index 0 is the prior, the last state is left untouched."""
    noise = noise or ZeroNoise()
    xs = np.array(means, dtype=np.float64, copy=True)
    Ps = np.array(covs, dtype=np.float64, copy=True)
    nstates, n = xs.shape
    rep = rts_rate_index(nstates, n_dts)
    sog_rep = np.repeat(np.asarray(sog_rate, dtype=np.float64), rep)
    cog_rep = np.repeat(np.asarray(cog_rate, dtype=np.float64), rep)
    w0, wi = ut_weights(n)
    for step in range(nstates - 2, -1, -1):
        X = sigma_points(xs[step], Ps[step], w0)
        Y = np.empty_like(X)
        for j in range(X.shape[1]):
            Y[:, j] = geodetic_dynamics(X[:, j], dt_array[step], sog_rep[step], cog_rep[step])
        w = np.full(X.shape[1], wi)
        w[0] = w0
        x_b = (Y * w).sum(axis=1)
        if isinstance(noise, TapeNoise):
            noise.bwd_step = step
        x_b = x_b + noise.draw("bwd", np.sqrt(np.diag(Q)))  # :320-323
        dev_f = Y - xs[step][:, None]
        P_b = _weighted_outer(dev_f, dev_f, w0, wi) + Q  # :324-325
        D = _weighted_outer(X - xs[step][:, None], Y - x_b[:, None], w0, wi)  # :328-330
        K = D @ np.linalg.pinv(P_b)  # :333
        y = xs[step + 1] - x_b
        y[3] = wrap180(y[3])  # :340
        xs[step] = xs[step] + K @ y
        xs[step][3] = xs[step][3] % 360.0  # :346
        Ps[step] = Ps[step] + (K @ (Ps[step + 1] - P_b)) @ K.T  # :349
    return xs, Ps


def run_track(x0, P0, H, Q, R, dt_array, dts, z, sog_rate, cog_rate, smoother=True, noise=None, gating=False):
    """Forward + (optionally) backward for one track; convenience for tests and the CPU baseline."""
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        out = run_forward(x0, P0, H, Q, R, dt_array, dts, z, sog_rate, cog_rate, noise=noise, gating=gating)
        if smoother:
            xs, Ps = run_smoother(out["means"], out["covs"], Q, dt_array, len(dts), sog_rate, cog_rate, noise=noise)
            out["means_s"], out["covs_s"] = xs, Ps
    return out


# --------------------------------------------------------------------------- #
# fit metrics (performance_metrics.py:4-58) and their per-track form          #
# --------------------------------------------------------------------------- #
def rmse(x, xref):
    """performance_metrics.py:4-20."""
    return np.sqrt(np.mean((np.asarray(x) - np.asarray(xref)) ** 2))


def cum_abs_diff(x, xref):
    """performance_metrics.py:23-39."""
    return np.cumsum(np.abs(np.asarray(x) - np.asarray(xref)))


def abs_diff(x, xref):
    """performance_metrics.py:42-58."""
    return np.abs(np.asarray(x) - np.asarray(xref))


def track_metrics(means, mask, z, rows=(0, 1)):
    """The three metrics of one track's estimate against its assimilated observations: state 0
    with fix 0, then the state after each update (kalman_filter.py:76-116 appends the state after
    the update of a step; ``mask`` is that step's update flag) with the fix it used."""
    states = [0] + [s + 1 for s in range(len(mask)) if mask[s]]
    states = states[: z.shape[1]]
    out = {}
    for r in rows:
        x, xref = np.asarray(means)[states, r], z[r, : len(states)]
        out[r] = dict(rmse=rmse(x, xref), cum_abs=cum_abs_diff(x, xref)[-1], max_abs=abs_diff(x, xref).max(), abs_diff=abs_diff(x, xref))
    return out


# --------------------------------------------------------------------------- #
# WGS84 leg (utils.py:9-72).  The reference takes s12 / azi1 from the third-    #
# party geographiclib (>= 2.0, requirements.txt:4; NOT present in              #
# /root/reference nor installed here).  Restated with Vincenty's published      #
# inverse formulae on the same ellipsoid; PINNED ONLY by the reference's one    #
# exact vector (examples/cli_example/output_01203823_predictions.txt:1) and its #
# loose unit tests (tests/test_utils.py:36-60, 87-100): parity otherwise        #
# unpinned.                                                                     #
# --------------------------------------------------------------------------- #
def wgs84_inverse(lat1, lon1, lat2, lon2):
    """(s12 metres, azi1 degrees) of the WGS84 geodesic from point 1 to point 2."""
    import math

    a, f = 6378137.0, 1.0 / 298.257223563
    b = (1.0 - f) * a
    U1, U2 = math.atan((1 - f) * math.tan(math.radians(lat1))), math.atan((1 - f) * math.tan(math.radians(lat2)))
    L = math.radians(lon2 - lon1)
    su1, cu1, su2, cu2 = math.sin(U1), math.cos(U1), math.sin(U2), math.cos(U2)
    lam = L
    for _ in range(200):
        sl, cl = math.sin(lam), math.cos(lam)
        ss = math.hypot(cu2 * sl, cu1 * su2 - su1 * cu2 * cl)
        cs = su1 * su2 + cu1 * cu2 * cl
        sig = math.atan2(ss, cs)
        sa = cu1 * cu2 * sl / ss
        c2a = 1 - sa * sa
        c2m = cs - 2 * su1 * su2 / c2a if c2a != 0 else 0.0
        C = f / 16 * c2a * (4 + f * (4 - 3 * c2a))
        new = L + (1 - C) * f * sa * (sig + C * ss * (c2m + C * cs * (-1 + 2 * c2m * c2m)))
        conv = abs(new - lam) < 1e-15
        lam = new
        if conv:
            break
    sl, cl = math.sin(lam), math.cos(lam)
    ty, tx = cu2 * sl, cu1 * su2 - su1 * cu2 * cl
    ss, cs = math.hypot(ty, tx), su1 * su2 + cu1 * cu2 * cl
    sig = math.atan2(ss, cs)
    sa = cu1 * cu2 * sl / ss
    c2a = 1 - sa * sa
    c2m = cs - 2 * su1 * su2 / c2a if c2a != 0 else 0.0
    u2 = c2a * (a * a - b * b) / (b * b)
    A = 1 + u2 / 16384 * (4096 + u2 * (-768 + u2 * (320 - 175 * u2)))
    B = u2 / 1024 * (256 + u2 * (-128 + u2 * (74 - 47 * u2)))
    ds = B * ss * (c2m + B / 4 * (cs * (-1 + 2 * c2m * c2m) - B / 6 * c2m * (-3 + 4 * ss * ss) * (-3 + 4 * c2m * c2m)))
    return b * A * (sig - ds), math.degrees(math.atan2(ty, tx))


def wgs84_leg(lon1, lat1, lon2, lat2):
    """(distance km, heading deg in [0, 360)) as geographiclib_distance / geographiclib_heading
    return them, with their same-point guard (utils.py:32-38, 64-70)."""
    if abs(lat1 - lat2) < 1e-8 and abs(lon1 - lon2) < 1e-8:
        return 0.0, 0.0
    s12, azi1 = wgs84_inverse(lat1, lon1, lat2, lon2)
    return s12 * 1e-3, (azi1 + 360) % 360
