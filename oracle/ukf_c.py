"""ctypes front end of the plain-C oracle (oracle/ukf_oracle.c).  TEST INFRASTRUCTURE ONLY: the
shipped package never imports this; tests, smoke() and bench.py's CPU baseline do."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_build", "libukf_oracle.so")
LIB_FMA = os.path.join(HERE, "_build", "libukf_oracle_fma.so")   # rounding variant (see Makefile)
LIB_EXT = os.path.join(HERE, "_build", "libukf_oracle_ext.so")   # x87 extended precision (see ukf_oracle.c: REAL)
_lib = None
_lib_fma = None
_lib_ext = None
_d = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")


def load(variant=False):
    """The oracle library; ``variant=True``: the same source built with FMA contraction (a
    rounding-only variant used to measure the oracle's self-uncertainty)."""
    global _lib, _lib_fma, _lib_ext
    if _lib is None:
        src = os.path.join(HERE, "ukf_oracle.c")
        if any(not os.path.exists(p) or os.path.getmtime(p) < os.path.getmtime(src) for p in (LIB, LIB_FMA, LIB_EXT)):
            subprocess.check_call(["make", "-C", HERE, "-s"])
        _lib, _lib_fma, _lib_ext = C.CDLL(LIB), C.CDLL(LIB_FMA), C.CDLL(LIB_EXT)
        for lib in (_lib, _lib_fma):
            lib.oracle_forward.restype = C.c_int
            lib.oracle_smoother.restype = C.c_int
            lib.oracle_batch.restype = C.c_int
        for lib in (_lib, _lib_fma, _lib_ext):
            lib.oracle_track.restype = C.c_int
    if variant == "extended":
        return _lib_ext
    return _lib_fma if variant else _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _c(a, dtype=np.float64):
    return None if a is None else np.ascontiguousarray(a, dtype=dtype)


def run_track(x0, P0, H, Q, R, dt_array, dts, z, sog_rate, cog_rate, smoother=True, noise=None, gating=False, mask=None,
              variant=False):
    """Same contract as oracle.ukf_numpy.run_track; ``noise`` = dict(pred, upd, bwd) of unit normals."""
    from . import ukf_numpy as O

    lib = load(variant)
    dt_array = _c(dt_array)
    n, nobs = len(dt_array), z.shape[1]
    mask = O.update_mask(dt_array, dts) if mask is None else mask
    m8 = _c(mask, np.uint8)
    means = np.zeros((n + 1, 4))
    covs = np.zeros((n + 1, 4, 4))
    n_upd = 1 + int(mask.sum())
    gi = np.zeros(nobs, dtype=np.int32)
    gl = np.ones(nobs)
    noise = noise or {}
    npred, nupd, nbwd = (_c(noise.get(k)) for k in ("pred", "upd", "bwd"))
    zc, sr, cr = _c(z), _c(sog_rate), _c(cog_rate)
    got = lib.oracle_forward(n, nobs, _p(_c(x0)), _p(_c(P0)), _p(_c(H)), _p(_c(Q)), _p(_c(R)), _p(dt_array), _p(m8), _p(zc), _p(sr),
                             _p(cr), _p(npred), _p(nupd), int(gating), C.c_double(50.0), 10000, _p(means), _p(covs), _p(gi), _p(gl))
    if got < 0:
        raise IndexError("update index ran past the observations")
    assert got == n_upd
    out = dict(means=means, covs=covs, mask=np.asarray(mask, dtype=bool), gate_iters=gi[:n_upd].copy(), gate_lambda=gl[:n_upd].copy())
    if smoother:
        ms, cs = means.copy(), covs.copy()
        rep = int((n + 1) / len(dts))
        rc = lib.oracle_smoother(n, len(sr), rep, _p(_c(Q)), _p(dt_array), _p(sr), _p(cr), _p(nbwd), _p(ms), _p(cs))
        if rc < 0:
            raise IndexError("smoother rate index out of range")
        out["means_s"], out["covs_s"] = ms, cs
    return out


def run_track_precision(x0, P0, H, Q, R, dt_array, dts, z, sog_rate, cog_rate, smoother=True, gating=False, mask=None,
                        precision="extended"):
    """One zero-noise track through ``oracle_track`` of the build named by ``precision``: ``"double"``
    (the pinned oracle), ``"fma"`` (its FMA-contracted rounding variant) or ``"extended"`` (the same
    formulas in x87 extended precision, filtered states handed to the smoother unrounded): the
    yardstick that tells, per track, how far an fp64 evaluation is from the exact value."""
    from . import ukf_numpy as O

    lib = load({"double": False, "fma": True, "extended": "extended"}[precision])
    dt_array = _c(dt_array)
    n, nobs = len(dt_array), z.shape[1]
    mask = O.update_mask(dt_array, dts) if mask is None else mask
    m8 = _c(mask, np.uint8)
    out = dict(means=np.zeros((n + 1, 4)), covs=np.zeros((n + 1, 4, 4)), mask=np.asarray(mask, dtype=bool))
    if smoother:
        out.update(means_s=np.zeros((n + 1, 4)), covs_s=np.zeros((n + 1, 4, 4)))
    gi, gl = np.zeros(nobs, dtype=np.int32), np.ones(nobs)
    rep = int((n + 1) / len(dts))
    got = lib.oracle_track(n, nobs, rep, _p(_c(x0)), _p(_c(P0)), _p(_c(H)), _p(_c(Q)), _p(_c(R)), _p(dt_array), _p(m8), _p(_c(z)),
                           _p(_c(sog_rate)), _p(_c(cog_rate)), int(gating), C.c_double(50.0), 10000, _p(out["means"]), _p(out["covs"]),
                           _p(out.get("means_s")), _p(out.get("covs_s")), _p(gi), _p(gl))
    if got < 0:
        raise IndexError("update or smoother rate index ran past the observations")
    out.update(gate_iters=gi[:got].copy(), gate_lambda=gl[:got].copy())
    return out


def run_batch(x0, dt, z_lon, z_lat, sog_rate, cog_rate, H, Q, R, P0, substeps=1, smoother=True, gating=False, z_sog=None, z_cog=None,
              threads=None):
    """Uniform tile in the kernels' [plane][T] layout -> dict of [N+1][4|16][T] arrays (OpenMP over tracks)."""
    lib = load()
    if threads:
        os.environ["OMP_NUM_THREADS"] = str(threads)
    n, T = dt.shape
    nobs = z_lon.shape[0]
    out = {k: np.zeros((n + 1, c, T)) for k, c in (("mean_f", 4), ("cov_f", 16), ("mean_s", 4), ("cov_s", 16))}
    arr = [_c(a) for a in (H, Q, R, P0, x0, dt, z_lon, z_lat)]
    zs, zc_ = _c(z_sog), _c(z_cog)
    sr, cr = _c(sog_rate), _c(cog_rate)
    bad = lib.oracle_batch(T, n, nobs, int(substeps), int(smoother), int(gating), *(_p(a) for a in arr), _p(zs), _p(zc_), _p(sr), _p(cr),
                           *(_p(out[k]) for k in ("mean_f", "cov_f", "mean_s", "cov_s")))
    if bad:
        raise RuntimeError(f"oracle_batch failed with code {bad}")
    return out
