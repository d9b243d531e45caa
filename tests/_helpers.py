"""Shared helpers for the parity tests: fixture loading and the tolerance definitions.

Tolerances (BASELINE.json north_star "1e-9 relative"):
  means        |d| <= tol * max(1, |ref|) per component, COG compared on the circle
  covariances  max|dP| <= tol * max|P_ref| per matrix (norm-relative; entries cross zero)
  decisions    (update mask, gating iteration counts) exact
"""
from __future__ import annotations

import os
import re
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(REPO, "tests", "golden")
if REPO not in sys.path:
    sys.path.insert(0, REPO)

TOL = 1e-9
# Where the reference's own output moves by more than TOL under a rounding-only change of its
# LAPACK calls (fixture key ``unc``, see tests/golden/make_golden.py: rounding_variant), parity is
# asserted at that self-uncertainty instead (factor 1): the smoother inverts covariances with
# condition numbers up to ~1e10 on long-gap tracks, which bounds what "equal to the reference"
# can mean there.  On the fixtures this widens the bound on 48 of 97 tracks, to at most 1.0e-7
# (smoothed means of c4_ragged_ungated); only that fixture has errors above 1e-9 at all.
UNC_FACTOR = 1.0
WORST = {}   # label -> worst [filtered mean, filtered cov, smoothed mean, smoothed cov] error seen (printed by the GPU tests)


def load_golden(name):
    """-> (list of per-track dicts, dict of extras)."""
    d = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    n = int(d["n_tracks"])
    tracks = [dict() for _ in range(n)]
    extra = {}
    for key in d.files:
        if key == "n_tracks":
            continue
        head, _, tail = key.partition("_")
        if head.startswith("t") and head[1:].isdigit():
            tracks[int(head[1:])][tail] = d[key]
        else:
            extra[key] = d[key]
    return tracks, extra


def circ_diff(a, b):
    return (np.asarray(a) - np.asarray(b) + 180.0) % 360.0 - 180.0


def mean_err(got, ref):
    """Worst per-component error of (N+1, 4) means, normalised by max(1, |ref|)."""
    got, ref = np.asarray(got), np.asarray(ref)
    d = got - ref
    d[..., 3] = circ_diff(got[..., 3], ref[..., 3])
    return float(np.max(np.abs(d) / np.maximum(1.0, np.abs(ref))))


def cov_err(got, ref):
    """Worst per-matrix norm-relative error of (N+1, 4, 4) covariances."""
    got, ref = np.asarray(got), np.asarray(ref)
    num = np.max(np.abs(got - ref), axis=(-2, -1))
    den = np.max(np.abs(ref), axis=(-2, -1))
    return float(np.max(num / np.maximum(den, 1e-300)))


def diag_err(got_diag, ref_diag, ref_scale=None):
    got_diag, ref_diag = np.asarray(got_diag), np.asarray(ref_diag)
    den = np.max(np.abs(ref_diag), axis=-1, keepdims=True) if ref_scale is None else ref_scale
    return float(np.max(np.abs(got_diag - ref_diag) / np.maximum(den, 1e-300)))


def track_errors(got, ref, smoother=True):
    """-> [filtered mean, filtered cov, smoothed mean, smoothed cov] errors (nan = not compared)."""
    e = [mean_err(got["means"], ref["means"]), np.nan, np.nan, np.nan]
    if "covs" in ref:
        e[1] = cov_err(got["covs"], ref["covs"])
    else:
        e[1] = diag_err(np.diagonal(got["covs"], axis1=1, axis2=2), ref["covs_diag"])
    if smoother and "means_s" in ref:
        e[2] = mean_err(got["means_s"], ref["means_s"])
        if "covs_s" in ref:
            e[3] = cov_err(got["covs_s"], ref["covs_s"])
        else:
            e[3] = diag_err(np.diagonal(got["covs_s"], axis1=1, axis2=2), ref["covs_s_diag"])
    return e


def assert_track_close(got, ref, tol=TOL, smoother=True, label="", unc=None):
    """``got``/``ref``: dicts with means, covs[, means_s, covs_s] (or *_diag for slim fixtures).
    ``unc``: the reference's self-uncertainty for the four quantities (defaults to ``ref["unc"]``)."""
    if unc is None:
        unc = ref.get("unc", np.zeros(4))
    names = ("filtered mean", "filtered cov", "smoothed mean", "smoothed cov")
    errs = track_errors(got, ref, smoother)
    key = re.sub(r"\[\d+\]|\s+\d+$", "", label.split(" track")[0].split(" ship")[0]).strip() or "unlabelled"
    WORST[key] = np.fmax(WORST.get(key, np.zeros(4)), np.nan_to_num(np.asarray(errs, dtype=float)))
    for name, e, u in zip(names, errs, unc):
        if np.isnan(e):
            continue
        bound = max(tol, UNC_FACTOR * float(u))
        assert e <= bound, f"{label} {name} err {e:.3e} > {bound:.3e} (reference self-uncertainty {u:.3e})"
