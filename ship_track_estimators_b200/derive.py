"""Derived filter inputs for many tracks on the device.

What ``ShipTrack`` computes per ship on the host - speed and course over ground between successive
fixes, the CLI's optional box smoothing, and the backward-difference rates (reference
``ship_track.py:197-304``, ``utils.py:75-172``, ``cli/main_cli.py:99-108``) - for a whole
structure-of-arrays tile of raw fixes in one kernel launch.  Distances and headings are either the
reference's spherical pair ``haversine_formula`` / ``heading`` (``geodesy="sphere"``) or its default,
the WGS84 inverse geodesic it takes from the third-party ``geographiclib`` (``geodesy="wgs84"``,
restated with Vincenty's inverse formulae; see ``include/ste_ukf.h``).
"""
from __future__ import annotations

from typing import Optional

import torch

from . import _native as nat
from .batch import TrackBatch
from .synthetic import SyntheticTracks


def derive_inputs(lon: torch.Tensor, lat: torch.Tensor, dts: torch.Tensor, n_obs: Optional[torch.Tensor] = None,
                  smooth_width: int = 0, geodesy: str = "sphere") -> SyntheticTracks:
    """``lon, lat [nobs][T]`` (deg), ``dts [nobs-1][T]`` (h), ``n_obs [T]`` int32 (ragged) -> all
    inputs of the filter as a :class:`SyntheticTracks` bundle (usable with ``TrackBatch.from_synthetic``)."""
    lib = nat.load()
    model = {"sphere": nat.STE_GEODESY_SPHERE, "wgs84": nat.STE_GEODESY_WGS84}[geodesy]
    nobs, T = lon.shape
    if lat.shape != lon.shape or dts.shape != (nobs - 1, T):
        raise ValueError("lon/lat must be [nobs][T] and dts [nobs-1][T]")
    lon, lat, dts = lon.contiguous(), lat.contiguous(), dts.contiguous()
    if n_obs is None:
        n_obs = torch.full((T,), nobs, dtype=torch.int32, device=lon.device)
    n_obs = n_obs.to(torch.int32).contiguous()
    out = [torch.empty_like(lon) for _ in range(4)]
    with torch.cuda.device(lon.device):
        nat.check(lib.ste_derive_inputs_f64(T, nobs, T, int(smooth_width), model, nat.ptr(lon), nat.ptr(lat), nat.ptr(dts), nat.ptr(n_obs),
                                            *(nat.ptr(o) for o in out), nat.current_stream()))
    sog, cog, sog_rate, cog_rate = out
    return SyntheticTracks(lon=lon, lat=lat, sog=sog, cog=cog, sog_rate=sog_rate, cog_rate=cog_rate, dts=dts, nobs=n_obs,
                           outlier=torch.zeros_like(lon, dtype=torch.bool))


def batch_from_fixes(lon, lat, dts, n_obs=None, substeps: int = 1, smooth_width: int = 0, need_rows=(True, True, False, False),
                     geodesy: str = "sphere") -> TrackBatch:
    """Raw fixes on the device -> a ready :class:`TrackBatch` (derivation + packing, no host round trip)."""
    return TrackBatch.from_synthetic(derive_inputs(lon, lat, dts, n_obs, smooth_width, geodesy), substeps=substeps, need_rows=need_rows)
