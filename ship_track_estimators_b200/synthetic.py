"""Seeded synthetic ship tracks in the structure-of-arrays layout the kernels consume.

The reference ships no generator; this one is the definition SURVEY.md section 8(d) fixes for the
BASELINE.json synthetic configs, so the oracle subset and the GPU run read the same arrays.
Written with torch ops only, so the same code runs on ``cpu`` (tests, golden fixtures) and on
``cuda`` (bench tiles).  Tensors are ``[obs][track]`` with the track index fastest.

Derived quantities follow the reference's ``ShipTrack`` exactly in *definition* (haversine /
great-circle heading between successive fixes, last value duplicated, backward-difference rates
with a leading zero: reference ``ship_track.py:197-304``); both sides of every parity test consume
the arrays produced here, so ULP differences between torch and numpy trigonometry are irrelevant.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Sequence

import torch

from .constants import EARTH_RADIUS

_D2R = 3.141592653589793 / 180.0


@dataclass
class SyntheticTracks:
    """Observations and derived inputs for ``T`` tracks, each ``nobs[t]`` fixes long."""

    lon: torch.Tensor  # [nobs_max][T] observed longitude (deg)
    lat: torch.Tensor  # [nobs_max][T] observed latitude (deg)
    sog: torch.Tensor  # [nobs_max][T] km/h between fix i and i+1 (last duplicated)
    cog: torch.Tensor  # [nobs_max][T] deg
    sog_rate: torch.Tensor  # [nobs_max][T] backward difference, leading 0
    cog_rate: torch.Tensor  # [nobs_max][T]
    dts: torch.Tensor  # [nobs_max-1][T] hours between fixes
    nobs: torch.Tensor  # [T] int32 fixes per track
    outlier: torch.Tensor  # [nobs_max][T] bool, fixes that were displaced on purpose

    @property
    def n_tracks(self) -> int:
        return self.lon.shape[1]

    def x0(self) -> torch.Tensor:
        """Initial state ``z[:, 0]`` per track -> [4][T] (reference main_cli.py:109)."""
        return torch.stack([self.lon[0], self.lat[0], self.sog[0], self.cog[0]])


def _haversine_km(lon1, lat1, lon2, lat2):
    lam1, phi1, lam2, phi2 = (v * _D2R for v in (lon1, lat1, lon2, lat2))
    a = torch.sin((phi2 - phi1) / 2) ** 2 + torch.cos(phi1) * torch.cos(phi2) * torch.sin((lam2 - lam1) / 2) ** 2
    return 2 * torch.atan2(torch.sqrt(a), torch.sqrt(1 - a)) * EARTH_RADIUS


def _heading_deg(lon1, lat1, lon2, lat2):
    lam1, phi1, lam2, phi2 = (v * _D2R for v in (lon1, lat1, lon2, lat2))
    dlam = lam2 - lam1
    east = torch.sin(dlam) * torch.cos(phi2)
    north = torch.cos(phi1) * torch.sin(phi2) - torch.sin(phi1) * torch.cos(phi2) * torch.cos(dlam)
    return torch.remainder(torch.atan2(east, north) / _D2R + 360.0, 360.0)


def _great_circle_step(lon, lat, sog, cog, dt):
    """One step of the truth model: the same spherical direct problem the filter's process model
    solves (reference ``non_linear_process.py:54-70``)."""
    phi, alpha = lat * _D2R, cog * _D2R
    delta = sog * dt / EARTH_RADIUS
    sd, cd = torch.sin(delta), torch.cos(delta)
    dlon = torch.atan2(sd * torch.sin(alpha), torch.cos(phi) * cd - torch.sin(phi) * sd * torch.cos(alpha))
    s = (torch.sin(phi) * cd + torch.cos(phi) * sd * torch.cos(alpha)).clamp(-1.0, 1.0)
    return lon + dlon / _D2R, torch.asin(s) / _D2R


def _box_smooth_same(y: torch.Tensor, width: int, nobs: torch.Tensor) -> torch.Tensor:
    """``np.convolve(y, ones(w)/w, mode="same")`` along dim 0 for every (ragged) column
    (reference ``utils.py:150-172``); samples past a track's end count as the zero padding."""
    n, T = y.shape
    valid = torch.arange(n, device=y.device)[:, None] < nobs[None, :]
    y = torch.where(valid, y, torch.zeros_like(y))
    # 'same' keeps full[(w-1)//2 : (w-1)//2 + n] where full[j] = sum_m y[m] box[j-m]
    start = (width - 1) // 2
    out = torch.zeros_like(y)
    for tap in range(width):
        shift = start - tap  # out[i] += y[i + shift] / w
        lo, hi = max(0, -shift), min(n, n - shift)
        if hi > lo:
            out[lo:hi] += y[lo + shift : hi + shift] / width
    # a ragged column must not see its neighbour's padding convention differently: entries with
    # i + shift >= nobs are zero already because y was zeroed there.
    return torch.where(valid, out, torch.zeros_like(out))


def make_tracks(
    n_tracks: int,
    nobs: int,
    *,
    seed: int = 0,
    device: str = "cpu",
    dts_choices: Sequence[float] = (1.0,),
    nobs_min: Optional[int] = None,
    outlier_frac: float = 0.0,
    smooth_width: int = 0,
    obs_sigma_deg: float = 0.05,
    lengths: Optional[torch.Tensor] = None,
) -> SyntheticTracks:
    """Generate ``n_tracks`` tracks with up to ``nobs`` fixes each.

    ``nobs_min`` (ragged): lengths are uniform in ``[nobs_min, nobs]`` and the tracks are returned
    sorted by decreasing length.  ``dts_choices``: hours between fixes, drawn per gap.
    ``outlier_frac``: fraction of fixes displaced by 5-50 degrees.  ``smooth_width`` >= 2 applies
    the CLI's box smoothing to SOG and COG before the rates are differenced
    (reference ``main_cli.py:99-108``).  ``lengths`` ([T] int, each in [2, nobs]) fixes the number
    of fixes per track instead of drawing it (a tile of a length-sorted fleet).
    """
    dev = torch.device(device)
    gen = torch.Generator(device=dev)
    gen.manual_seed(int(seed))
    f64 = dict(dtype=torch.float64, device=dev)
    T = int(n_tracks)

    def uniform(lo, hi, *shape):
        return lo + (hi - lo) * torch.rand(*shape, generator=gen, **f64)

    def normal(*shape):
        return torch.randn(*shape, generator=gen, **f64)

    if lengths is not None:
        lengths = torch.as_tensor(lengths).to(device=dev, dtype=torch.int32)
        if lengths.shape != (T,) or int(lengths.min()) < 2 or int(lengths.max()) > nobs:
            raise ValueError("lengths must hold one value in [2, nobs] per track")
    elif nobs_min is None or nobs_min >= nobs:
        lengths = torch.full((T,), nobs, dtype=torch.int32, device=dev)
    else:
        lengths = torch.randint(nobs_min, nobs + 1, (T,), generator=gen, device=dev, dtype=torch.int32)
        lengths, _ = torch.sort(lengths, descending=True)

    choices = torch.tensor(list(dts_choices), **f64)
    if len(dts_choices) == 1:
        dts = choices[0].expand(nobs - 1, T).clone()
    else:
        dts = choices[torch.randint(0, len(dts_choices), (nobs - 1, T), generator=gen, device=dev)]

    lon_t = uniform(-180.0, 180.0, T)
    lat_t = uniform(-60.0, 60.0, T)
    sog_t = uniform(5.0, 30.0, T)
    cog_t = uniform(0.0, 360.0, T)

    lon = torch.empty(nobs, T, **f64)
    lat = torch.empty(nobs, T, **f64)
    for i in range(nobs):
        lon[i] = lon_t + obs_sigma_deg * normal(T)
        lat[i] = lat_t + obs_sigma_deg * normal(T)
        if i + 1 < nobs:
            lon_t, lat_t = _great_circle_step(lon_t, lat_t, sog_t, cog_t, dts[i])
            cog_t = cog_t + 2.0 * normal(T)
            # keep the synthetic fleet out of the polar caps (the filter itself has no such limit)
            cog_t = torch.where(lat_t > 60.0, torch.full_like(cog_t, 180.0), cog_t)
            cog_t = torch.where(lat_t < -60.0, torch.zeros_like(cog_t), cog_t)
            cog_t = torch.remainder(cog_t, 360.0)
            sog_t = (sog_t + 0.2 * normal(T)).clamp(0.5, 60.0)

    outlier = torch.zeros(nobs, T, dtype=torch.bool, device=dev)
    if outlier_frac > 0.0:
        outlier = torch.rand(nobs, T, generator=gen, **f64) < outlier_frac
        outlier[0] = False  # x0 comes from the first fix
        size = uniform(5.0, 50.0, nobs, T)
        sign = torch.where(torch.rand(nobs, T, generator=gen, **f64) < 0.5, -1.0, 1.0)
        lon = torch.where(outlier, lon + sign * size, lon)
        lat = torch.where(outlier, (lat - sign * size * 0.5).clamp(-89.0, 89.0), lat)

    # derived inputs, as ShipTrack computes them (ship_track.py:197-304)
    idx = torch.arange(nobs, device=dev)[:, None]
    sog = torch.empty(nobs, T, **f64)
    cog = torch.empty(nobs, T, **f64)
    sog[:-1] = _haversine_km(lon[:-1], lat[:-1], lon[1:], lat[1:]) / dts
    cog[:-1] = _heading_deg(lon[:-1], lat[:-1], lon[1:], lat[1:])
    sog[-1], cog[-1] = sog[-2], cog[-2]
    # ragged: each track's own last fix duplicates its own previous value
    last = (lengths.long() - 1)[None, :]
    prev = (last - 1).clamp(min=0)
    sog = torch.where(idx == last, torch.gather(sog, 0, prev).expand_as(sog), sog)
    cog = torch.where(idx == last, torch.gather(cog, 0, prev).expand_as(cog), cog)

    if smooth_width and smooth_width > 1:
        sog = _box_smooth_same(sog, smooth_width, lengths)
        cog = _box_smooth_same(cog, smooth_width, lengths)

    sog_rate = torch.zeros(nobs, T, **f64)
    cog_rate = torch.zeros(nobs, T, **f64)
    sog_rate[1:] = (sog[1:] - sog[:-1]) / dts
    cog_rate[1:] = (cog[1:] - cog[:-1]) / dts

    valid = idx < lengths[None, :]
    zero = torch.zeros(nobs, T, **f64)
    sog_rate = torch.where(valid, sog_rate, zero)
    cog_rate = torch.where(valid, cog_rate, zero)
    return SyntheticTracks(
        lon=lon, lat=lat, sog=sog, cog=cog, sog_rate=sog_rate, cog_rate=cog_rate,
        dts=dts, nobs=lengths, outlier=outlier & valid,
    )
