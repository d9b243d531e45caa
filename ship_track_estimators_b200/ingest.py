"""Bulk CSV -> structure-of-arrays ingest for whole fleets.

The reference reads one ship at a time: ``ShipTrack.read_csv`` parses the whole file, keeps the
rows of one id and builds that ship's ``lon / lat / dts`` (``ship_track.py:107-195``); its batch
example therefore re-parses the CSV once per ship (``example_ukf_rts_smoother_batch.py:15-34``).
``read_csv_fleet`` parses the file once and lays every ship out as the device kernels want it:
``lon, lat [max_obs][T]``, ``dts [max_obs-1][T]``, ``n_obs [T]`` - same row selection (ids compared
as strings), same row order (``DataFrame.sort_index`` inside a ship, no sort by date), same time arithmetic (``yr-mo-dy`` +
``hr`` -> gaps in hours) as the per-ship reader.  ``FleetFixes.to_batch`` then derives speed,
course and their rates on the device (``derive.batch_from_fixes``) without a host round trip.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np
import torch


@dataclass
class FleetFixes:
    """Raw fixes of ``T`` ships, padded to the longest (host arrays, track index fastest)."""

    ids: List[str]
    lon: np.ndarray       # [max_obs][T] degrees
    lat: np.ndarray       # [max_obs][T] degrees
    dts: np.ndarray       # [max_obs-1][T] hours between successive fixes (0 beyond a ship's last gap)
    n_obs: np.ndarray     # [T] int32

    @property
    def n_tracks(self) -> int:
        return len(self.ids)

    def track(self, i: int):
        """(lat, lon, dts) of ship ``i`` exactly as ``ShipTrack.read_csv`` returns them."""
        n = int(self.n_obs[i])
        return self.lat[:n, i].copy(), self.lon[:n, i].copy(), self.dts[: max(n - 1, 0), i].copy()

    def select(self, keep: Sequence[int]) -> "FleetFixes":
        keep = list(keep)
        n = int(self.n_obs[keep].max()) if keep else 0
        return FleetFixes([self.ids[i] for i in keep], self.lon[:n, keep].copy(), self.lat[:n, keep].copy(),
                          self.dts[: max(n - 1, 0), keep].copy(), self.n_obs[keep].copy())

    def to_batch(self, device="cuda", substeps: int = 1, smooth_width: int = 0, need_rows=(True, True, False, False),
                 geodesy: str = "sphere", sort_by_length: bool = True):
        """Upload the fixes and build the filter inputs on the device (ships need >= 2 fixes).

        Ships are packed in order of decreasing length (``sort_by_length``, stable): the lanes of a
        warp then finish together.  The permutation is recorded in ``TrackBatch.order`` and undone by
        ``TrackResults.track(i)``, so ``i`` keeps meaning ship ``ids[i]``."""
        from .derive import batch_from_fixes

        if self.n_tracks and int(self.n_obs.min()) < 2:
            raise ValueError("every ship needs at least two fixes; drop the others with select()")
        order = None
        lon, lat, dts, n_obs = self.lon, self.lat, self.dts, self.n_obs
        if sort_by_length and self.n_tracks > 1 and np.any(n_obs[:-1] < n_obs[1:]):
            order = np.argsort(-n_obs.astype(np.int64), kind="stable")
            lon, lat, dts, n_obs = lon[:, order], lat[:, order], dts[:, order], n_obs[order]
        up = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(device)   # noqa: E731
        batch = batch_from_fixes(up(lon), up(lat), up(dts), up(n_obs), substeps=substeps,
                                 smooth_width=smooth_width, need_rows=need_rows, geodesy=geodesy)
        batch.order = order
        return batch


def read_csv_fleet(csv_file: str, id_col: str = "id", lat_col: str = "lat", lon_col: str = "lon",
                   ship_ids: Optional[Sequence] = None, reverse: bool = False, on_bad_rows: str = "raise") -> FleetFixes:
    """All ships of ``csv_file`` (or those in ``ship_ids``, in that order) in one pass.

    Ships appear in order of first occurrence in the file unless ``ship_ids`` is given.  A requested
    id without rows raises ``ValueError`` as ``ShipTrack.read_csv`` does.  A ship with a row whose
    date or position does not parse raises too (the per-ship reader would, for that ship) unless
    ``on_bad_rows="skip"``, which leaves such ships out (the reference's batch example drops the
    stray repeated-header "ship" of the historical file by hand, ``example_ukf_rts_smoother_batch.py:17``)."""
    if on_bad_rows not in ("raise", "skip"):
        raise ValueError("on_bad_rows must be 'raise' or 'skip'")
    import pandas as pd

    df = pd.read_csv(csv_file)
    ids_col = df[id_col].astype(str)
    stamp = (df["yr"].astype(str) + "-" + df["mo"].astype(str) + "-" + df["dy"].astype(str)
             + "T" + df["hr"].astype(str).str.zfill(2) + ":00:00")
    # Files in the wild carry stray rows (the historical data repeats its header line); the per-ship
    # reader never sees them because it filters by id first.  Parse leniently here and complain
    # below only if a row that is actually used did not parse.
    when = pd.to_datetime(stamp, errors="coerce")
    lat_num, lon_num = pd.to_numeric(df[lat_col], errors="coerce"), pd.to_numeric(df[lon_col], errors="coerce")
    unparsed = (when.isna() | (lat_num.isna() & df[lat_col].notna()) | (lon_num.isna() & df[lon_col].notna())).to_numpy()
    hours = when.to_numpy().astype("datetime64[s]").astype(np.int64)                       # seconds since epoch
    lat_all = lat_num.to_numpy(dtype=np.float64)
    lon_all = lon_num.to_numpy(dtype=np.float64)

    codes, uniques = pd.factorize(ids_col, sort=False)              # first-occurrence order
    # Inside a ship the per-ship reader orders rows with DataFrame.sort_index (ship_track.py:151).
    # That is file order for a default RangeIndex, but these files carry a leading unnamed column
    # that pandas turns into the index - and a stray header row makes its labels strings, so the
    # order is lexicographic by label.  Reproduce it: rank every row by its index label first.
    labels = df.index.to_numpy()
    if labels.dtype == object and all(isinstance(v, str) for v in labels[:64]):
        labels = labels.astype(str)                                 # fixed-width unicode: same order, C-speed compares
    by_label = np.argsort(labels, kind="stable")
    order = by_label[np.argsort(codes[by_label], kind="stable")]    # rows grouped by ship, label order inside
    counts = np.bincount(codes, minlength=len(uniques))
    starts = np.concatenate(([0], np.cumsum(counts)[:-1]))
    if ship_ids is None:
        wanted = np.arange(len(uniques))
    else:
        index = {u: i for i, u in enumerate(uniques)}
        missing = [str(s) for s in ship_ids if str(s) not in index]
        if missing:
            raise ValueError(f"No data found for ship '{missing[0]}' in '{csv_file}'.")
        wanted = np.asarray([index[str(s)] for s in ship_ids], dtype=np.int64)
    if on_bad_rows == "skip":
        bad_codes = np.unique(codes[unparsed])
        wanted = wanted[~np.isin(wanted, bad_codes)]
    T = len(wanted)
    n_obs = counts[wanted].astype(np.int32)
    max_obs = int(n_obs.max()) if T else 0
    lon = np.zeros((max_obs, T))
    lat = np.zeros((max_obs, T))
    dts = np.zeros((max(max_obs - 1, 0), T))
    # scatter: position of every kept row inside its ship, then one fancy-indexed assignment
    col_of_code = np.full(len(uniques), -1, dtype=np.int64)
    col_of_code[wanted] = np.arange(T)
    g_codes = codes[order]
    pos = np.arange(len(order)) - starts[g_codes]
    cols = col_of_code[g_codes]
    keep = cols >= 0
    rows, cols, src = pos[keep], cols[keep], order[keep]
    if unparsed[src].any():
        bad = int(src[unparsed[src]][0])
        raise ValueError(f"row {bad} of '{csv_file}' (ship '{ids_col.iloc[bad]}') has an unparsable date or position")
    if reverse:                                                     # ShipTrack.read_csv(reverse=True): arrays flipped
        rows = n_obs[cols] - 1 - rows
    lon[rows, cols] = lon_all[src]
    lat[rows, cols] = lat_all[src]
    sec = np.zeros((max_obs, T), dtype=np.int64)
    sec[rows, cols] = hours[src]
    if max_obs > 1:
        gap = np.diff(sec, axis=0).astype(np.float64) / 3600.0
        if reverse:                                                 # gaps are taken in file order, then flipped
            gap = -gap
        valid = np.arange(max_obs - 1)[:, None] < (n_obs[None, :] - 1)
        dts[valid] = gap[valid]
    return FleetFixes([str(uniques[i]) for i in wanted], lon, lat, dts, n_obs)
