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


# ------------------------------------------------------------------------------------------------ #
# device ingest: the file's bytes are parsed on the GPU                                              #
# ------------------------------------------------------------------------------------------------ #
@dataclass
class DeviceFleetFixes:
    """Raw fixes of ``T`` ships as device tensors (same layout as :class:`FleetFixes`)."""

    ids: List[str]
    lon: torch.Tensor      # [max_obs][T]
    lat: torch.Tensor      # [max_obs][T]
    dts: torch.Tensor      # [max_obs-1][T]
    n_obs: torch.Tensor    # [T] int32
    stats: dict            # rows, bytes and the seconds spent in each stage

    @property
    def n_tracks(self) -> int:
        return len(self.ids)

    def to_host(self) -> FleetFixes:
        return FleetFixes(list(self.ids), self.lon.cpu().numpy(), self.lat.cpu().numpy(), self.dts.cpu().numpy(),
                          self.n_obs.cpu().numpy().astype(np.int32))

    def to_batch(self, substeps: int = 1, smooth_width: int = 0, need_rows=(True, True, False, False), geodesy: str = "sphere",
                 sort_by_length: bool = True):
        """Filter inputs derived on the device from the device-resident fixes (no host round trip at all)."""
        from .derive import batch_from_fixes

        if self.n_tracks and int(self.n_obs.min()) < 2:
            raise ValueError("every ship needs at least two fixes; drop the others first")
        lon, lat, dts, n_obs, order = self.lon, self.lat, self.dts, self.n_obs, None
        if sort_by_length and self.n_tracks > 1 and bool((n_obs[:-1] < n_obs[1:]).any()):
            perm = torch.argsort(-n_obs.to(torch.int64), stable=True)
            lon, lat, dts, n_obs = lon[:, perm].contiguous(), lat[:, perm].contiguous(), dts[:, perm].contiguous(), n_obs[perm].contiguous()
            order = perm.cpu().numpy()
        batch = batch_from_fixes(lon, lat, dts, n_obs, substeps=substeps, smooth_width=smooth_width, need_rows=need_rows, geodesy=geodesy)
        batch.order = order
        return batch


_CSV_BAD_DATE, _CSV_BAD_POS, _CSV_ID_TEXT, _CSV_LABEL_TEXT, _CSV_SLOW, _CSV_SHORT = 0x1, 0x2, 0x4, 0x8, 0x10, 0x20


def _split_header(line: bytes) -> List[str]:
    return [f.strip().strip('"') for f in line.decode("utf-8", "replace").rstrip("\r\n").split(",")]


def read_csv_fleet_device(csv_file: str, id_col: str = "id", lat_col: str = "lat", lon_col: str = "lon",
                          ship_ids: Optional[Sequence] = None, reverse: bool = False, on_bad_rows: str = "raise",
                          device="cuda") -> DeviceFleetFixes:
    """``read_csv_fleet`` with the parsing on the GPU: the file's bytes go to the device once, one thread
    per row extracts time stamp, position and ship id (``ste_csv_parse_rows``), ships are grouped and
    rows ordered by device sorts, and the fixes are scattered into the ``[max_obs][T]`` layout - the host
    only reads the header line and one id string per ship.  Same row selection, row order inside a
    ship, time arithmetic, ``reverse`` and error behaviour as :func:`read_csv_fleet` /
    ``ShipTrack.read_csv``, with one restriction: when the file carries an index column its labels
    must be integers (pandas then sorts them numerically); a file whose labels are text - the
    historical data set with its stray repeated header line - raises ``NotImplementedError`` and
    belongs to the host reader."""
    import ctypes as C
    import time

    from . import _native as nat

    if on_bad_rows not in ("raise", "skip"):
        raise ValueError("on_bad_rows must be 'raise' or 'skip'")
    lib, dev = nat.load(), torch.device(device)
    t0 = time.perf_counter()
    raw = np.fromfile(csv_file, dtype=np.uint8)
    if raw.size == 0 or raw[-1] != 10:
        raw = np.concatenate([raw, np.array([10], dtype=np.uint8)])
    t_read = time.perf_counter() - t0
    data = torch.from_numpy(raw).to(dev, non_blocking=False)
    nl = torch.nonzero(data == 10).reshape(-1)                       # offsets of the newlines
    row_start = torch.cat([torch.zeros(1, dtype=torch.int64, device=dev), nl + 1])   # [rows + 1]
    n_lines = int(nl.numel())
    if n_lines < 1:
        raise ValueError(f"'{csv_file}' is empty")
    first_nl = int(nl[0])
    header = _split_header(raw[:first_nl].tobytes())
    n_rows = n_lines - 1
    cols = np.full(8, -1, dtype=np.int32)
    shift = 0
    if n_rows > 0:
        second_nl = int(nl[1])
        n_fields = len(_split_header(raw[first_nl + 1:second_nl].tobytes()))
        if n_fields == len(header) + 1:     # one name fewer than fields: pandas makes the first field the index
            shift, cols[7] = 1, 0
    for k, name in enumerate(("yr", "mo", "dy", "hr", lat_col, lon_col, id_col)):
        if name not in header:
            raise KeyError(name)
        cols[k] = header.index(name) + shift
    i64 = dict(dtype=torch.int64, device=dev)
    out = dict(hours=torch.empty(n_rows, **i64), lat=torch.empty(n_rows, dtype=torch.float64, device=dev),
               lon=torch.empty(n_rows, dtype=torch.float64, device=dev), id_key=torch.empty(n_rows, **i64), id_int=torch.empty(n_rows, **i64),
               id_off=torch.empty(n_rows, dtype=torch.int32, device=dev), id_len=torch.empty(n_rows, dtype=torch.int32, device=dev),
               label=torch.empty(n_rows, **i64), flags=torch.empty(n_rows, dtype=torch.int32, device=dev))
    torch.cuda.synchronize(dev)
    t1 = time.perf_counter()
    with torch.cuda.device(dev):
        nat.check(lib.ste_csv_parse_rows(nat.ptr(data), nat.ptr(row_start[1:]), n_rows, cols.ctypes.data_as(C.POINTER(C.c_int32)),
                                         *(nat.ptr(out[k]) for k in ("hours", "lat", "lon", "id_key", "id_int", "id_off", "id_len", "label", "flags")),
                                         nat.current_stream()))
    torch.cuda.synchronize(dev)
    t_parse = time.perf_counter() - t1
    data_start = row_start[1:]       # row r of the data starts here (row 0 of the file is the header)
    flags = out["flags"]
    any_flags = int(torch.bitwise_or(flags, torch.zeros_like(flags)).max()) if n_rows else 0
    any_flags = 0
    if n_rows:
        for bit in (_CSV_BAD_DATE, _CSV_BAD_POS, _CSV_ID_TEXT, _CSV_LABEL_TEXT, _CSV_SLOW, _CSV_SHORT):
            if bool((flags & bit).any()):
                any_flags |= bit
    if shift and (any_flags & _CSV_LABEL_TEXT):
        raise NotImplementedError("the index labels of this file are not all integers (pandas would order the rows of a ship "
                                  "lexicographically): use read_csv_fleet")
    if any_flags & _CSV_SLOW:        # a handful of over-long literals: the host re-reads exactly those fields
        rows = torch.nonzero((flags & _CSV_SLOW) != 0).reshape(-1).cpu().numpy()
        starts = data_start[torch.from_numpy(rows).to(dev)].cpu().numpy()
        ends = row_start[2:][torch.from_numpy(rows).to(dev)].cpu().numpy()
        for r, s, e in zip(rows, starts, ends):
            fields = _split_header(raw[s:e - 1].tobytes())
            for name, ci in (("lat", cols[4]), ("lon", cols[5])):
                try:
                    out[name][r] = float(fields[ci]) if fields[ci] not in ("", "NA", "NaN", "nan", "NULL") else float("nan")
                except ValueError:
                    flags[r] |= _CSV_BAD_POS
    bad = (flags & (_CSV_BAD_DATE | _CSV_BAD_POS | _CSV_SHORT)) != 0
    # ship key: pandas compares ids as str(value); an all-integer id column is parsed as integers first
    numeric_ids = not (any_flags & _CSV_ID_TEXT)
    key = out["id_int"] if numeric_ids else out["id_key"]
    t2 = time.perf_counter()
    rows_idx = torch.arange(n_rows, **i64)
    uniq, code = torch.unique(key, return_inverse=True)
    G = int(uniq.numel())
    first = torch.full((G,), n_rows, **i64).scatter_reduce(0, code, rows_idx, reduce="amin")     # first occurrence of each ship
    by_first = torch.argsort(first)                      # ships in order of first occurrence, as the host reader lists them
    rank_of_group = torch.empty(G, **i64)
    rank_of_group[by_first] = torch.arange(G, **i64)
    code = rank_of_group[code]
    first = first[by_first]
    # the id text of one row per ship, read from the host copy of the bytes
    f_cpu = first.cpu().numpy()
    offs = (data_start[first] + out["id_off"][first].to(torch.int64)).cpu().numpy()
    lens = out["id_len"][first].cpu().numpy()
    names = [raw[o:o + n].tobytes().decode("utf-8", "replace") for o, n in zip(offs, lens)]
    if numeric_ids:
        names = [str(int(v)) for v in out["id_int"][first].cpu().numpy()]
    del f_cpu
    counts = torch.bincount(code, minlength=G)
    if ship_ids is None:
        wanted = torch.arange(G, **i64)
    else:
        index = {u: i for i, u in enumerate(names)}
        missing = [str(s) for s in ship_ids if str(s) not in index]
        if missing:
            raise ValueError(f"No data found for ship '{missing[0]}' in '{csv_file}'.")
        wanted = torch.tensor([index[str(s)] for s in ship_ids], **i64)
    bad_groups = torch.zeros(G, dtype=torch.bool, device=dev)
    if bool(bad.any()):
        bad_groups[code[bad]] = True
        if on_bad_rows == "skip":
            wanted = wanted[~bad_groups[wanted]]
        elif bool(bad_groups[wanted].any()):
            r = int(torch.nonzero(bad & bad_groups[code] & torch.isin(code, wanted)).reshape(-1)[0])
            raise ValueError(f"row {r} of '{csv_file}' (ship '{names[int(code[r])]}') has an unparsable date or position")
    T = int(wanted.numel())
    col_of_group = torch.full((G,), -1, **i64)
    col_of_group[wanted] = torch.arange(T, **i64)
    # order: by ship, inside a ship by index label (numeric) or by position in the file
    inner = out["label"] if shift else rows_idx
    by_inner = torch.argsort(inner, stable=True)
    order = by_inner[torch.argsort(code[by_inner], stable=True)]
    starts = torch.cumsum(counts, 0) - counts
    g_sorted = code[order]
    pos = torch.arange(n_rows, **i64) - starts[g_sorted]
    cols_t = col_of_group[g_sorted]
    keep = cols_t >= 0
    rows_, cols_, src = pos[keep], cols_t[keep], order[keep]
    n_obs = counts[wanted].to(torch.int32)
    max_obs = int(n_obs.max()) if T else 0
    if reverse:
        rows_ = n_obs.to(torch.int64)[cols_] - 1 - rows_
    f64 = dict(dtype=torch.float64, device=dev)
    lon, lat = torch.zeros(max_obs, T, **f64), torch.zeros(max_obs, T, **f64)
    hrs = torch.zeros(max_obs, T, **i64)
    lon[rows_, cols_], lat[rows_, cols_], hrs[rows_, cols_] = out["lon"][src], out["lat"][src], out["hours"][src]
    dts = torch.zeros(max(max_obs - 1, 0), T, **f64)
    if max_obs > 1:
        gap = (hrs[1:] - hrs[:-1]).to(torch.float64)
        if reverse:
            gap = -gap
        valid = torch.arange(max_obs - 1, device=dev)[:, None] < (n_obs.to(torch.int64)[None, :] - 1)
        dts = torch.where(valid, gap, torch.zeros_like(gap))
    torch.cuda.synchronize(dev)
    t_group = time.perf_counter() - t2
    stats = dict(rows=n_rows, bytes=int(raw.size), read_s=t_read, h2d_and_lines_s=t1 - t0 - t_read, parse_s=t_parse, group_scatter_s=t_group,
                 total_s=time.perf_counter() - t0)
    return DeviceFleetFixes([names[int(g)] for g in wanted.cpu().numpy()], lon, lat, dts, n_obs, stats)
