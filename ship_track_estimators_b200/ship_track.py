"""Per-ship observation container: CSV ingest and the derived inputs of the UKF.

Drop-in for reference ``src/track_estimators/ship_track.py``.  Host-side O(nobs) preprocessing
(SURVEY.md section 8 "next" rows N1/N2) - it feeds the GPU path but is not part of it.  The attribute
names, method names, argument meanings and the returned values follow the reference; the
implementation is written around numpy arrays rather than Python lists.
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple, Union

import numpy as np

from .utils import geographiclib_distance, geographiclib_heading


def _pairwise(func: Callable, lon: np.ndarray, lat: np.ndarray) -> np.ndarray:
    """``func(lon[i-1], lat[i-1], lon[i], lat[i])`` for successive fixes (scalar calls: ``func`` may
    be any user callable, e.g. a geodesic solver that does not broadcast)."""
    n = len(lon)
    return np.fromiter((func(lon[i - 1], lat[i - 1], lon[i], lat[i]) for i in range(1, n)), dtype=np.float64, count=max(n - 1, 0))


class ShipTrack:
    """Observations of one ship: ``lon``/``lat`` fixes, gaps ``dts`` (hours), and the derived speed
    over ground ``sog`` (km/h), course over ground ``cog`` (deg), their rates and the measurement
    matrix ``z`` (reference ``ship_track.py:9-105``)."""

    def __init__(
        self,
        csv_file: Optional[str] = None,
        estimate_cog: bool = False,
        estimate_sog: bool = False,
        estimate_sog_rate: bool = False,
        estimate_cog_rate: bool = False,
        calc_distance_func: Callable = geographiclib_distance,
        calc_heading_func: Callable = geographiclib_heading,
    ) -> None:
        self.lat = self.lon = self.cog = self.sog = self.dts = self.dates = None
        self.df = None
        self.sog_rate = self.cog_rate = self.z = None
        self.calc_distance_func = calc_distance_func
        self.calc_heading_func = calc_heading_func
        if csv_file is not None:
            self.read_csv(csv_file=csv_file)
            if estimate_sog_rate:
                self.calculate_sog_rate()
            elif estimate_sog:
                self.calculate_sog()
            if estimate_cog_rate:
                self.calculate_cog_rate()
            elif estimate_cog:
                self.calculate_cog()
            self.z = self.get_measurements()

    # ------------------------------------------------------------------ #
    def read_csv(
        self,
        csv_file: str,
        ship_id: Optional[Union[str, int]] = None,
        id_col: str = "id",
        lat_col: str = "lat",
        lon_col: str = "lon",
        reverse: bool = False,
    ) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
        """Rows of ``ship_id`` from a CSV with ``yr, mo, dy, hr`` columns, in file order (the
        reference does not sort by date, ``ship_track.py:151, 179``); gaps in hours.
        As in the reference ``ship_id`` is compared as a string, so an id is effectively mandatory."""
        import pandas as pd

        df = pd.read_csv(csv_file)
        df[id_col] = df[id_col].astype(str)
        df = df.loc[df[id_col] == str(ship_id)].sort_index(axis=0, ignore_index=True)
        if df.empty:
            raise ValueError(f"No data found for ship '{ship_id}' in '{csv_file}'.")
        stamp = (df["yr"].astype(str) + "-" + df["mo"].astype(str) + "-" + df["dy"].astype(str)
                 + "T" + df["hr"].astype(str).str.zfill(2) + ":00:00")
        df["date"] = stamp
        self.df = df
        when = pd.to_datetime(stamp)
        self.dates = when.to_list()
        self.dts = (when.diff().dt.total_seconds().to_numpy()[1:] / 3600.0).astype(np.float64)
        self.lat = pd.to_numeric(df[lat_col]).values
        self.lon = pd.to_numeric(df[lon_col]).values
        assert len(self.lon) > 0, f"Longitude list is empty for column '{lon_col}'."
        assert len(self.lat) > 0, f"Latitude list is empty for column '{lat_col}'."
        assert len(self.lat) == len(self.lon)
        if reverse:
            self.dts, self.lat, self.lon = self.dts[::-1], self.lat[::-1], self.lon[::-1]
        return self.lat, self.lon, self.dts

    # ------------------------------------------------------------------ #
    def calculate_sog(self) -> np.ndarray:
        """Distance between successive fixes over the gap; the last value is repeated
        (reference ``ship_track.py:197-224``)."""
        leg = _pairwise(self.calc_distance_func, self.lon, self.lat) / self.dts
        self.sog = np.append(leg, leg[-1])
        return self.sog

    def calculate_cog(self) -> np.ndarray:
        """Heading from each fix to the next; the last value is repeated (``:252-278``)."""
        leg = _pairwise(self.calc_heading_func, self.lon, self.lat)
        self.cog = np.append(leg, leg[-1])
        return self.cog

    @staticmethod
    def _backward_rate(values: np.ndarray, dts: np.ndarray) -> np.ndarray:
        return np.concatenate(([0.0], np.diff(values) / dts[: len(values) - 1]))

    def calculate_sog_rate(self) -> np.ndarray:
        """Backward difference of ``sog`` with a leading 0 (``:226-250``)."""
        if self.sog is None:
            self.calculate_sog()
        self.sog_rate = self._backward_rate(np.asarray(self.sog, dtype=np.float64), self.dts)
        return self.sog_rate

    def calculate_cog_rate(self) -> np.ndarray:
        """Backward difference of ``cog`` with a leading 0 (``:280-304``)."""
        if self.cog is None:
            self.calculate_cog()
        self.cog_rate = self._backward_rate(np.asarray(self.cog, dtype=np.float64), self.dts)
        return self.cog_rate

    def get_measurements(self, include_sog: bool = False, include_cog: bool = False) -> np.ndarray:
        """Measurement matrix with rows lon, lat [, sog] [, cog] (``:306-338``)."""
        rows = [self.lon, self.lat]
        if include_sog:
            if self.sog is None:
                self.calculate_sog()
            rows.append(self.sog)
        if include_cog:
            if self.cog is None:
                self.calculate_cog()
            rows.append(self.cog)
        self.z = np.vstack(rows)
        return self.z

    def plot_trajectory(self, figsize: tuple = (20, 15), scatter_kwargs: dict = {"s": 10, "color": "red"},
                        savefig: Optional[str] = None, show: bool = True):
        """Scatter of the fixes on a PlateCarree map (needs cartopy + matplotlib; ``:340-392``)."""
        import cartopy.crs as ccrs
        import matplotlib.pyplot as plt

        assert self.lat is not None, "Latitude is not set."
        assert self.lon is not None, "Longitude is not set."
        fig, ax = plt.subplots(nrows=1, ncols=1, subplot_kw={"projection": ccrs.PlateCarree()}, figsize=figsize)
        ax.stock_img()
        ax.coastlines()
        ax.gridlines(crs=ccrs.PlateCarree(), draw_labels=True, linewidth=0.6, color="gray", alpha=0.5, linestyle="-.")
        ax.scatter(self.lon, self.lat, transform=ccrs.PlateCarree(), **scatter_kwargs)
        if savefig:
            plt.savefig(savefig, bbox_inches="tight", dpi=300)
        if show:
            plt.show()
        return fig, ax
