#!/usr/bin/env python
"""One ship through the reference-style class API on the GPU: UKF + URTSS, as the reference's single-ship example does
(examples/example_ukf_rts_smoother.py there), with this package's drop-in names.

    python examples/single_ship.py <tracks.csv> <ship id> [--id-col primary.id --lat-col lat --lon-col lon --substeps 2]
"""
import argparse

import numpy as np

from ship_track_estimators_b200.kalman_filters import UnscentedKalmanFilter, geodetic_dynamics
from ship_track_estimators_b200.ship_track import ShipTrack
from ship_track_estimators_b200.utils import generate_dts, haversine_formula, heading

ap = argparse.ArgumentParser()
ap.add_argument("csv")
ap.add_argument("ship_id")
ap.add_argument("--id-col", default="primary.id")
ap.add_argument("--lat-col", default="lat")
ap.add_argument("--lon-col", default="lon")
ap.add_argument("--substeps", type=int, default=2)
a = ap.parse_args()

track = ShipTrack(calc_distance_func=haversine_formula, calc_heading_func=heading)
track.read_csv(a.csv, ship_id=a.ship_id, id_col=a.id_col, lat_col=a.lat_col, lon_col=a.lon_col)
track.calculate_sog_rate()
track.calculate_cog_rate()
track.get_measurements(include_sog=True, include_cog=True)

H = np.diag([1.0, 1.0, 0.0, 0.0])
R = np.diag([1e-3, 1e-3, 0.0, 0.0])
Q = np.diag([1e-2, 1e-2, 1e-4, 1e-4])
ukf = UnscentedKalmanFilter(H=H, Q=Q, R=R, P=np.eye(4), x0=track.z[:, 0], non_linear_process=geodetic_dynamics)
dt = generate_dts(track.dts, a.substeps)
means, covs = ukf.run(len(dt), dt, track)            # one forward launch for the whole track
smoothed, covs_s = ukf.run_rts_smoother(track)       # one backward launch
print(f"ship {a.ship_id}: {track.z.shape[1]} fixes, {len(dt)} filter steps")
print("last filtered state [lon, lat, sog, cog]:", means[-1])
print("first smoothed state                    :", smoothed[0])
print("largest smoothed position variance      :", float(np.max(covs_s[:, :2, :2])))
