"""``track_estimator``: filter (and optionally smooth) one ship of a CSV file on the GPU.

Same flags, ``input.json`` keys and output files as the reference (``cli/main_cli.py:54-169``):
``<prefix>_<id>_predictions.txt`` (N+1, 4), ``_variances.txt`` (diagonals), ``_dts.txt``,
``original_<id>_track.txt`` and, with ``-rts``, the two ``*_smoothed.txt`` files.
"""
from __future__ import annotations

import logging
import os
import sys
from typing import Optional, Tuple

import numpy as np

from .. import __version__
from ..kalman_filters.non_linear_process import geodetic_dynamics
from ..kalman_filters.unscented import UnscentedKalmanFilter
from ..ship_track import ShipTrack
from ..utils import generate_dts, smooth
from .argument_parser import create_parser
from .writers import write_track_outputs
from .json_loader import load_input_json

logger = logging.getLogger(__name__)

_BOAT = r"""
             /|~~~

           ///|

         /////|

       ///////|

     /////////|

   \==========|===/
~~~~~~~~~~~~~~~~~~~~~
"""


def start_banner():
    logger.info(_BOAT, extra={"simple": True})
    logger.info(f"version: {__version__}", extra={"simple": True})


def exit_banner():
    logger.info("Track estimator has terminated succesfully! :)", extra={"simple": True})


def _get_input_matrix(settings: dict, matrix_name: str, dim: int) -> np.ndarray:
    """A ``dim`` vector becomes a diagonal matrix, a ``dim x dim`` list is taken as is
    (reference ``main_cli.py:222-258``)."""
    if matrix_name not in settings:
        raise KeyError(f"{matrix_name} not found in input settings")
    matrix = np.asarray(settings[matrix_name])
    assert matrix.shape[0] == dim, f"Dimension mismatch: {matrix.shape[0]} != {dim} for {matrix_name}"
    if matrix.ndim == 1:
        return np.diag(matrix)
    if matrix.ndim == 2:
        assert matrix.shape[1] == dim, f"Dimension mismatch: {matrix.shape[1]} != {dim} for {matrix_name}"
        return matrix
    raise ValueError(f"{matrix_name} must be 1 or 2 dimensional")


def get_input_settings(settings: dict) -> Tuple[int, float, int, np.ndarray, np.ndarray, np.ndarray, np.ndarray, Optional[int]]:
    """``dim, dt, nsteps, H, Q, R, P, smooth`` from the JSON dictionary (``main_cli.py:172-219``)."""
    for key in ("dim", "dt", "nsteps"):
        if key not in settings:
            raise KeyError(f"{key} not found in input settings")
    dim, dt, nsteps = int(settings["dim"]), settings["dt"], int(settings["nsteps"])
    smooth_control = int(settings["smooth"]) if "smooth" in settings else None
    H, Q, R, P = (_get_input_matrix(settings, name, dim) for name in ("H", "Q", "R", "P"))
    return dim, dt, nsteps, H, Q, R, P, smooth_control


def track_estimator(argv=None):
    """Entry point of the ``track_estimator`` console script."""
    logging.basicConfig(format="Track estimator | %(levelname)s | %(asctime)s | %(message)s", level=logging.INFO,
                        datefmt="%Y-%m-%d %H:%M:%S", stream=sys.stdout)
    start_banner()
    args = create_parser().parse_args(argv)
    for path, what in ((args.input_file, "Input"), (args.track_file, "Track")):
        if not os.path.isfile(path):
            logger.error(f"{what} file '{path}' does not exist.")
            exit_banner()
            return

    logger.info(f"Reading input JSON from '{args.input_file}'...")
    dim, dt, nsteps, H, Q, R, P, smooth_control = get_input_settings(load_input_json(args.input_file))

    ship_track = ShipTrack()
    ship_track.read_csv(args.track_file, ship_id=args.ship_id, id_col=args.id_col, lat_col=args.lat_id,
                        lon_col=args.lon_id, reverse=bool(args.reverse))
    if smooth_control not in [-1, 0, 1, None]:
        logger.info(f"Smoothing SOG and COG by {smooth_control}.")
        ship_track.calculate_cog()
        ship_track.calculate_sog()
        ship_track.sog = smooth(ship_track.sog, smooth_control)
        ship_track.cog = smooth(ship_track.cog, smooth_control)
    z = ship_track.get_measurements(include_sog=True, include_cog=True)
    ship_track.calculate_cog_rate()
    ship_track.calculate_sog_rate()
    x0 = z[:, 0].reshape(-1, 1).copy()

    # a JSON dt other than -1 / 0 / null means "one step per observation gap" (main_cli.py:114-117)
    dt_array = generate_dts(ship_track.dts, nsteps if dt in [-1, 0, None] else 1)
    nsteps = len(dt_array)

    logger.info("Running the Unscented Kalman Filter.")
    ukf = UnscentedKalmanFilter(H=H, Q=Q, R=R, P=P, x0=x0, non_linear_process=geodetic_dynamics)
    predictions, estimate_vars = ukf.run(nsteps, dt_array, ship_track)
    logger.info("Finished running the Unscented Kalman Filter.")
    if args.apply_rts_smoother:
        predictions_smoothed, estimate_vars_smoothed = ukf.run_rts_smoother(ship_track=ship_track)

    logger.info(f"Writing outputs with prefix '{args.output_prefix}'.")
    write_track_outputs(args.output_prefix, args.ship_id, predictions, estimate_vars, dt_array, ship_track.lon, ship_track.lat,
                        predictions_smoothed if args.apply_rts_smoother else None,
                        estimate_vars_smoothed if args.apply_rts_smoother else None)
    exit_banner()
