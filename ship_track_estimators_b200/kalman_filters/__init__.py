"""Kalman filters: the reference's ``track_estimators.kalman_filters`` package surface."""
from .kalman_filter import KalmanFilterBase
from .non_linear_process import geodetic_dynamics, geodetic_dynamics_turn
from .unscented import UnscentedKalmanFilter

__all__ = ["KalmanFilterBase", "UnscentedKalmanFilter", "geodetic_dynamics", "geodetic_dynamics_turn"]
