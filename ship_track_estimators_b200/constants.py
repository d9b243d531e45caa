"""Physical constants shared by host code and (restated) by the CUDA kernels.

Mirrors reference ``src/track_estimators/constants.py:1``.
"""

EARTH_RADIUS = 6378.137  # Radius of the earth in km
