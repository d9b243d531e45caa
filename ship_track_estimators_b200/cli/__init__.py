"""``track_estimator`` command line front end (drop-in for reference ``track_estimators/cli``)."""
