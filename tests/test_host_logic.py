"""CPU: host-side logic of the drop-in API (input shaping, packing, error behaviour).

Known answers for the utilities are the reference's own (reference tests/test_utils.py:11-149)."""
from types import SimpleNamespace

import numpy as np
import pytest

from ship_track_estimators_b200.batch import FilterModel, TrackBatch, exact_update_mask
from ship_track_estimators_b200.utils import generate_dts, haversine_formula, heading, smooth


def test_generate_dts_reference_vectors():
    dt = [10, 5, 3, 2, 1, 10, 10]
    assert np.array_equal(generate_dts(dt, 2), np.array([5, 5, 2.5, 2.5, 1.5, 1.5, 1, 1, 0.5, 0.5, 5, 5, 5, 5]))
    ref4 = np.repeat(np.array([2.5, 1.25, 0.75, 0.5, 0.25, 2.5, 2.5]), 4)
    assert np.array_equal(generate_dts(dt, 4), ref4)
    assert generate_dts([], 3).shape == (0,)
    # bitwise the same as the reference's per-element Python loop
    rng = np.random.default_rng(0)
    dts = rng.uniform(0.1, 30, 50)
    loop = np.asarray([t / 7 for t in dts for _ in range(7)])
    assert np.array_equal(generate_dts(dts, 7), loop)


def test_haversine_and_heading_reference_answers():
    assert np.isclose(haversine_formula(-74.0060, 40.7128, -118.2437, 34.0522), 3933.96, rtol=1e-2)
    assert np.isclose(haversine_formula(-9.13333, 38.7167, -8.6291, 41.1579), 273.59, rtol=1e-2)
    assert np.isclose(haversine_formula(12.3, 45.6, 12.3, 45.6), 0.0)
    assert np.isclose(heading(-94.581213, 39.099912, -90.200203, 38.627089), 96.51, rtol=1e-3)
    assert np.isclose(heading(5.0, 5.0, 5.0, 5.0), 0.0)
    # derived quantities of the 5-fix track
    lon = np.array([-30.5, -31.5, -32.5, -33.5, -33.5]); lat = np.array([-0.5, -3.5, -6.5, -8.5, -11.5])
    dts = np.array([24.0, 24, 24, 12])
    sog = haversine_formula(lon[:-1], lat[:-1], lon[1:], lat[1:]) / dts
    cog = heading(lon[:-1], lat[:-1], lon[1:], lat[1:])
    np.testing.assert_allclose(sog, [14.666569517825678, 14.66188856000385, 10.35378642995031, 27.829872698318386], rtol=1e-14)
    np.testing.assert_allclose(cog, [198.40941994487474, 198.32831442189988, 206.30510826660182, 180], rtol=1e-14)


def test_smooth_matches_numpy_same_mode():
    y = np.arange(10, dtype=float) ** 2
    for w in (2, 3, 5):
        assert np.array_equal(smooth(y, w), np.convolve(y, np.ones(w) / w, mode="same"))


def test_exact_update_mask_against_python_loop():
    rng = np.random.default_rng(1)
    for k in (1, 2, 3, 4, 6, 7, 10):
        dts = rng.choice([1.0, 2.0, 3.0, 5.0, 12.0, 48.0], 40)
        dta = generate_dts(dts, k)
        t, cs, loop = 0, np.cumsum(dts), []
        for d in dta:  # reference kalman_filter.py:98-101 verbatim semantics
            t += d
            loop.append(t in cs)
        assert np.array_equal(exact_update_mask(dta, dts), np.asarray(loop)), k
    # a filter that has already advanced (self.time != 0)
    dts = np.array([1.0, 1.0, 1.0, 1.0])
    assert exact_update_mask(np.array([1.0, 1.0]), dts, time0=2.0).tolist() == [True, True]
    assert exact_update_mask(np.array([0.5, 0.5]), dts, time0=0.25).tolist() == [False, False]


def _track(nobs, dts=None, seed=0):
    rng = np.random.default_rng(seed)
    dts = np.ones(nobs - 1) if dts is None else np.asarray(dts, dtype=float)
    return SimpleNamespace(dts=dts, z=rng.normal(size=(4, nobs)), sog_rate=rng.normal(size=nobs), cog_rate=rng.normal(size=nobs))


def test_from_tracks_packing_cpu():
    a, b = _track(5, seed=1), _track(3, dts=[2.0, 4.0], seed=2)
    batch = TrackBatch.from_tracks([a, b], [generate_dts(a.dts, 2), generate_dts(b.dts, 2)], device="cpu")
    assert batch.n_tracks == 2 and batch.max_steps == 8 and batch.max_obs == 5
    assert batch.n_steps.tolist() == [8, 4] and batch.rate_repeat.tolist() == [2, 2]
    assert batch.upd_mask[:, 0].tolist() == [0, 1] * 4 and batch.upd_mask[:, 1].tolist() == [0, 1, 0, 1, 0, 0, 0, 0]
    assert np.array_equal(batch.z[1][:3, 1].numpy(), b.z[1]) and float(batch.z[1][3, 1]) == 0.0
    assert np.array_equal(batch.x0[:, 0].numpy(), a.z[:, 0])
    assert batch.dt[:, 1].tolist() == [1.0, 1.0, 2.0, 2.0, 0, 0, 0, 0]
    assert batch.track_steps() == 12
    # nobs == 2: the reference's repeat factor is k + 1 (int((k + 1) / 1)), SURVEY 7.3 item 9
    c = _track(2, dts=[3.0])
    assert TrackBatch.from_tracks([c], [generate_dts(c.dts, 2)], device="cpu").rate_repeat.tolist() == [3]


def test_from_tracks_raises_where_the_reference_does():
    # zero gap: the time matches twice -> update index runs past the observations (IndexError in
    # the reference, kalman_filter.py:101-108)
    t = _track(3, dts=[0.0, 1.0])
    with pytest.raises(IndexError):
        TrackBatch.from_tracks([t], [generate_dts(t.dts, 2)], device="cpu")
    # 3-row measurement matrices are not the 4-state filter's input
    t = _track(4)
    t.z = t.z[:3]
    with pytest.raises(NotImplementedError):
        TrackBatch.from_tracks([t], [generate_dts(t.dts, 1)], device="cpu")


def test_filter_model_rows_and_validation():
    H, R = np.diag([1.0, 1, 0, 0]), np.diag([1e-3, 1e-3, 0, 0])
    m = FilterModel(H, np.eye(4), R, np.eye(4))
    assert m.rows_needed() == [True, True, False, False]
    assert FilterModel(H, np.eye(4), R, np.eye(4), force_generic=True).rows_needed() == [True, True, False, False]
    assert FilterModel(np.eye(4), np.eye(4), np.eye(4), np.eye(4)).rows_needed() == [True] * 4
    assert FilterModel(H, np.eye(4), R, np.eye(4), gating=True, force_generic=True).rows_needed() == [True] * 4
    with pytest.raises(NotImplementedError):
        FilterModel(np.eye(2), np.eye(2), np.eye(2), np.eye(2))
    Q = np.eye(4); Q[0, 1] = 0.1
    with pytest.raises(NotImplementedError):
        FilterModel(H, Q, R, np.eye(4))


def test_reference_error_conventions():
    from ship_track_estimators_b200.kalman_filters import KalmanFilterBase, UnscentedKalmanFilter, geodetic_dynamics

    with pytest.raises(ValueError, match="Set proper system dynamics."):
        UnscentedKalmanFilter()
    u = UnscentedKalmanFilter(H=np.eye(4))
    assert u.n == 4 and u.n_sigma_points == 9 and u.weights.shape == (9, 9) and u.sigma_points.shape == (4, 9)
    assert np.array_equal(u.Q, np.eye(4)) and np.array_equal(u.P_orig, np.eye(4)) and u.x.shape == (4, 1)
    assert u.time == 0 and u.means == [] and u.covariances == [] and u.dt is None and u.c is None
    with pytest.raises(AssertionError):
        u.predict(dt=1.0)  # no process model set
    with pytest.raises(NotImplementedError):
        u.predict(lambda x, **k: x, dt=1.0)  # arbitrary callables cannot run inside the kernel
    with pytest.raises(AssertionError):
        UnscentedKalmanFilter(H=np.eye(4), non_linear_process=geodetic_dynamics).run(3, [1.0, 1.0], None)
    with pytest.raises(AssertionError):
        u.compute_weights(1.5)
    W = UnscentedKalmanFilter(H=np.eye(2)).compute_weights()
    assert np.isclose(np.trace(W), 1.0) and np.count_nonzero(W - np.diag(np.diagonal(W))) == 0
    with pytest.raises(NotImplementedError):
        KalmanFilterBase().predict()
    with pytest.raises(NotImplementedError):
        KalmanFilterBase().update()


def test_synthetic_generator_definitions():
    """Derived inputs follow ShipTrack's definitions (ship_track.py:197-304) on every track."""
    from ship_track_estimators_b200.synthetic import make_tracks

    syn = make_tracks(5, 20, seed=11, nobs_min=8, dts_choices=(1, 2, 6), smooth_width=0)
    assert syn.nobs.tolist() == sorted(syn.nobs.tolist(), reverse=True)
    for t in range(5):
        m = int(syn.nobs[t])
        lon, lat, dts = syn.lon[:m, t].numpy(), syn.lat[:m, t].numpy(), syn.dts[: m - 1, t].numpy()
        sog = np.append(haversine_formula(lon[:-1], lat[:-1], lon[1:], lat[1:]) / dts, 0.0); sog[-1] = sog[-2]
        cog = np.append(heading(lon[:-1], lat[:-1], lon[1:], lat[1:]), 0.0); cog[-1] = cog[-2]
        np.testing.assert_allclose(syn.sog[:m, t].numpy(), sog, rtol=1e-9)
        np.testing.assert_allclose(syn.cog[:m, t].numpy(), cog, rtol=1e-9, atol=1e-9)
        rate = np.append(0.0, np.diff(syn.sog[:m, t].numpy()) / dts)
        np.testing.assert_allclose(syn.sog_rate[:m, t].numpy(), rate, rtol=1e-12, atol=1e-15)
    # box smoothing as the CLI applies it (main_cli.py:99-104) on ragged columns
    syn2 = make_tracks(3, 12, seed=11, nobs_min=6, smooth_width=2)
    raw = make_tracks(3, 12, seed=11, nobs_min=6, smooth_width=0)
    for t in range(3):
        m = int(raw.nobs[t])
        np.testing.assert_allclose(syn2.sog[:m, t].numpy(), smooth(raw.sog[:m, t].numpy(), 2), rtol=1e-14)
    # determinism
    again = make_tracks(5, 20, seed=11, nobs_min=8, dts_choices=(1, 2, 6))
    assert np.array_equal(again.lon.numpy(), syn.lon.numpy())


def test_long_step_fraction_flags_mixed_tiles():
    """TrackBatch.long_step_fraction: ~0 for hourly 20 km/h tracks, well inside (0, 1) for gaps of
    1-24 h (the tiles BatchedUKF(long_steps=True) is meant for)."""
    from ship_track_estimators_b200.batch import TrackBatch
    from ship_track_estimators_b200.synthetic import make_tracks

    dense = TrackBatch.from_synthetic(make_tracks(64, 60, seed=1, device="cpu"), 1)
    mixed = TrackBatch.from_synthetic(make_tracks(64, 120, seed=2, device="cpu", nobs_min=30, dts_choices=(1, 2, 3, 6, 12, 24)), 2)
    assert dense.long_step_fraction() < 0.02
    assert 0.2 < mixed.long_step_fraction() < 0.8


def test_bench_reference_arm_line_for_the_real_data_config():
    """`bench.py --impl reference --config c2`: the reference (or, where it is not installed, the numpy port) over the
    fleet of the committed real-data fixtures on the host cores; one JSON line with the contract's keys and the GPU arm's
    config block."""
    import json
    import os
    import subprocess
    import sys

    repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(repo, "bench.py"), "--impl", "reference", "--config", "c2", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=900, cwd=repo)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
    assert line["impl"] == "reference" and line["unit"] == "track-steps/s" and line["value"] > 0
    assert line["config"]["config"] == "c2" and line["config"]["job_tracks"] == 72 and line["config"]["fixtures"] == ["c2_historical_batch", "c2_modern_ship"]
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["gpu_launches"] == 0
