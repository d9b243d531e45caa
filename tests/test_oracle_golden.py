"""CPU: pin the oracle (oracle/ukf_numpy.py) against the reference-generated golden vectors.

The fixtures were produced by running the unmodified reference (tests/golden/make_golden.py);
the numpy oracle calls the same scipy/numpy routines in the same order, so agreement is expected
to the last few ulps (asserted at 1e-12, observed 0).  The reference's own unit-test identities
(reference tests/test_unscented_kf.py:24-87) are restated for the oracle as well.
"""
import os

import numpy as np
import pytest

from _helpers import GOLDEN, cov_err, load_golden, mean_err
from oracle import ukf_numpy as O

STRICT = 1e-12
FIXTURES = ["c1_single_ship", "c2_modern_ship", "c3_const_dt", "c4_ragged_ungated", "c4_ragged_gated", "tape_noise", "dense_h"]


def _oracle_run(tr):
    noise = None
    if "noise_pred" in tr:
        noise = O.TapeNoise(pred=tr["noise_pred"], upd=tr["noise_upd"], bwd=tr.get("noise_bwd"))
    return O.run_track(tr["x0"], tr["P0"], tr["H"], tr["Q"], tr["R"], tr["dt_array"], tr["dts"], tr["z"],
                       tr["sog_rate"], tr["cog_rate"], smoother="means_s" in tr, noise=noise, gating="gate_iters" in tr)


def _check(tr, out, label):
    assert np.array_equal(out["mask"], tr["mask"]), label
    assert mean_err(out["means"], tr["means"]) <= STRICT, label
    if "covs" in tr:
        assert cov_err(out["covs"], tr["covs"]) <= STRICT, label
    if "means_s" in tr:
        assert mean_err(out["means_s"], tr["means_s"]) <= STRICT, label
        if "covs_s" in tr:
            assert cov_err(out["covs_s"], tr["covs_s"]) <= STRICT, label
    if "gate_iters" in tr:
        assert np.array_equal(out["gate_iters"], tr["gate_iters"]), label
        np.testing.assert_allclose(out["gate_lambda"], tr["gate_lambda"], rtol=STRICT)


@pytest.mark.parametrize("name", FIXTURES)
def test_oracle_matches_reference_fixture(name):
    tracks, _ = load_golden(name)
    for i, tr in enumerate(tracks):
        _check(tr, _oracle_run(tr), f"{name}[{i}]")


def test_oracle_matches_reference_historical_batch_sample():
    """BASELINE config 2 (71 runnable historical ships): every 6th track keeps the CPU suite short."""
    tracks, _ = load_golden("c2_historical_batch")
    for i in range(0, len(tracks), 6):
        _check(tracks[i], _oracle_run(tracks[i]), f"c2_historical_batch[{i}]")


def test_building_block_known_answers():
    d = np.load(os.path.join(GOLDEN, "kat_blocks.npz"))
    for x, dt, sr, cr, ref in zip(d["geo_x"], d["geo_dt"], d["geo_sog_rate"], d["geo_cog_rate"], d["geo_out"]):
        np.testing.assert_allclose(O.geodetic_dynamics(x, dt, sr, cr), ref, rtol=1e-14, atol=1e-13)
    w0, wi = O.ut_weights(4)
    W = d["weights"]
    assert W[0, 0] == w0 and np.all(np.diag(W)[1:] == wi)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for x, P, X in zip(d["sp_x"], d["sp_P"], d["sp_X"]):
            np.testing.assert_allclose(O.sigma_points(x, P, w0), X, rtol=1e-13, atol=1e-13)
    # SURVEY section 9.3
    np.testing.assert_allclose(
        O.geodetic_dynamics(np.array([-30.5, -0.5, 14.5, 198.5]), 12.0, 0.01, -0.02),
        [-30.99621056441692, -1.9822576603313298, 14.62, 198.26], rtol=1e-15)
    assert w0 == -0.33333333333333326 and wi == 0.16666666666666666


def test_gating_known_answers():
    rows = np.load(os.path.join(GOLDEN, "kat_gating.npz"))["rows"]
    H, R, P = np.diag([1.0, 1, 0, 0]), np.diag([0.25, 0.25, 0, 0]), np.diag([0.3, 0.3, 1, 1])
    x = np.array([10.0, 20, 12, 90])
    for d, iters, lam, scale in rows:
        Rs, it, lm = O.check_robustness(x, x + np.array([d, -d, 0, 0]), P, H, R)
        assert it == int(iters)
        assert abs(lm - lam) <= 1e-12 * lam and abs(Rs[0, 0] / 0.25 - scale) <= 1e-12 * scale


def test_reference_unit_test_identities():
    """reference tests/test_unscented_kf.py: n = 2, sigma points recover (x, P) unweighted with
    W0 = 0 (:24-39) and weighted with the default weights (:64-87); weights sum to 1 (:42-61)."""
    rng = np.random.default_rng(7)
    P = np.diag(rng.uniform(0, 1, 2))
    x = rng.uniform(0, 1, 2)
    X = O.sigma_points(x, P, 0.0)
    assert np.allclose(x, X.mean(axis=1)) and np.allclose(P, np.cov(X))
    w0, wi = O.ut_weights(2)
    assert np.isclose(w0 + 4 * wi, 1.0) and -1 < w0 < 1
    X = O.sigma_points(x, P, w0)
    w = np.array([w0] + [wi] * 4)
    assert np.allclose(x, (X * w).sum(axis=1))
    dev = X - x[:, None]
    assert np.allclose(P, (dev * w) @ dev.T)


def test_update_mask_exact_equality():
    """kalman_filter.py:101: k = 2, 4 re-sum exactly on integer-hour gaps, k = 3 / 7 mostly miss."""
    dts = np.array([1.0, 2.0, 3.0, 6.0, 12.0, 24.0, 1.0, 3.0])
    for k, expect_all in ((1, True), (2, True), (4, True), (3, False), (7, False)):
        m = O.update_mask(O.generate_dts(dts, k), dts)
        hits = m.reshape(-1, k)[:, -1]
        assert (hits.all() and m.sum() == len(dts)) == expect_all


# --------------------------------------------------------------------------------------------- #
# the plain-C oracle (oracle/ukf_oracle.c): same fixtures, same bound as the CUDA path             #
# --------------------------------------------------------------------------------------------- #
ALL_FIXTURES = FIXTURES + ["c2_historical_batch"]


@pytest.mark.parametrize("name", ALL_FIXTURES)
def test_c_oracle_matches_reference_fixture(name):
    from _helpers import assert_track_close
    from oracle import ukf_c as OC

    tracks, _ = load_golden(name)
    for i, tr in enumerate(tracks):
        noise = dict(pred=tr["noise_pred"], upd=tr["noise_upd"], bwd=tr.get("noise_bwd")) if "noise_pred" in tr else None
        out = OC.run_track(tr["x0"], tr["P0"], tr["H"], tr["Q"], tr["R"], tr["dt_array"], tr["dts"], tr["z"], tr["sog_rate"],
                           tr["cog_rate"], smoother="means_s" in tr, noise=noise, gating="gate_iters" in tr)
        assert np.array_equal(out["mask"], tr["mask"])
        if "gate_iters" in tr:
            assert np.array_equal(out["gate_iters"], tr["gate_iters"]), f"{name}[{i}]"
            np.testing.assert_allclose(out["gate_lambda"], tr["gate_lambda"], rtol=1e-9)
        assert_track_close(out, tr, smoother="means_s" in tr, label=f"C oracle {name}[{i}]")


def test_c_oracle_batch_layout_matches_per_track_calls():
    """oracle_batch (the [plane][T] driver used for large parity runs and the CPU baseline) against
    run_track on the same synthetic tile."""
    from oracle import ukf_c as OC
    from ship_track_estimators_b200.synthetic import make_tracks

    T, nobs, k = 5, 33, 2
    syn = make_tracks(T, nobs, seed=12, device="cpu", dts_choices=(1.0, 2.0))
    H, R = np.diag([1.0, 1.0, 0.0, 0.0]), np.diag([1e-3, 1e-3, 0.0, 0.0])
    Q, P = np.diag([1e-2, 1e-2, 1e-4, 1e-4]), np.eye(4)
    dt = np.repeat(syn.dts.numpy() / k, k, axis=0)
    out = OC.run_batch(syn.x0().numpy(), dt, syn.lon.numpy(), syn.lat.numpy(), syn.sog_rate.numpy(), syn.cog_rate.numpy(), H, Q, R, P,
                       substeps=k, z_sog=syn.sog.numpy(), z_cog=syn.cog.numpy())
    for t in range(T):
        z = np.stack([syn.lon[:, t].numpy(), syn.lat[:, t].numpy(), syn.sog[:, t].numpy(), syn.cog[:, t].numpy()])
        one = OC.run_track(z[:, 0], P, H, Q, R, dt[:, t], syn.dts[:, t].numpy(), z, syn.sog_rate[:, t].numpy(), syn.cog_rate[:, t].numpy())
        assert np.array_equal(out["mean_s"][:, :, t], one["means_s"]) and np.array_equal(out["cov_f"][:, :, t].reshape(-1, 4, 4), one["covs"])
        ref = O.run_track(z[:, 0], P, H, Q, R, dt[:, t], syn.dts[:, t].numpy(), z, syn.sog_rate[:, t].numpy(), syn.cog_rate[:, t].numpy())
        assert mean_err(one["means_s"], ref["means_s"]) <= 1e-10 and cov_err(one["covs_s"], ref["covs_s"]) <= 1e-10


def test_metrics_against_reference_known_answers():
    """oracle rmse / cum_abs_diff / abs_diff and the package's own array helpers against values the
    reference's performance_metrics produced (tests/golden/make_golden.py metrics_kat)."""
    import os

    from oracle import ukf_numpy as O
    from ship_track_estimators_b200 import performance_metrics as PM

    kat = np.load(os.path.join(os.path.dirname(__file__), "golden", "kat_metrics.npz"))
    for i in range(int(kat["n"])):
        x, xref = kat[f"x{i}"], kat[f"xref{i}"]
        for mod in (O, PM):
            assert np.array_equal(mod.abs_diff(x, xref), kat[f"abs{i}"])
            assert np.array_equal(mod.cum_abs_diff(x, xref), kat[f"cum{i}"])
            assert float(mod.rmse(x, xref)) == float(kat[f"rmse{i}"])


def test_wgs84_leg_known_answers():
    """The WGS84 leg (Vincenty restatement of the geographiclib call) in the oracle and in the package:
    the reference's one exact vector and the vectors of its own unit tests (tests/test_utils.py)."""
    from oracle import ukf_numpy as O
    from ship_track_estimators_b200.utils import geographiclib_distance, geographiclib_heading

    for dist_fn, head_fn in ((lambda *a: O.wgs84_leg(*a)[0], lambda *a: O.wgs84_leg(*a)[1]), (geographiclib_distance, geographiclib_heading)):
        # examples/cli_example/output_01203823_predictions.txt:1: 24 h leg of ship 01203823
        assert abs(dist_fn(-30.5, -0.5, -31.5, -3.5) / 24.0 - 14.578418614021368) <= 1e-11 * 14.58
        assert abs(head_fn(-30.5, -0.5, -31.5, -3.5) - 198.52495095065817) <= 1e-9
        # tests/test_utils.py:36-60, 87-100
        assert np.isclose(dist_fn(-74.0060, 40.7128, -118.2437, 34.0522), 3933.96, rtol=1e-2)
        assert np.isclose(dist_fn(-9.13333, 38.7167, -8.6291, 41.1579), 273.59, rtol=1e-2)
        assert np.isclose(head_fn(-94.581213, 39.099912, -90.200203, 38.627089), 96.51, rtol=1e-3)
        assert dist_fn(12.3, 45.6, 12.3, 45.6) == 0.0 and head_fn(12.3, 45.6, 12.3, 45.6) == 0.0
    # equatorial and meridional legs against closed forms: a * dlon, and the meridian arc (series in n)
    s, az = O.wgs84_inverse(0.0, 10.0, 0.0, 11.0)
    assert abs(s - 6378137.0 * np.radians(1.0)) < 1e-6 and abs(az - 90.0) < 1e-12
    s, az = O.wgs84_inverse(-10.0, 5.0, 20.0, 5.0)
    n = (1 / 298.257223563) / (2 - 1 / 298.257223563)
    A0 = 6378137.0 / (1 + n)      # Helmert's meridian arc
    arc = lambda p: A0 * ((1 + n**2 / 4 + n**4 / 64) * p - 1.5 * (n - n**3 / 8) * np.sin(2 * p) + 15 / 16 * (n**2 - n**4 / 4) * np.sin(4 * p)
                          - 35 / 48 * n**3 * np.sin(6 * p) + 315 / 512 * n**4 * np.sin(8 * p))   # noqa: E731
    assert abs(s - (arc(np.radians(20.0)) - arc(np.radians(-10.0)))) < 1e-4 and abs(az) < 1e-12


def test_oracle_criterion_index_against_reference_kats():
    """The numpy oracle's judging index against the reference's criterion_index known answers."""
    import os

    from _helpers import GOLDEN
    from oracle import ukf_numpy as O

    d = np.load(os.path.join(GOLDEN, "kat_gating_noisy.npz"))
    for i in range(int(d["n"])):
        g = O.criterion_index(d[f"c{i}_x"], d[f"c{i}_z"], d[f"c{i}_P"], d[f"c{i}_H"], d[f"c{i}_R"])
        np.testing.assert_allclose(g, float(d[f"c{i}_gamma"]), rtol=1e-12)
