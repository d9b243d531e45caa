"""GPU parity against the reference-generated golden fixtures (tests/golden/*.npz), through the
C ABI.  Tolerance 1e-9 (BASELINE.json north_star); decisions (update mask, gating iterations) exact."""
from types import SimpleNamespace

import numpy as np
import pytest

from _helpers import TOL, assert_track_close, load_golden

pytestmark = pytest.mark.gpu


def _as_track(tr):
    return SimpleNamespace(dts=tr["dts"], z=tr["z"], sog_rate=tr["sog_rate"], cog_rate=tr["cog_rate"])


def _groups(tracks):
    """Tracks sharing (H, Q, R, P0, smoother, gating) run as one batch."""
    out = {}
    for i, tr in enumerate(tracks):
        key = (tr["H"].tobytes(), tr["Q"].tobytes(), tr["R"].tobytes(), tr["P0"].tobytes(), "means_s" in tr, "gate_iters" in tr)
        out.setdefault(key, []).append(i)
    return list(out.values())


def _run_group(tracks, idx, cuda, force_generic=False):
    from ship_track_estimators_b200.batch import BatchedUKF, TrackBatch

    t0 = tracks[idx[0]]
    smoother, gating = "means_s" in t0, "gate_iters" in t0
    noise = None
    if "noise_pred" in t0:
        noise = [dict(pred=tracks[i]["noise_pred"], upd=tracks[i]["noise_upd"], bwd=tracks[i].get("noise_bwd")) for i in idx]
    batch = TrackBatch.from_tracks(
        [_as_track(tracks[i]) for i in idx], [tracks[i]["dt_array"] for i in idx], device=cuda,
        x0=[tracks[i]["x0"] for i in idx], noise=noise, smoother=smoother,
    )
    ukf = BatchedUKF(t0["H"], t0["Q"], t0["R"], t0["P0"], gating=gating, force_generic=force_generic)
    res = ukf.run(batch, smoother=smoother)
    return [res.track(j) for j in range(len(idx))]


@pytest.mark.parametrize("force_generic", [False, True])
@pytest.mark.parametrize(
    "name",
    ["c1_single_ship", "c2_historical_batch", "c2_modern_ship", "c3_const_dt", "c4_ragged_ungated", "c4_ragged_gated", "tape_noise", "dense_h"],
)
def test_golden_fixture(name, force_generic, cuda, native_lib):
    tracks, _ = load_golden(name)
    for idx in _groups(tracks):
        got = _run_group(tracks, idx, cuda, force_generic)
        for g, i in zip(got, idx):
            ref = tracks[i]
            label = f"{name}[{i}] generic={force_generic}"
            assert g["n_updates"] == 1 + int(ref["mask"].sum()), label
            assert not (g["status"] & 0x1), f"{label}: non-finite state"
            if "gate_iters" in ref:
                assert np.array_equal(g["gate_iters"], ref["gate_iters"]), f"{label}: gating decisions differ"
                np.testing.assert_allclose(g["gate_lambda"], ref["gate_lambda"], rtol=1e-9, err_msg=label)
            assert_track_close(g, ref, tol=TOL, smoother="means_s" in ref, label=label)


def test_geodetic_known_answers(cuda, native_lib):
    from ship_track_estimators_b200.kalman_filters.non_linear_process import geodetic_dynamics

    d = np.load(__import__("os").path.join(__import__("_helpers").GOLDEN, "kat_blocks.npz"))
    got = geodetic_dynamics(d["geo_x"].T.copy(), None, d["geo_dt"], d["geo_sog_rate"], d["geo_cog_rate"]).T
    np.testing.assert_allclose(got, d["geo_out"], rtol=1e-13, atol=1e-12)
    # SURVEY section 9.3 known answer
    one = geodetic_dynamics(np.array([-30.5, -0.5, 14.5, 198.5]), None, 12.0, sog_rate=0.01, cog_rate=-0.02)
    np.testing.assert_allclose(one, [-30.99621056441692, -1.9822576603313298, 14.62, 198.26], rtol=1e-13)


def test_sigma_points_known_answers(cuda, native_lib):
    from ship_track_estimators_b200.kalman_filters.unscented import UnscentedKalmanFilter

    d = np.load(__import__("os").path.join(__import__("_helpers").GOLDEN, "kat_blocks.npz"))
    for x, P, X in zip(d["sp_x"], d["sp_P"], d["sp_X"]):
        u = UnscentedKalmanFilter(H=np.eye(4), P=P, x0=x)
        W = u.compute_weights()
        np.testing.assert_array_equal(W, d["weights"])
        got = u.compute_sigma_points()
        scale = np.max(np.abs(X - x[:, None]))
        assert np.max(np.abs(got - X)) <= 1e-10 * max(scale, 1.0)


def test_gating_known_answers(cuda, native_lib):
    """SURVEY section 9.3 / tests/golden/kat_gating.npz: iterations, final lambda, R scale."""
    from ship_track_estimators_b200.kalman_filters.unscented import UnscentedKalmanFilter

    rows = np.load(__import__("os").path.join(__import__("_helpers").GOLDEN, "kat_gating.npz"))["rows"]
    for d, iters, lam, scale in rows:
        g = UnscentedKalmanFilter(H=np.diag([1.0, 1, 0, 0]), R=np.diag([0.25, 0.25, 0, 0]), P=np.diag([0.3, 0.3, 1, 1]),
                                  x0=np.array([10.0, 20, 12, 90]), noise="zero")   # the known answers are noise-free
        Rs = g.check_robustness(np.array([10.0 + d, 20.0 - d, 12, 90]), g.P, g.R)
        assert g.last_gate["iterations"] == int(iters)
        assert abs(g.last_gate["lambda_factor"] - lam) <= 1e-9 * lam
        assert abs(Rs[0, 0] / 0.25 - scale) <= 1e-9 * scale


def _gate_kats():
    import os

    from _helpers import GOLDEN

    d = np.load(os.path.join(GOLDEN, "kat_gating_noisy.npz"))
    return [{k.split("_", 1)[1]: d[k] for k in d.files if k.startswith(f"c{i}_")} for i in range(int(d["n"]))]


def test_robustification_entry_points_against_reference_kats(cuda, native_lib):
    """criterion_index, update_lambda_factor and the reference's NOISY judging loop followed by the update
    (unscented.py:209-265, 353-483 with the call at :228 enabled), the recorded np.random.normal draws replayed."""
    from ship_track_estimators_b200.kalman_filters import UnscentedKalmanFilter, geodetic_dynamics

    Q = np.diag([1e-2, 1e-2, 1e-4, 1e-4])
    for i, c in enumerate(_gate_kats()):
        ukf = UnscentedKalmanFilter(H=c["H"], Q=Q, R=c["R"], P=c["P"], x0=c["x"], non_linear_process=geodetic_dynamics, gating=True)
        gamma = ukf.criterion_index(c["z"], c["P"], c["R"])
        np.testing.assert_allclose(gamma, float(c["gamma"]), rtol=1e-9, err_msg=f"case {i} criterion_index")
        lam = ukf.update_lambda_factor(1.0, float(c["gamma"]), 50.0, c["z"], c["P"], c["R"])
        np.testing.assert_allclose(lam, float(c["lam"]), rtol=1e-9, err_msg=f"case {i} update_lambda_factor")
        queue = [row for row in c["tape"]]
        orig = np.random.normal

        def replay(loc=0.0, scale=1.0, size=None):
            rows = size[0] if isinstance(size, tuple) and len(size) == 2 else 1
            out = np.stack([queue.pop(0) for _ in range(rows)])
            return out if isinstance(size, tuple) and len(size) == 2 else out[0]

        np.random.normal = replay
        try:
            ukf.update(c["z"])
        finally:
            np.random.normal = orig
        assert not queue, f"case {i}: {len(queue)} recorded draws were not consumed"
        assert ukf.last_gate["iterations"] == int(c["n_gate_draws"]) - 1
        np.testing.assert_allclose(ukf.last_gate["scale"], c["R_out"][0, 0] / c["R"][0, 0], rtol=1e-9)
        d = ukf.x.reshape(-1) - c["x_post"]
        d[3] = (d[3] + 180.0) % 360.0 - 180.0
        assert float(np.max(np.abs(d) / np.maximum(1.0, np.abs(c["x_post"])))) <= TOL, f"case {i} state"
        assert float(np.max(np.abs(ukf.P - c["P_post"])) / np.max(np.abs(c["P_post"]))) <= TOL, f"case {i} covariance"


def test_per_track_measurement_covariance(cuda, native_lib):
    """every track of a tile with its OWN dense 4x4 R (SteInputs.R_tracks), H = I, UKF + URTSS, against
    one reference run per track (tests/golden/per_track_r.npz); the filter's shared R must not matter."""
    from ship_track_estimators_b200.batch import BatchedUKF, TrackBatch

    tracks, _ = load_golden("per_track_r")
    assert not np.array_equal(tracks[0]["R"], tracks[1]["R"]) and np.count_nonzero(tracks[0]["R"]) == 16
    batch = TrackBatch.from_tracks([_as_track(t) for t in tracks], [t["dt_array"] for t in tracks], device=cuda,
                                   x0=[t["x0"] for t in tracks], R=[t["R"] for t in tracks])
    for shared in (np.eye(4) * 123.0, np.diag([1e-3, 1e-3, 0.0, 0.0])):   # the second would select the position-only update
        res = BatchedUKF(tracks[0]["H"], tracks[0]["Q"], shared, tracks[0]["P0"]).run(batch)
        for i, ref in enumerate(tracks):
            assert_track_close(res.track(i), ref, tol=TOL, label=f"per_track_r[{i}]")


def test_generic_dimension_class_api_against_reference_kats(cuda, native_lib):
    """the class at n = 5 (state [lon, lat, sog, cog, cog_rate], process geodetic_dynamics_turn):
    the process model, sigma points, predict and update against the reference's class run with the same five-state
    process built from its own geodetic_dynamics (tests/golden/kat_n5.npz)."""
    import os

    from _helpers import GOLDEN
    from ship_track_estimators_b200.kalman_filters import UnscentedKalmanFilter, geodetic_dynamics, geodetic_dynamics_turn

    d = np.load(os.path.join(GOLDEN, "kat_n5.npz"))
    H, Q, R = d["H"], d["Q"], d["R"]
    rel = lambda got, ref: float(np.max(np.abs(np.asarray(got) - ref)) / max(1.0, float(np.max(np.abs(ref)))))   # noqa: E731
    for i in range(int(d["n"])):
        c = {k.split("_", 1)[1]: d[k] for k in d.files if k.startswith(f"c{i}_")}
        dt, sr = float(c["dt"]), float(c["sog_rate"])
        assert rel(geodetic_dynamics_turn(c["x"], None, dt, sog_rate=sr), c["f_x"]) <= 1e-13
        ukf = UnscentedKalmanFilter(H=H, Q=Q, R=R, P=c["P"], x0=c["x"], non_linear_process=geodetic_dynamics_turn, noise="zero")
        assert ukf.n == 5 and ukf.n_sigma_points == 11
        ukf.compute_weights()
        assert rel(ukf.compute_sigma_points(), c["X0"]) <= 1e-11
        ukf.predict(dt=dt, c=None, sog_rate=sr)
        assert rel(ukf.sigma_points_orig, c["X0"]) <= 1e-11 and rel(ukf.sigma_points, c["X_pred"]) <= 1e-11
        assert rel(ukf.x.reshape(-1), c["x_pred"]) <= TOL and rel(ukf.P, c["P_pred"]) <= TOL, f"case {i} predict"
        ukf.update(c["z"])
        dx = ukf.x.reshape(-1) - c["x_post"]
        dx[3] = (dx[3] + 180.0) % 360.0 - 180.0
        assert float(np.max(np.abs(dx) / np.maximum(1.0, np.abs(c["x_post"])))) <= TOL and rel(ukf.P, c["P_post"]) <= TOL, f"case {i} update"
    with pytest.raises(ValueError, match="propagates"):
        UnscentedKalmanFilter(H=H, Q=Q, R=R, non_linear_process=geodetic_dynamics).predict(dt=1.0, c=None)
    with pytest.raises(NotImplementedError, match="arbitrary Python callables"):
        UnscentedKalmanFilter(H=H, Q=Q, R=R, non_linear_process=lambda x, **k: x).predict(dt=1.0, c=None)


def test_generic_dimension_run_and_smoother_against_the_reference(cuda, native_lib):
    """``run`` and ``run_rts_smoother`` of the class at n = 5 (geodetic_dynamics_turn) against the
    unmodified reference run with the same five-state process (tests/golden/kat_n5_tracks.npz: k = 1, 2, 2)."""
    import os
    from types import SimpleNamespace

    from _helpers import GOLDEN
    from ship_track_estimators_b200.kalman_filters import UnscentedKalmanFilter, geodetic_dynamics_turn

    d = np.load(os.path.join(GOLDEN, "kat_n5_tracks.npz"))
    H, Q, R, P0 = d["H"], d["Q"], d["R"], d["P0"]

    def errs(got_m, got_c, ref_m, ref_c):
        dm = np.asarray(got_m).reshape(ref_m.shape) - ref_m
        dm[:, 3] = (dm[:, 3] + 180.0) % 360.0 - 180.0
        em = float(np.max(np.abs(dm) / np.maximum(1.0, np.abs(ref_m))))
        ec = max(float(np.max(np.abs(g - r)) / np.max(np.abs(r))) for g, r in zip(np.asarray(got_c), ref_c))
        return em, ec

    for i in range(int(d["n"])):
        t = {k.split("_", 1)[1]: d[k] for k in d.files if k.startswith(f"t{i}_")}
        st = SimpleNamespace(dts=t["dts"], z=t["z"], sog=t["z"][2], cog=t["z"][3], sog_rate=t["sog_rate"].copy(), cog_rate=t["cog_rate"].copy())
        ukf = UnscentedKalmanFilter(H=H, Q=Q, R=R, P=P0, x0=t["x0"], non_linear_process=geodetic_dynamics_turn, noise="zero")
        m, c = ukf.run(len(t["dt_array"]), t["dt_array"], st)
        em, ec = errs(m, c, t["means"], t["covs"])
        assert m.shape == t["means"].shape and em <= TOL and ec <= TOL, (i, em, ec)
        ms, cs = ukf.run_rts_smoother(st)
        es, ecs = errs(ms, cs, t["means_s"], t["covs_s"])
        print(f"  n = 5 track {i}: filtered {em:.1e} {ec:.1e}  smoothed {es:.1e} {ecs:.1e}")
        assert ms.shape == t["means_s"].shape and es <= TOL and ecs <= TOL, (i, es, ecs)
        assert ukf.status == 0


def test_generic_dimension_run_loop(cuda, native_lib):
    """n = 5 through KalmanFilterBase.run (host loop around the generic single-step kernels): with a zero turn-rate
    state that nothing excites (zero process noise and prior variance on it) the first four states must reproduce
    the n = 4 filter of the batched path fed cog_rate = 0."""
    from types import SimpleNamespace

    from ship_track_estimators_b200.kalman_filters import UnscentedKalmanFilter, geodetic_dynamics, geodetic_dynamics_turn
    from ship_track_estimators_b200.synthetic import make_tracks
    from ship_track_estimators_b200.utils import generate_dts

    syn = make_tracks(1, 13, seed=5, device="cpu", dts_choices=(1.0, 2.0))
    z = np.stack([syn.lon[:, 0].numpy(), syn.lat[:, 0].numpy(), syn.sog[:, 0].numpy(), syn.cog[:, 0].numpy()])
    st = SimpleNamespace(dts=syn.dts[:, 0].numpy(), z=z, sog_rate=syn.sog_rate[:, 0].numpy(), cog_rate=np.zeros(13), sog=z[2], cog=z[3])
    dt = generate_dts(st.dts, 2)
    H4, Q4, R4 = np.diag([1.0, 1, 0, 0]), np.diag([1e-2, 1e-2, 1e-4, 1e-4]), np.diag([1e-3, 1e-3, 0, 0])
    u4 = UnscentedKalmanFilter(H=H4, Q=Q4, R=R4, P=np.eye(4), x0=z[:, 0], non_linear_process=geodetic_dynamics, noise="zero")
    m4, c4 = u4.run(len(dt), dt, st)
    pad = lambda M: np.pad(M, ((0, 1), (0, 1)))   # noqa: E731
    u5 = UnscentedKalmanFilter(H=pad(H4), Q=pad(Q4), R=pad(R4), P=pad(np.eye(4)), x0=np.append(z[:, 0], 0.0),
                               non_linear_process=geodetic_dynamics_turn, noise="zero")
    m5, c5 = u5.run(len(dt), dt, st)
    assert m5.shape == (len(dt) + 1, 5) and c5.shape == (len(dt) + 1, 5, 5)
    # with an inert fifth state the n = 5 transform IS the n = 4 one: the scale n / (1 - W0) = 3 and Wi = 1/6 are the same,
    # and the two extra sigma points sit on the mean, moving its weight from W0 = -2/3 back to -1/3
    d = m5[:, :4] - m4
    d[:, 3] = (d[:, 3] + 180.0) % 360.0 - 180.0
    assert np.allclose(m5[:, 4], 0.0) and float(np.max(np.abs(d) / np.maximum(1.0, np.abs(m4)))) <= 1e-8
    assert np.all(np.isfinite(c5)) and np.allclose(c5[:, 4, :], 0.0)
    assert float(np.max(np.abs(c5[:, :4, :4] - c4)) / np.max(np.abs(c4))) <= 1e-8


def test_process_model_near_the_poles(cuda, native_lib):
    """The reference takes the new latitude as arcsin(.), the CUDA path as atan2(up, hypot(east, north)): the same angle,
    conditioned differently as |lat| -> 90.  Reference known answers at |lat| = 89.5 ... 89.9999 degrees
    (tests/golden/kat_polar.npz): the process model to 1e-9 of max(1, |value|) and one unscented predict to 1e-9."""
    import os

    from _helpers import GOLDEN
    from ship_track_estimators_b200.kalman_filters import UnscentedKalmanFilter, geodetic_dynamics

    d = np.load(os.path.join(GOLDEN, "kat_polar.npz"))
    worst = 0.0
    for i in range(len(d["x"])):
        y = geodetic_dynamics(d["x"][i], None, float(d["dt"][i]), sog_rate=float(d["sog_rate"][i]), cog_rate=float(d["cog_rate"][i]))
        err = float(np.max(np.abs(y - d["y"][i]) / np.maximum(1.0, np.abs(d["y"][i]))))
        worst = max(worst, err)
        assert err <= TOL, (i, d["x"][i], y, d["y"][i])
        ukf = UnscentedKalmanFilter(H=np.diag([1.0, 1, 0, 0]), Q=d["Q"], R=np.eye(4), P=d["P0"][i], x0=d["x"][i],
                                    non_linear_process=geodetic_dynamics, noise="zero")
        ukf.predict(dt=float(d["dt"][i]), c=None, sog_rate=float(d["sog_rate"][i]), cog_rate=float(d["cog_rate"][i]))
        assert float(np.max(np.abs(ukf.x.reshape(-1) - d["x_pred"][i]) / np.maximum(1.0, np.abs(d["x_pred"][i])))) <= TOL, i
        assert float(np.max(np.abs(ukf.P - d["P_pred"][i])) / np.max(np.abs(d["P_pred"][i]))) <= 1e-8, i
    print(f"process model at |lat| up to 89.9999 deg: worst deviation from the reference {worst:.2e}")
