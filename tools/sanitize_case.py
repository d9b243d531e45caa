"""Small end-to-end case for compute-sanitizer (memcheck): ragged gated tracks through every kernel."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ship_track_estimators_b200.batch import BatchedUKF, TrackBatch
from ship_track_estimators_b200.derive import batch_from_fixes
from ship_track_estimators_b200.synthetic import make_tracks
from ship_track_estimators_b200.kalman_filters import UnscentedKalmanFilter, geodetic_dynamics
dev = torch.device("cuda:0")
H = np.diag([1.0, 1, 0, 0]); R = np.diag([0.05, 0.05, 0, 0]); Q = np.diag([1e-2, 1e-2, 1e-4, 1e-4]); P = np.eye(4)
syn = make_tracks(333, 40, seed=3, device="cpu", nobs_min=9, dts_choices=(1, 2, 3), outlier_frac=0.05, smooth_width=2)
for kw in (dict(), dict(gating=True), dict(gating=True, force_generic=True), dict(packed_cov=True)):
    ukf = BatchedUKF(H, Q, R, P, **kw)
    b = TrackBatch.from_synthetic(syn, substeps=2, need_rows=ukf.model.rows_needed()).to(dev)
    r = ukf.run(b)
    r2 = ukf.run(b, res=ukf.allocate(b, reuse_stats=False, in_place=True))
    torch.cuda.synchronize()
    print(kw, int((r.status & 1).sum()), float(r.mean_s[0, 0, 0]), float(r2.mean_s[0, 0, 0]))
b2 = batch_from_fixes(syn.lon.to(dev), syn.lat.to(dev), syn.dts.to(dev), syn.nobs.to(dev), substeps=1, smooth_width=3)
print(float(BatchedUKF(H, Q, R, P).run(b2).mean_s[1, 1, 1]))
u = UnscentedKalmanFilter(H=H, Q=Q, R=R, P=P, x0=np.array([1.0, 2, 10, 45]), non_linear_process=geodetic_dynamics, noise="zero")
u.predict(dt=1.0, c=None, sog_rate=0.0, cog_rate=0.0); u.update(np.array([1.1, 2.1, 0, 0])); u.compute_sigma_points()
print(u.x.ravel())
