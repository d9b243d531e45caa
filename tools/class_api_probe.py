"""Dev tool: cost of the reference-style single-step calls (one track, one step per call) on the GPU."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ship_track_estimators_b200.kalman_filters import UnscentedKalmanFilter, geodetic_dynamics
H = np.diag([1.0, 1, 0, 0]); R = np.diag([1e-3, 1e-3, 0, 0]); Q = np.diag([1e-2, 1e-2, 1e-4, 1e-4])
ukf = UnscentedKalmanFilter(H=H, Q=Q, R=R, P=np.eye(4), x0=np.array([10.0, 20.0, 15.0, 90.0]), non_linear_process=geodetic_dynamics, noise="zero")
z = np.array([10.1, 20.05, 15.0, 90.0])
for _ in range(20):
    ukf.predict(dt=1.0, c=None, sog_rate=0.0, cog_rate=0.0); ukf.update(z)
torch.cuda.synchronize(); n = 500; t0 = time.perf_counter()
for _ in range(n):
    ukf.predict(dt=1.0, c=None, sog_rate=0.0, cog_rate=0.0); ukf.update(z)
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / n
print(json.dumps({"single_step_predict_plus_update_us": dt * 1e6, "note": "class API, batch of one, one H2D + launch + D2H per call; the reference's numpy step is ~430 us"}))
