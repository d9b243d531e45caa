"""Throughput of the other BASELINE configs' shapes on one GPU (dev tool; numbers quoted in DESIGN.md):
C3 forward only, k = 1;  C4 ragged tracks, k = 2, box smoothing 2, outliers + Mahalanobis gating, URTSS."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ship_track_estimators_b200.batch import BatchedUKF, TrackBatch
from ship_track_estimators_b200.synthetic import make_tracks
dev = "cuda:0"
H = np.diag([1.0, 1, 0, 0]); Q = np.diag([1e-2, 1e-2, 1e-4, 1e-4]); P = np.eye(4)
def timed(fn, reps=2):
    fn(); torch.cuda.synchronize(); best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    return best
# C3: forward only
T, N = 148 * 128 * 8, 1024
b = TrackBatch.from_synthetic(make_tracks(T, N + 1, seed=1, device=dev), 1)
ukf = BatchedUKF(H, Q, np.diag([1e-3, 1e-3, 0, 0]), P, packed_cov=True)
res = ukf.allocate(b, smoother=False)
ms = timed(lambda: ukf.forward(b, res))
print(json.dumps({"config": "C3 shape: forward only, k=1, 1024 steps", "tracks": T, "ms": ms, "track_steps_per_s": T * N / ms * 1e3, "flagged": int((res.status != 0).sum())}))
del b, res; torch.cuda.empty_cache()
# C4: ragged, gated
T, nmax, nmin = 148 * 128 * 2, 1200, 100
syn = make_tracks(T, nmax, seed=2, device=dev, nobs_min=nmin, dts_choices=(1, 2, 3, 6, 12, 24), outlier_frac=0.01, smooth_width=2)
for gating, long_steps in ((False, True), (True, True), (False, False)):
    # gaps of 1-24 h at k = 2 mix <= 50 km and longer steps inside every warp: long_steps=True keeps one geodetic tier
    ukf = BatchedUKF(H, Q, np.diag([0.05, 0.05, 0, 0]), P, gating=gating, packed_cov=True, long_steps=long_steps)
    b = TrackBatch.from_synthetic(syn, 2, need_rows=ukf.model.rows_needed())
    res = ukf.allocate(b, smoother=True)
    steps = b.track_steps()
    f = timed(lambda: ukf.forward(b, res)); bw = timed(lambda: ukf.backward(b, res))
    out = {"config": f"C4 shape: ragged {nmin}-{nmax} fixes, k=2, smooth 2, 1% outliers, gating={gating}, long_steps={long_steps}, URTSS", "tracks": T, "track_steps": steps,
           "fwd_ms": f, "bwd_ms": bw, "track_steps_per_s": steps / (f + bw) * 1e3, "nonfinite": int((res.status & 1).ne(0).sum()),
           "recompute_flag": int((res.status & 0x100).ne(0).sum())}
    if gating:
        out["gated_updates"] = int((res.gate_iters > 0).sum()); out["max_iters"] = int(res.gate_iters.max())
    print(json.dumps(out))
    del b, res; torch.cuda.empty_cache()
