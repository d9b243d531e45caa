"""Developer tool: Jacobi sweeps per square root (per root, and the maximum over the 32 lanes of a warp, which is what
a warp executes) on the config-5 shape (`c5`) or the config-4 shape (`c4`), through the HOST build of the device code
(tools/host_emul compiled with -DSTE_EMUL_STATS); on `c5` also the relative size of the filtered covariances' off-diagonal
entries by pair.  usage: python tools/sweep_stats.py c5|c4"""
import os, sys, ctypes as C
os.environ["STE_EMUL_FLAGS"] = "-DSTE_EMUL_STATS"
REPO=os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, REPO); sys.path.insert(0, REPO+"/tests"); sys.path.insert(0, REPO+"/tools/host_emul")
import numpy as np, torch
import emul
from emul import HostUKF
from ship_track_estimators_b200.batch import TrackBatch
from ship_track_estimators_b200.synthetic import make_tracks
shape = sys.argv[1] if len(sys.argv) > 1 else "c5"
H = np.diag([1.0, 1, 0, 0]); R = np.diag([1e-3, 1e-3, 0, 0]); Q = np.diag([1e-2, 1e-2, 1e-4, 1e-4]); P = np.eye(4)
if shape == "c5":
    T, N, k = 256, 512, 1
    syn = make_tracks(T, N + 1, seed=11, device="cpu")
    u = HostUKF(H, Q, R, P, packed_cov=True, long_steps=False)
else:
    T, N, k = 256, 400, 2
    syn = make_tracks(T, 400, seed=5, device="cpu", nobs_min=399, dts_choices=(1, 2, 3, 6, 12, 24), outlier_frac=0.01, smooth_width=2)
    u = HostUKF(H, Q, R, P, gating=True, packed_cov=True, long_steps=True)
b = TrackBatch.from_synthetic(syn, substeps=k, need_rows=u.model.rows_needed())
res = u.allocate(b, smoother=False) if hasattr(u, "allocate") else None
lib = emul.build()
buf = np.zeros(T * (b.max_steps + 4) * 2, np.uint8)
lib.emul_sweep_log(buf.ctypes.data_as(C.c_void_p), C.c_longlong(buf.size))
u.forward(b, res)
lib.emul_sweep_log_count.restype = C.c_longlong
n = lib.emul_sweep_log_count()
steps = n // T
print("roots", n, "per track", n / T, "max_steps", b.max_steps)
log = buf[:steps * T].reshape(T, steps).astype(int)   # track-major call order
print("per-root hist", np.bincount(log.ravel(), minlength=5), "mean", log[log > 0].mean())
w = log.reshape(T // 32, 32, steps)
wmax = w.max(axis=1)
print("per-warp max: hist", np.bincount(wmax.ravel(), minlength=5), "mean", wmax.mean(), "cold in warp", (w.min(axis=1) == 0).mean())
print("mean per-root by step (first 12)", log.mean(axis=0)[:12], " steady", log[:, steps // 2:].mean(), " per-warp steady", wmax[:, steps // 2:].mean())
if shape == "c5":
    rel = []
    for t in range(0, T, 8):
        Pc = res.track(t)["covs"][:-1]
        d = np.sqrt(np.einsum("sii->si", Pc))
        rel.append(np.abs(Pc) / (d[:, :, None] * d[:, None, :]))
    rel = np.concatenate(rel)
    for (p, q) in [(0,1),(2,3),(0,2),(1,3),(0,3),(1,2)]:
        v = rel[:, p, q]
        print((p, q), "median %.2e  90%% %.2e  max %.2e  frac<1e-4 %.3f" % (np.median(v), np.quantile(v, .9), v.max(), (v < 1e-4).mean()))
