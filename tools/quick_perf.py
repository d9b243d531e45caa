"""Quick device timing of the forward / backward kernels and the FP64 FMA probe (dev tool)."""
import argparse, json, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ship_track_estimators_b200 import _native as nat
from ship_track_estimators_b200.batch import BatchedUKF, TrackBatch
from ship_track_estimators_b200.synthetic import make_tracks

ap = argparse.ArgumentParser()
ap.add_argument("--tracks", type=int, default=148 * 128 * 4)
ap.add_argument("--steps", type=int, default=256)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--generic", action="store_true")
ap.add_argument("--packed", action="store_true")
ap.add_argument("--fused", action="store_true")
ap.add_argument("--fused-fwd-tracks", type=int, default=0, help="tile size of the forward side of the fused launch (default: --tracks)")
ap.add_argument("--fused-bwd-tracks", type=int, default=0, help="tile size of the backward side of the fused launch (default: --tracks)")
ap.add_argument("--two-streams", action="store_true", help="forward of one tile and backward of another launched on two streams")
ap.add_argument("--no-probe", action="store_true")
ap.add_argument("--no-metrics", action="store_true")
ap.add_argument("--label", default="")
a = ap.parse_args()
dev = torch.device("cuda:0")
lib = nat.load()

def timed(fn, reps):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts), ts

if not a.no_probe:
    blocks, threads, iters = 148 * 16, 256, 20000
    sink = torch.empty(blocks * threads, dtype=torch.float64, device=dev)
    ms, _ = timed(lambda: nat.check(lib.ste_probe_fp64_fma(blocks, threads, iters, nat.ptr(sink), nat.current_stream())), 3)
    print(json.dumps({"probe": "fp64_fma", "ms": ms, "tflops": 2.0 * 8 * iters * blocks * threads / ms / 1e9}))

syn = make_tracks(a.tracks, a.steps + 1, seed=1, device="cuda:0")
batch = TrackBatch.from_synthetic(syn, substeps=1)
H = np.diag([1.0, 1, 0, 0]); R = np.diag([1e-3, 1e-3, 0, 0]); Q = np.diag([1e-2, 1e-2, 1e-4, 1e-4]); P = np.eye(4)
ukf = BatchedUKF(H, Q, R, P, force_generic=a.generic, packed_cov=a.packed, long_steps=False)
res = ukf.allocate(batch, smoother=True)
res_nt = ukf.allocate(batch, smoother=False)
f_ms, f_all = timed(lambda: ukf.forward(batch, res), a.reps)
b_ms, b_all = timed(lambda: ukf.backward(batch, res), a.reps)
n_ms, n_all = timed(lambda: ukf.forward(batch, res_nt), a.reps)
ts = a.tracks * a.steps
print(json.dumps({"label": a.label or os.environ.get("STE_UKF_LIB", "default"), "tracks": a.tracks, "steps": a.steps, "fwd_ms": f_ms, "bwd_ms": b_ms, "fwd_no_tape_ms": n_ms,
                  "fwd_steps_per_s": ts / f_ms * 1e3, "bwd_steps_per_s": ts / b_ms * 1e3, "fwd_no_tape_steps_per_s": ts / n_ms * 1e3,
                  "both_steps_per_s": ts / (f_ms + b_ms) * 1e3,
                  "fwd_GBs": ts * 200 / f_ms / 1e6, "both_GBs": ts * 544 / (f_ms + b_ms) / 1e6,
                  "status_nonzero": int((res.status != 0).sum().item()), "all": [f_all, b_all, n_all]}))
print("sample smoothed", res.mean_s[0, :, 0].tolist(), res.mean_s[-1, :, 0].tolist())

if not a.no_metrics:
    from ship_track_estimators_b200.performance_metrics import track_metrics
    m_ms, _ = timed(lambda: track_metrics(ukf, batch, res, which="smoothed"), a.reps)
    print(json.dumps({"track_metrics_ms": m_ms, "GBs_algorithmic": ts * 32 / m_ms / 1e6}))

if a.fused:
    nf, nb = a.fused_fwd_tracks or a.tracks, a.fused_bwd_tracks or a.tracks
    batch2 = TrackBatch.from_synthetic(make_tracks(nf, a.steps + 1, seed=2, device="cuda:0"), substeps=1)
    res2 = ukf.allocate(batch2, smoother=True)
    if nb != a.tracks:
        batch = TrackBatch.from_synthetic(make_tracks(nb, a.steps + 1, seed=1, device="cuda:0"), substeps=1)
        res = ukf.allocate(batch, smoother=True)
    ukf.forward(batch, res)
    fu_ms, fu_all = timed(lambda: ukf.fused(batch2, res2, batch, res), a.reps)
    print(json.dumps({"fused_ms": fu_ms, "fwd_tracks": nf, "bwd_tracks": nb, "separate_ms": f_ms + b_ms, "fused_steps_per_s": ts / fu_ms * 1e3,
                      "gain": (f_ms + b_ms) / fu_ms, "all": fu_all}))

if a.two_streams:
    batch2 = TrackBatch.from_synthetic(make_tracks(a.tracks, a.steps + 1, seed=2, device="cuda:0"), substeps=1)
    res2 = ukf.allocate(batch2, smoother=True)
    ukf.forward(batch, res)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream(priority=-1)
    def both(order):
        e0 = torch.cuda.Event(); e0.record()
        s1.wait_event(e0); s2.wait_event(e0)
        def f():
            with torch.cuda.stream(s1): ukf.forward(batch2, res2)
        def b():
            with torch.cuda.stream(s2): ukf.backward(batch, res)
        (f(), b()) if order == "fb" else (b(), f())
        torch.cuda.current_stream().wait_stream(s1); torch.cuda.current_stream().wait_stream(s2)
    for order in ("fb", "bf"):
        ms, allv = timed(lambda: both(order), a.reps)
        print(json.dumps({"two_streams": order, "ms": ms, "separate_ms": f_ms + b_ms, "gain": (f_ms + b_ms) / ms, "all": allv}))
