"""Developer tool: BASELINE config-4-shaped tracks (ragged 100-5000 fixes, gaps 1-24 h, k = 2, box smoothing 2,
1 % displaced fixes, Mahalanobis gating, URTSS) through the HOST build of the device code (tools/host_emul) against
the plain-C oracle and its FMA-contracted rounding variant.  Prints, per track, the error against the oracle and the
oracle's own self-uncertainty - how the tolerance policy of tests/test_gpu_parity.py::test_c4_shape_* was chosen."""
import os, sys, time
from concurrent.futures import ThreadPoolExecutor
HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "tests")); sys.path.insert(0, os.path.join(HERE, "host_emul"))
import numpy as np, torch
from _helpers import track_errors
from emul import HostUKF
from oracle import ukf_c as OC, ukf_numpy as O
from ship_track_estimators_b200.batch import TrackBatch
from ship_track_estimators_b200.synthetic import make_tracks

T = int(sys.argv[1]) if len(sys.argv) > 1 else 32
nmax = int(sys.argv[2]) if len(sys.argv) > 2 else 5000
gating = (sys.argv[3] != "0") if len(sys.argv) > 3 else True
k = 2
H = np.diag([1.0, 1, 0, 0]); R = np.diag([1e-3, 1e-3, 0, 0]); Q = np.diag([1e-2, 1e-2, 1e-4, 1e-4]); P = np.eye(4)
syn = make_tracks(T, nmax, seed=int(os.environ.get("SEED", "5")), device="cpu", nobs_min=100, dts_choices=(1, 2, 3, 6, 12, 24), outlier_frac=0.01, smooth_width=2)
u = HostUKF(H, Q, R, P, gating=gating, packed_cov=True, long_steps=True)
b = TrackBatch.from_synthetic(syn, substeps=k, need_rows=u.model.rows_needed())
t0 = time.time(); res = u.run(b); print("emul s", time.time() - t0, "track-steps", b.track_steps())
print("status counts", res.check_status(raise_on=0))

def one(t):
    m = int(syn.nobs[t])
    z = np.stack([syn.lon[:m, t].numpy(), syn.lat[:m, t].numpy(), syn.sog[:m, t].numpy(), syn.cog[:m, t].numpy()])
    dts = syn.dts[: m - 1, t].numpy()
    kw = dict(gating=gating, mask=np.tile(np.arange(1, k + 1) == k, m - 1))
    a = (z[:, 0], P, H, Q, R, O.generate_dts(dts, k), dts, z, syn.sog_rate[:m, t].numpy(), syn.cog_rate[:m, t].numpy())
    return OC.run_track(*a, **kw), OC.run_track(*a, variant=True, **kw), OC.run_track_precision(*a, precision="extended", **kw)

t0 = time.time()
with ThreadPoolExecutor(8) as ex:
    refs = list(ex.map(one, range(T)))
print("oracle s", time.time() - t0)
worst, ratios = 0, []
for t, (ref, var, ext) in enumerate(refs):
    got = res.track(t)
    e = np.array(track_errors(got, ext)); eo = np.array(track_errors(ref, ext)); ev = np.array(track_errors(var, ext))
    same = np.array_equal(got["gate_iters"], ref["gate_iters"]) if gating else True
    same_ext = np.array_equal(ext["gate_iters"], ref["gate_iters"]) if gating else True
    ratio = float(np.max(e / np.maximum(1e-9, eo)))
    ratios.append(ratio)
    worst = max(worst, ratio)
    if ratio > 1.0 or not same or not same_ext or t < 4:
        print(f"track {t:4d} nobs {int(syn.nobs[t]):5d} device-vs-ext {e}  oracle-vs-ext {eo}  fma-vs-ext {ev} ratio {ratio:.2f} gates_equal {same} {same_ext}")
print("worst device error / max(1e-9, fp64 oracle error), both against extended precision:", worst, " median", float(np.median(ratios)))
