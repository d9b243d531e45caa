"""Developer sandbox driver: run the device code compiled for the host (see emul.cpp) on the
golden fixtures.  Not product code; never imported by the package, the tests or the bench."""
import ctypes as C, os, subprocess, sys
HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "tests"))
import numpy as np, torch
from ship_track_estimators_b200 import _native as nat
from ship_track_estimators_b200.batch import BatchedUKF, TrackBatch, TrackResults

_EMUL = None


def build():
    """Compile the host build once per process (and only when a source is newer than the library)."""
    global _EMUL
    if _EMUL is not None:
        return _EMUL
    so = os.path.join(HERE, "libste_emul.so")
    csrc = os.path.join(REPO, "ship_track_estimators_b200", "csrc")
    deps = [os.path.join(HERE, "emul.cpp")] + [os.path.join(csrc, f) for f in os.listdir(csrc) if f.endswith(".cuh")]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps) or os.environ.get("STE_EMUL_FLAGS"):
        _compile(so)
    _EMUL = C.CDLL(so)
    return _EMUL


def _compile(so):
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=fast", "-march=native", *os.environ.get("STE_EMUL_FLAGS", "").split(), "-fPIC", "-shared", "-x", "c++",
                           os.path.join(HERE, "emul.cpp"), "-o", so])

class HostUKF(BatchedUKF):
    """BatchedUKF whose launches go to the host build (CPU tensors)."""
    def __init__(self, *a, **k):
        self._emul = build()
        nat_load = nat.load
        try:
            super().__init__(*a, **k)
        finally:
            pass
    def _call(self, fn, b, res):
        keep = nat.ptr
        nat.ptr = lambda t: None if t is None else t.data_ptr()
        try:
            p, i, o = self._problem(b), self._inputs(b), self._outputs(res)
        finally:
            nat.ptr = keep
        assert fn(C.byref(p), C.byref(i), C.byref(o)) == 0
    def forward(self, b, res): self._call(self._emul.emul_forward, b, res)
    def backward(self, b, res): self._call(self._emul.emul_backward, b, res)
    def fused(self, fb, fr, bb, br):
        keep = nat.ptr
        nat.ptr = lambda t: None if t is None else t.data_ptr()
        try:
            a = (self._problem(fb), self._inputs(fb), self._outputs(fr), self._problem(bb), self._inputs(bb), self._outputs(br))
        finally:
            nat.ptr = keep
        assert self._emul.emul_fused(*[C.byref(v) for v in a]) == 0

def check_fused(names):
    """fused(A, B) must reproduce forward(A) + backward(B) bit for bit (host build)."""
    from types import SimpleNamespace
    from _helpers import load_golden
    for name in names:
        tracks, _ = load_golden(name)
        if "means_s" not in tracks[0]:
            continue
        tr0 = tracks[0]
        gating = "gate_iters" in tr0
        def batch(sel):
            sts = [SimpleNamespace(dts=t["dts"], z=t["z"], sog_rate=t["sog_rate"], cog_rate=t["cog_rate"]) for t in sel]
            noise = [dict(pred=t["noise_pred"], upd=t["noise_upd"], bwd=t.get("noise_bwd")) for t in sel] if "noise_pred" in tr0 else None
            return TrackBatch.from_tracks(sts, [t["dt_array"] for t in sel], device="cpu", x0=[t["x0"] for t in sel], noise=noise, smoother=True)
        A, B = batch(tracks[: max(1, len(tracks) // 2 + 1)]), batch(tracks[::-1][: max(1, len(tracks) // 2)])
        for packed in (False, True):
            u = HostUKF(tr0["H"], tr0["Q"], tr0["R"], tr0["P0"], gating=gating, packed_cov=packed)
            ra, rb = u.allocate(A, smoother=True), u.allocate(B, smoother=True)
            u.forward(A, ra); u.forward(B, rb); u.backward(B, rb)
            fa, fb = u.allocate(A, smoother=True), u.allocate(B, smoother=True)
            u.forward(B, fb); u.fused(A, fa, B, fb)
            u.backward(A, ra)
            ga = u.allocate(A, smoother=True)           # and A smoothed by a fused launch with an empty forward tile
            u.forward(A, ga); u.fused(B, u.allocate(B, smoother=True), A, ga)
            for x, y, n, what in ((ra, fa, A.n_tracks, "A fwd"), (rb, fb, B.n_tracks, "B bwd"), (ra, ga, A.n_tracks, "A bwd")):
                for i in range(n):
                    p, q = x.track(i), y.track(i)
                    for k in p:
                        if what == "A fwd" and k in ("means_s", "covs_s", "status"):
                            continue
                        assert np.array_equal(p[k], q[k], equal_nan=True), (name, packed, what, i, k)
            print(f"{name:22s} packed={packed}: fused == separate (A {A.n_tracks} tracks, B {B.n_tracks} tracks)")

if __name__ == "__main__":
    from types import SimpleNamespace
    from _helpers import load_golden, track_errors
    if sys.argv[1:2] == ["--fused"]:
        check_fused(sys.argv[2:] or ["c1_single_ship", "c2_historical_batch", "c4_ragged_ungated", "c4_ragged_gated", "tape_noise"])
        sys.exit(0)
    names = sys.argv[1:] or ["c1_single_ship", "c2_historical_batch", "c2_modern_ship", "c3_const_dt", "c4_ragged_ungated", "c4_ragged_gated", "tape_noise", "dense_h"]
    for name in names:
        tracks, _ = load_golden(name)
        worst = np.zeros(4); worst_ratio = 0.0
        for generic in (False, True):
            for i, tr in enumerate(tracks):
                st = SimpleNamespace(dts=tr["dts"], z=tr["z"], sog_rate=tr["sog_rate"], cog_rate=tr["cog_rate"])
                sm, gating = "means_s" in tr, "gate_iters" in tr
                noise = [dict(pred=tr["noise_pred"], upd=tr["noise_upd"], bwd=tr.get("noise_bwd"))] if "noise_pred" in tr else None
                b = TrackBatch.from_tracks([st], [tr["dt_array"]], device="cpu", x0=[tr["x0"]], noise=noise, smoother=sm)
                u = HostUKF(tr["H"], tr["Q"], tr["R"], tr["P0"], gating=gating, force_generic=generic)
                g = u.run(b, smoother=sm).track(0)
                assert g["n_updates"] == 1 + int(tr["mask"].sum())
                if gating:
                    assert np.array_equal(g["gate_iters"], tr["gate_iters"]), (name, i, g["gate_iters"], tr["gate_iters"])
                e = np.nan_to_num(np.array(track_errors(g, tr, sm)))
                worst = np.maximum(worst, e)
                worst_ratio = max(worst_ratio, float(np.max(e / np.maximum(1e-9, 10 * tr["unc"]))))
        print(f"{name:22s} worst err {worst}  worst err/bound {worst_ratio:.3f}")
