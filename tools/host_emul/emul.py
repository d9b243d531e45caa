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

if __name__ == "__main__":
    from types import SimpleNamespace
    from _helpers import load_golden, track_errors
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
