"""Instruction / FP64 / DRAM counts of the kernels on a config-4-shaped tile (ragged, k = 2, box smoothing 2, 1 % displaced
fixes, gating, URTSS, full-range geodetic tier chosen automatically), for profiles/kernel_counts_c4.json.

    python tools/c4_counts.py run                      # the workload: one forward + one backward launch; prints its track-steps
    ncu --metrics <ncu_counts.METRICS> --clock-control none -k regex:'ukf_forward|urtss_backward' --csv --log-file c.csv python tools/c4_counts.py run
    python tools/c4_counts.py parse c.csv <track-steps> > profiles/kernel_counts_c4.json
"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

T, NMAX, NMIN = 148 * 128 * 2, 400, 100


def run():
    import numpy as np, torch
    from ship_track_estimators_b200.batch import BatchedUKF, TrackBatch
    from ship_track_estimators_b200.synthetic import make_tracks
    H = np.diag([1.0, 1, 0, 0]); Q = np.diag([1e-2, 1e-2, 1e-4, 1e-4]); R = np.diag([1e-3, 1e-3, 0, 0]); P = np.eye(4)
    lengths = torch.from_numpy(np.sort(np.random.default_rng(3).integers(NMIN, NMAX + 1, size=T))[::-1].astype(np.int32).copy())
    syn = make_tracks(T, NMAX, seed=2, device="cuda:0", dts_choices=(1.0, 2.0, 3.0, 6.0, 12.0, 24.0), outlier_frac=0.01, smooth_width=2, lengths=lengths)
    ukf = BatchedUKF(H, Q, R, P, gating=True, packed_cov=True)
    b = TrackBatch.from_synthetic(syn, substeps=2, need_rows=ukf.model.rows_needed())
    res = ukf.allocate(b, smoother=True, in_place=True)
    ukf._long_steps_for(b)
    torch.cuda.synchronize()
    ukf.forward(b, res); ukf.backward(b, res)
    torch.cuda.synchronize()
    print(json.dumps({"track_steps": b.track_steps(), "tracks": T, "gated_updates": int((res.gate_iters > 0).sum()), "long_steps": bool(ukf._long_steps_for(b))}))


def parse(path, track_steps):
    import ncu_counts as N
    launches = N.parse(path)
    fwd = [l for l in launches if "ukf_forward" in l["kernel"]][-1]
    bwd = [l for l in launches if "urtss_backward" in l["kernel"]][-1]
    out = {"forward": N.per_step(fwd, track_steps), "backward": N.per_step(bwd, track_steps),
           "source": f"ncu --metrics ... --clock-control none over tools/c4_counts.py run ({T} ragged tracks of {NMIN}-{NMAX} fixes, k = 2, gating, "
                     f"full-range geodetic tier; {track_steps} track-steps): per track-step = per launch / track-steps"}
    json.dump(out, sys.stdout, indent=1)
    print()


if __name__ == "__main__":
    if sys.argv[1] == "run":
        run()
    else:
        parse(sys.argv[2], float(sys.argv[3]))
