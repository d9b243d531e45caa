"""Dev tool: time BatchedUKF.run_host_pipelined per output set, with the copy and kernel shares (CUDA events)."""
import os, sys, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ship_track_estimators_b200.batch import BatchedUKF, TrackBatch
from ship_track_estimators_b200.synthetic import make_tracks
dev = torch.device("cuda:0")
H = np.diag([1.0, 1, 0, 0]); R = np.diag([1e-3, 1e-3, 0, 0]); Q = np.diag([1e-2, 1e-2, 1e-4, 1e-4]); P = np.eye(4)
ukf = BatchedUKF(H, Q, R, P, packed_cov=True, long_steps=False)
T, N, n = 148 * 128, 1024, 6
tiles = [TrackBatch.from_synthetic(make_tracks(T, N + 1, seed=5 + j, device="cuda:0"), substeps=1).pin_memory() for j in range(2)]
seq = [tiles[i % 2] for i in range(n)]
# raw copy rates of this box
big = torch.empty(1 << 28, dtype=torch.float64, device=dev); hbig = torch.empty(1 << 28, dtype=torch.float64).pin_memory()
for name, fn in (("d2h", lambda: hbig.copy_(big, non_blocking=True)), ("h2d", lambda: big.copy_(hbig, non_blocking=True))):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); torch.cuda.synchronize()
    print(json.dumps({"copy": name, "GBs": big.numel() * 8 / (time.perf_counter() - t0) / 1e9}))
del big, hbig
for name in ("smoothed", "cli", "all", "summary"):
    t0 = time.perf_counter()
    outs = [ukf.host_outputs(tiles[0], outputs=name) for _ in range(2)]
    alloc_s = time.perf_counter() - t0
    so = [outs[i % 2] for i in range(n)]
    ukf.run_host_pipelined(seq[:2], so[:2], device=dev, outputs=name); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); moved = ukf.run_host_pipelined(seq, so, device=dev, outputs=name); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(json.dumps({"set": name, "ms_per_tile": ms / n, "track_steps_per_s": T * N * n / ms * 1e3, "d2h_GBs": moved["d2h_bytes"] / ms / 1e6,
                      "h2d_GBs": moved["h2d_bytes"] / ms / 1e6, "pin_alloc_s": alloc_s, "d2h_bytes_per_state": moved["d2h_bytes"] / n / (T * (N + 1))}))
    del outs, so
