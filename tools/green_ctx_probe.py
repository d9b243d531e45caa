"""Developer probe: the two passes on disjoint sets of SMs (CUDA green contexts, driver API through cuda-python).
The forward filter is bound by the FP64 pipe and the tape smoother by DRAM; on the SAME SMs their instruction streams do
not fit the instruction cache together (ukf_roles_kernel).  This measures (i) each pass alone on n SMs and
(ii) forward of one tile on 148 - n SMs beside backward of another on n SMs, against the two launches back to back.
usage: python tools/green_ctx_probe.py [--tracks 113664] [--steps 512] [--splits 32,40,48,56,64]"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from cuda.bindings import driver as cu
from ship_track_estimators_b200.batch import BatchedUKF, TrackBatch
from ship_track_estimators_b200.synthetic import make_tracks

ap = argparse.ArgumentParser()
ap.add_argument("--tracks", type=int, default=148 * 128 * 6)
ap.add_argument("--steps", type=int, default=512)
ap.add_argument("--splits", default="32,40,48,56,64,80")
ap.add_argument("--reps", type=int, default=3)
a = ap.parse_args()


def ck(r):
    err, rest = r[0], r[1:]
    if err != cu.CUresult.CUDA_SUCCESS:
        raise RuntimeError(str(err))
    return rest[0] if len(rest) == 1 else rest


torch.cuda.init(); torch.zeros(1, device="cuda:0")
dev = ck(cu.cuDeviceGet(0))
sm_all = ck(cu.cuDeviceGetDevResource(dev, cu.CUdevResourceType.CU_DEV_RESOURCE_TYPE_SM))
print(json.dumps({"sm_total": sm_all.sm.smCount}))


def split_streams(n):
    """-> (stream on a green context of >= n SMs, stream on the remaining SMs, the two SM counts, the contexts)."""
    groups, nb, rem = ck(cu.cuDevSmResourceSplitByCount(1, sm_all, 0, n))
    out = []
    for res in (groups[0], rem):
        desc = ck(cu.cuDevResourceGenerateDesc([res], 1))
        g = ck(cu.cuGreenCtxCreate(desc, dev, cu.CUgreenCtxCreate_flags.CU_GREEN_CTX_DEFAULT_STREAM))
        s = ck(cu.cuGreenCtxStreamCreate(g, cu.CUstream_flags.CU_STREAM_NON_BLOCKING, 0))
        out.append((torch.cuda.ExternalStream(int(s)), res.sm.smCount, g))
    return out


H = np.diag([1.0, 1, 0, 0]); R = np.diag([1e-3, 1e-3, 0, 0]); Q = np.diag([1e-2, 1e-2, 1e-4, 1e-4]); P = np.eye(4)
ukf = BatchedUKF(H, Q, R, P, packed_cov=True, long_steps=False)
tiles = []
for seed in (1, 2):
    b = TrackBatch.from_synthetic(make_tracks(a.tracks, a.steps + 1, seed=seed, device="cuda:0"), substeps=1)
    r = ukf.allocate(b, smoother=True)
    ukf.forward(b, r)
    tiles.append((b, r))
torch.cuda.synchronize()
(b0, r0), (b1, r1) = tiles


def timed(fn):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(a.reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts)


def on(stream, fn):
    cur = torch.cuda.current_stream()
    stream.wait_stream(cur)
    with torch.cuda.stream(stream):
        fn()
    cur.wait_stream(stream)


f_ms = timed(lambda: ukf.forward(b1, r1)); b_ms = timed(lambda: ukf.backward(b0, r0))
print(json.dumps({"all_sms": True, "fwd_ms": f_ms, "bwd_ms": b_ms, "back_to_back_ms": f_ms + b_ms}))
for n in [int(x) for x in a.splits.split(",")]:
    (sb, nb_sm, gb), (sf, nf_sm, gf) = split_streams(n)
    fb = timed(lambda: on(sb, lambda: ukf.backward(b0, r0)))
    ff = timed(lambda: on(sf, lambda: ukf.forward(b1, r1)))

    def both():
        cur = torch.cuda.current_stream()
        sb.wait_stream(cur); sf.wait_stream(cur)
        with torch.cuda.stream(sf):
            ukf.forward(b1, r1)
        with torch.cuda.stream(sb):
            ukf.backward(b0, r0)
        cur.wait_stream(sb); cur.wait_stream(sf)
    bo = timed(both)
    print(json.dumps({"bwd_sms": nb_sm, "fwd_sms": nf_sm, "bwd_alone_ms": fb, "fwd_alone_ms": ff, "both_ms": bo,
                      "back_to_back_ms": f_ms + b_ms, "gain": (f_ms + b_ms) / bo}))
    torch.cuda.synchronize()
