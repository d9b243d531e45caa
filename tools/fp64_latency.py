"""DFMA latency / issue-interval microbenchmark on the FP64 pipe (dev tool)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ship_track_estimators_b200 import _native as nat
lib = nat.load(); dev = torch.device("cuda:0")
sink = torch.zeros(1024, dtype=torch.float64, device=dev); cyc = torch.zeros(1, dtype=torch.int64, device=dev)
iters = 20000
for warps in (1, 4, 8, 12, 16, 24, 32):
    row = []
    for chains in (1, 2, 4, 8):
        nat.check(lib.ste_probe_fp64_latency(warps, iters, chains, nat.ptr(sink), nat.ptr(cyc), nat.current_stream()))
        torch.cuda.synchronize()
        row.append(cyc.item() / iters)
    print(f"warps/block {warps:2d} (per SMSP {warps/4:4.2f}): cycles/iter for 1/2/4/8 chains:", [round(r, 2) for r in row], " -> SMSP cycles per warp-DFMA:", [round(r / (ch * max(warps / 4, 1)), 2) for r, ch in zip(row, (1, 2, 4, 8))])
