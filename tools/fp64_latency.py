"""FP64-pipe microbenchmark: latency / issue interval for DFMA and for the instruction mixes the
filter kernels actually issue (dev tool; numbers quoted in DESIGN.md)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ship_track_estimators_b200 import _native as nat
lib = nat.load(); dev = torch.device("cuda:0")
sink = torch.zeros(1024, dtype=torch.float64, device=dev); cyc = torch.zeros(1, dtype=torch.int64, device=dev)
iters = 20000
names = {0: "DFMA reg operands", 1: "DFMA constant-bank operands", 2: "DMUL/DADD alternating", 3: "DFMA + compare/select",
         4: "DFMA, three changing register sources", 5: "DMUL, two changing register sources"}
for mode in (0, 1, 2, 3, 4, 5):
    print(names[mode])
    for warps in (1, 4, 12, 16, 32):
        row = []
        for chains in ((1, 2, 4, 8) if mode == 0 else ((4, 8) if mode >= 4 else (1, 4, 8))):
            nat.check(lib.ste_probe_fp64_latency(warps, iters, 100 * mode + chains, nat.ptr(sink), nat.ptr(cyc), nat.current_stream()))
            torch.cuda.synchronize()
            row.append((chains, cyc.item() / iters))
        print(f"  warps/block {warps:2d} (per SMSP {warps/4:4.2f}): SMSP cycles per FP64 op:", [(c, round(r / (c * max(warps / 4, 1)), 2)) for c, r in row])
