"""Dev tool: is the host<->device bandwidth of a pinned buffer a property of the BUFFER (where its pages landed) or of
the moment?  Allocates several pinned buffers, times repeated D2H and H2D copies of each."""
import json, os, sys, time
import torch
dev = torch.device("cuda:0")
n = 1 << 27   # 1 GiB of float64
src = torch.randn(n, dtype=torch.float64, device=dev)
print(open("/proc/cpuinfo").read().count("processor"), "cpus;", "numa nodes:", [d for d in os.listdir("/sys/devices/system/node") if d.startswith("node")] if os.path.exists("/sys/devices/system/node") else "?")
bufs = []
for i in range(6):
    t0 = time.perf_counter(); b = torch.empty(n, dtype=torch.float64).pin_memory(); ta = time.perf_counter() - t0
    bufs.append(b)
    rates = []
    for rep in range(4):
        torch.cuda.synchronize(); t0 = time.perf_counter(); b.copy_(src, non_blocking=True); torch.cuda.synchronize()
        rates.append(round(n * 8 / (time.perf_counter() - t0) / 1e9, 1))
    up = []
    for rep in range(2):
        torch.cuda.synchronize(); t0 = time.perf_counter(); src.copy_(b, non_blocking=True); torch.cuda.synchronize()
        up.append(round(n * 8 / (time.perf_counter() - t0) / 1e9, 1))
    print(json.dumps({"buffer": i, "pin_s": round(ta, 2), "d2h_GBs": rates, "h2d_GBs": up}), flush=True)
# full duplex on two streams, as the pipeline does
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
dst2 = torch.empty(n, dtype=torch.float64, device=dev)
for i in range(0, 6, 2):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    with torch.cuda.stream(s1): bufs[i].copy_(src, non_blocking=True)
    with torch.cuda.stream(s2): dst2.copy_(bufs[i + 1], non_blocking=True)
    torch.cuda.synchronize(); el = time.perf_counter() - t0
    print(json.dumps({"duplex_pair": i, "each_direction_GBs": round(n * 8 / el / 1e9, 1)}), flush=True)
