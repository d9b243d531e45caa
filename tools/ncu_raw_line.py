"""Two summary lines from an `ncu --page raw --csv` dump on stdin (first kernel): duration, registers, FP64-pipe and issue
activity, instruction-cache hit rate, and the stall cycles per issued instruction in descending order.  Dev tool."""
import csv, sys
rows = list(csv.reader(sys.stdin))
hdr, val = rows[0], rows[2]
g = {h: v for h, v in zip(hdr, val)}
f = lambda k: float(g[k].replace(",", ""))
dur = f("gpu__time_duration.sum") * {"ms": 1e3, "us": 1.0, "s": 1e6, "ns": 1e-3}[rows[1][hdr.index("gpu__time_duration.sum")]]
print("raw page: duration %.1f us (ncu), %d registers, FP64 pipe %.1f %% active, issue slots %.1f %% active, instruction-cache hit rate %.1f %%;" % (
    dur, f("launch__registers_per_thread"), f("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
    f("smsp__issue_active.avg.pct_of_peak_sustained_active"), f("sm__icc_request_hit_rate.pct")))
st = []
for h in hdr:
    if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") and "selected_per" not in h.replace("not_selected", "x"):
        st.append((f(h), h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
st.sort(reverse=True)
print("stall cycles per issued instruction: " + ", ".join("%s %.2f" % (n, v) for v, n in st[:7]) + ".")
