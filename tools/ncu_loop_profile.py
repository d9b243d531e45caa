"""Dev tool: stall samples per address bucket of a kernel's SASS (ncu source page, sass view).
usage: ncu -i X.ncu-rep --page source --csv --print-source sass > sass.csv
       python tools/ncu_loop_profile.py sass.csv [bucket_instrs] [lo_hex hi_hex]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
h = rows[hdr]
col = {n: i for i, n in enumerate(h)}
bucket = int(sys.argv[2]) if len(sys.argv) > 2 else 128
lo = int(sys.argv[3], 16) if len(sys.argv) > 4 else 0
hi = int(sys.argv[4], 16) if len(sys.argv) > 4 else 1 << 60
data = []
for r in rows[hdr + 1:]:
    try:
        a = int(r[col["Address"]], 16)
    except ValueError:
        continue
    data.append((a, r))
base = data[0][0]
tot = sum(float(r[col["# Samples"]] or 0) for _, r in data)
stalls = ["stall_long_sb", "stall_math", "stall_wait", "stall_no_inst", "stall_short_sb", "stall_lg", "stall_not_selected", "stall_selected", "stall_branch_resolving", "stall_barrier"]
print(f"total samples {tot:.0f}; bucket = {bucket} instrs")
print("offset    samples  %     execs/instr  " + " ".join(s[6:10] for s in stalls) + "  first-instr")
agg = {}
for a, r in data:
    off = a - base
    if not (lo <= off <= hi):
        continue
    b = off // (16 * bucket)
    e = agg.setdefault(b, [0.0, 0.0, 0, [0.0] * len(stalls), r[col["Source"]][:40]])
    e[0] += float(r[col["# Samples"]] or 0)
    e[1] += float(r[col["Instructions Executed"]] or 0)
    e[2] += 1
    for k, s in enumerate(stalls):
        e[3][k] += float(r[col[s]] or 0)
for b in sorted(agg):
    s, ex, n, st, first = agg[b]
    print(f"{b * bucket * 16:#8x} {s:8.0f} {100 * s / tot:5.1f} {ex / n:10.0f}   " + " ".join(f"{100 * v / max(s, 1):4.0f}" for v in st) + "  " + first)
