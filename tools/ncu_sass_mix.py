"""Summarise an `ncu --page source --csv --print-source sass` dump: per-kernel opcode mix
(warp-level executed instructions) and stall-reason totals.  Dev tool."""
import csv, sys, collections, re
path = sys.argv[1]
per_step = float(sys.argv[2]) if len(sys.argv) > 2 else None   # warps*steps to normalise by
rows = list(csv.reader(open(path)))
kernels = []
cur = None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "rows": []}
        kernels.append(cur)
    elif r and r[0] == "Address":
        cur["hdr"] = r
    elif cur is not None and cur["hdr"] is not None and r:
        cur["rows"].append(r)
for k in kernels:
    h = {n: i for i, n in enumerate(k["hdr"])}
    mix = collections.Counter(); stall = collections.Counter(); total = 0
    for r in k["rows"]:
        try: n = float(r[h["Instructions Executed"]])
        except ValueError: continue
        op = r[h["Source"]].strip()
        m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_]+)", op)
        opc = m.group(2) if m else op.split()[0]
        mix[opc] += n; total += n
        for name in h:
            if name.startswith("stall_") and "Not Issued" not in name:
                try: stall[name] += float(r[h[name]])
                except ValueError: pass
    print("==", k["name"][:70], "total warp-inst", int(total), ("per warp-step %.0f" % (total / per_step)) if per_step else "")
    for opc, n in mix.most_common(28):
        print(f"   {opc:12s} {n/total*100:6.2f}%", ("%8.1f /step" % (n / per_step)) if per_step else "")
    st = sum(stall.values())
    print("   stalls:", ", ".join(f"{a[6:]} {b/st*100:.1f}%" for a, b in stall.most_common(8)))
