"""Attribute executed SASS instructions (ncu source page, sass view) to source functions using the
inline chains nvdisasm -gi prints.  Usage:
  ncu -i X.ncu-rep --page source --csv --print-source sass > sass.csv
  cuobjdump -xelf all libste_ukf.so && nvdisasm -gi -c ste_ukf.sm_100a.cubin > disasm_gi.txt
  python tools/ncu_attrib.py sass.csv disasm_gi.txt <kernel-substring> <warps*steps>
Dev tool (round-1 profiling)."""
import collections, csv, os, re, sys

sass_csv, disasm, kernel_sub, per = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4])
# "mangled-substring|demangled-substring" selects the nvdisasm section and the ncu kernel separately
dis_sub, csv_sub = (kernel_sub.split("|") + [kernel_sub])[:2]
SRC = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "ship_track_estimators_b200", "csrc")

# function line ranges per source file
ranges = {}
for fn in os.listdir(SRC):
    if not fn.endswith((".cuh", ".cu")):
        continue
    lines = open(os.path.join(SRC, fn)).read().split("\n")
    starts = []
    for i, l in enumerate(lines, 1):
        m = re.match(r"^(?:template.*>\s*)?(?:STE_DEV|__global__|static|extern|STE_HD|int|__device__).*?\b([A-Za-z_][A-Za-z0-9_]*)\s*\(", l)
        if m and not l.startswith(" ") and m.group(1) not in ("defined", "__launch_bounds__", "if"):
            starts.append((i, m.group(1)))
        m2 = re.match(r"^\s+auto (\w+) = \[&\]", l)
        if m2:
            starts.append((i, m2.group(1)))
    ranges[fn] = starts

def func_of(path, line):
    fn = os.path.basename(path)
    best = fn
    for s, name in ranges.get(fn, []):
        if s <= line:
            best = name
    return best

# per-address inline chains for the kernel
chains, cur_chain, in_kernel, pending_new = {}, [], False, True
for l in open(disasm):
    if l.startswith(".text."):
        in_kernel = dis_sub in l
        continue
    if not in_kernel:
        continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', l)
    if m:
        if not cur_chain or cur_chain[-1][2]:
            pass
        cur_chain_entry = (m.group(1), int(m.group(2)), m.group(3) is not None)
        if pending_new:
            cur_chain = []
            pending_new = False
        cur_chain.append(cur_chain_entry)
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        chains[int(m.group(1), 16)] = list(cur_chain)
        pending_new = True

rows = list(csv.reader(open(sass_csv)))
hdr, take, counts = None, False, {}
for r in rows:
    if r and r[0] == "Kernel Name":
        take = csv_sub in r[1] and not counts
        hdr = None
    elif r and r[0] == "Address":
        hdr = {n: i for i, n in enumerate(r)}
    elif take and hdr and r:
        try:
            counts[int(r[hdr["Address"]], 16) if r[hdr["Address"]].startswith("0x") else int(r[hdr["Address"]])] = (float(r[hdr["Instructions Executed"]]), r[hdr["Source"]], float(r[hdr["# Samples"]] or 0), {k: float(r[i] or 0) for k, i in hdr.items() if k.startswith("stall_") and "Not Issued" not in k})
        except ValueError:
            pass
base = min(counts) if counts else 0
inner, path_tot, fp64_inner = collections.Counter(), collections.Counter(), collections.Counter()
WATCH = set(os.environ.get("STE_ATTRIB_OPS", "").split())   # e.g. STE_ATTRIB_OPS="FSEL IMAD": where these opcodes are executed
watch_inner = collections.Counter()
tot = 0
samples, stall_by = collections.Counter(), collections.defaultdict(collections.Counter)
for addr, (n, src, smp, stl) in counts.items():
    ch = chains.get(addr - base, [])
    names = [func_of(p, ln) for p, ln, _ in ch] or ["?"]
    # collapse consecutive duplicates, innermost first
    dedup = [names[0]]
    for x in names[1:]:
        if x != dedup[-1]:
            dedup.append(x)
    inner[dedup[0]] += n
    path_tot[" < ".join(dedup[:4])] += n
    key3 = " < ".join(dedup[:3])
    samples[key3] += smp
    for k, v in stl.items():
        stall_by[key3][k] += v
    op = re.match(r"(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", src.strip())
    if op and op.group(1) in ("DFMA", "DMUL", "DADD", "DSETP"):
        fp64_inner[" < ".join(dedup[:4])] += n
    if op and op.group(1) in WATCH:
        watch_inner[(op.group(1), " < ".join(dedup[:3]))] += n
    tot += n
print(f"kernel ~{kernel_sub}: {tot/per:.0f} warp-instructions per warp-step")
print("-- by innermost function")
for k, v in inner.most_common(20):
    print(f"   {k:28s} {v/per:8.1f}  {100*v/tot:5.1f}%")
print("-- by inline path (innermost < callers), all / FP64-pipe")
for k, v in path_tot.most_common(40):
    print(f"   {k:90s} {v/per:8.1f} {fp64_inner[k]/per:8.1f}")

tot_s = sum(samples.values())
print("-- stall samples by inline path (share of all samples; top stall reasons)")
for k, v in samples.most_common(16):
    top = ", ".join(f"{a[6:]} {100*b/max(v,1):.0f}%" for a, b in stall_by[k].most_common(4))
    print(f"   {k:72s} {100*v/tot_s:5.1f}%   {top}")

if WATCH:
    print("-- watched opcodes by inline path (per warp-step)")
    for (o, k), v in watch_inner.most_common(30):
        print(f"   {o:8s} {k:80s} {v/per:8.1f}")
