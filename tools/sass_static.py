"""Static opcode counts of one kernel by innermost source function (nvdisasm -gi inline chains): a quick look at what a
change did to the generated code before GPU time is spent.  Counts are per instruction in the binary (rolled loops once,
cold code included), so compare like with like.
usage: python tools/sass_static.py <library.so> <mangled-kernel-substring> [OPCODE ...]   (default opcodes: IMAD.MOV.U32 FSEL)"""
import collections, os, re, subprocess, sys, tempfile

lib, sub = sys.argv[1], sys.argv[2]
ops = sys.argv[3:] or ["IMAD.MOV.U32", "FSEL"]
SRC = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "ship_track_estimators_b200", "csrc")
ranges = {}
for fn in os.listdir(SRC):
    if fn.endswith((".cuh", ".cu")):
        starts = []
        for i, l in enumerate(open(os.path.join(SRC, fn)).read().split("\n"), 1):
            m = re.match(r"^(?:template.*>\s*)?(?:STE_DEV|__global__|static|extern|STE_HD|int|__device__|STE_COLD).*?\b([A-Za-z_][A-Za-z0-9_]*)\s*\(", l)
            if m and not l.startswith(" ") and m.group(1) not in ("defined", "__launch_bounds__", "if"):
                starts.append((i, m.group(1)))
        ranges[fn] = starts


def func_of(path, line):
    best = os.path.basename(path)
    for s, name in ranges.get(best, []):
        if s <= line:
            best = name
    return best


with tempfile.TemporaryDirectory() as d:
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=d, check=True, capture_output=True)
    cubin = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
    txt = subprocess.run(["nvdisasm", "-gi", "-c", os.path.join(d, cubin)], capture_output=True, text=True).stdout
in_k, chain, pending = False, [], True
cnt, total, by_fn = collections.defaultdict(collections.Counter), 0, collections.Counter()
for l in txt.splitlines():
    if l.startswith(".text."):
        in_k = sub in l
        continue
    if not in_k:
        continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m:
        if pending:
            chain, pending = [], False
        chain.append((m.group(1), int(m.group(2))))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        pending = True
        op = re.match(r"(?:@!?U?P[T\d]+\s+)?([A-Z0-9_.]+)", m.group(2).strip()).group(1)
        names = [func_of(p, ln) for p, ln in chain] or ["?"]
        total += 1
        by_fn[names[0]] += 1
        if op in ops:
            cnt[op][names[0] + " < " + (names[1] if len(names) > 1 else "")] += 1
print(f"{sub}: {total} instructions ({total * 16 / 1024:.1f} KB)")
print("by function:", ", ".join(f"{k} {v}" for k, v in by_fn.most_common(25)))
for op in ops:
    print(op, sum(cnt[op].values()))
    for k, v in cnt[op].most_common(12):
        print(f"    {k:60s} {v}")
