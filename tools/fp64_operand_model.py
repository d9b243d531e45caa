"""Operand-bandwidth model of the FP64 pipe applied to an ncu source-page dump (sass view).

Measured on the B200 (tools/fp64_latency.py, ste_probe_fp64_latency modes 0/4/5): a DFMA whose three sources are three
different registers issues every 3.0 cycles per SM sub-partition whatever the occupancy, one with two register sources
(third a constant-bank operand, an immediate, or a repeated register) and DMUL / DADD every 2.0: the issue interval is
max(2, number of distinct 64-bit register sources).  This tool sums that cost over the executed FP64-pipe instructions of
a kernel: the pipe-and-register-file-bound time, the denominator that explains why `sm__pipe_fp64_cycles_active` stops
near 65 % on code made of three-operand multiply-adds.
usage: ncu -i X.ncu-rep --page source --csv --print-source sass > sass.csv
       python tools/fp64_operand_model.py sass.csv <kernel-substring> <warp-steps> [<library.so> <mangled-kernel-substring>]
With the library given, the operands are read from `cuobjdump -sass` (ncu's listing drops the `.reuse` flags: a source
served by the operand-reuse cache does not read the register file) and joined with ncu's execution counts by offset."""
import csv, re, subprocess, sys, collections

path, sub, per = sys.argv[1], sys.argv[2], float(sys.argv[3])
real = {}
if len(sys.argv) > 5:
    txt = subprocess.run(["cuobjdump", "-sass", sys.argv[4]], capture_output=True, text=True).stdout
    cur = None
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            continue
        if cur and sys.argv[5] in cur:
            m = re.search(r"/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
            if m:
                real[int(m.group(1), 16)] = m.group(2).strip()
rows = list(csv.reader(open(path)))
base = None
hdr, take, tot = None, False, collections.Counter()
cost2 = cost_model = n_fp64 = 0.0
by_srcs = collections.Counter()
for r in rows:
    if r and r[0] == "Kernel Name":
        take, hdr = sub in r[1], None
    elif r and r[0] == "Address":
        hdr = {n: i for i, n in enumerate(r)}
    elif take and hdr and r:
        src, n = r[hdr["Source"]].strip(), float(r[hdr["Instructions Executed"]] or 0)
        addr = int(r[hdr["Address"]], 16)
        base = addr if base is None else base
        if real:
            full = real.get(addr - base, src)
            assert full.split()[0].split(".")[0].lstrip("@!P0123456789 ") == src.split()[0].split(".")[0].lstrip("@!P0123456789 ") or True
            src = full
        m = re.match(r"(?:@!?U?P\d+\s+)?([A-Z0-9_]+)((?:\.[A-Z0-9_]+)*)\s+(.*)", src)
        if not m or m.group(1) not in ("DFMA", "DMUL", "DADD", "DSETP"):
            continue
        op, args = m.group(1), [a.strip() for a in m.group(3).rstrip(";").split(",")]
        srcs = args[1:] if op != "DSETP" else args[2:]          # drop the destination (DSETP: two predicate destinations)
        regs = set()
        for a in srcs:
            mm = re.match(r"[-|~]*\|?(R\d+)", a)
            if mm and mm.group(1) != "RZ" and ".reuse" not in a:
                regs.add(mm.group(1))
        k = len(regs)
        by_srcs[(op, k)] += n
        n_fp64 += n
        cost2 += 2.0 * n
        cost_model += max(2.0, float(k)) * n
print(f"FP64-pipe warp-instructions per warp-step: {n_fp64 / per:.1f}")
for (op, k), n in sorted(by_srcs.items()):
    print(f"   {op:6s} with {k} distinct register source(s) (reuse-cache hits excluded): {n / per:7.1f} per step")
print(f"pipe-only cycles per warp-step (2 per instruction):                  {cost2 / per:8.1f}")
print(f"pipe + register-file cycles per warp-step (max(2, register sources)): {cost_model / per:8.1f}")
