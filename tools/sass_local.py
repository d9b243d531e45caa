"""Dev tool: where are the local-memory (LDL/STL) instructions of a kernel, relative to its loops?
usage: python tools/sass_local.py <lib.so> <kernel-name-substring>"""
import re, subprocess, sys
so, pat = sys.argv[1], sys.argv[2]
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
cur, rows = None, []
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        continue
    if cur and pat in cur:
        m = re.search(r"/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m:
            rows.append((int(m.group(1), 16), m.group(2)))
loops = []
for a, ins in rows:
    if "BRA" in ins:
        m = re.search(r"0x([0-9a-f]+)", ins)
        if m and int(m.group(1), 16) < a:
            loops.append((int(m.group(1), 16), a))
loops.sort()
print(f"{len(rows)} instructions; loops:")
for s, e in loops:
    l = sum(1 for a, i in rows if s <= a <= e and re.search(r"\bLDL", i))
    st = sum(1 for a, i in rows if s <= a <= e and re.search(r"\bSTL", i))
    depth = sum(1 for s2, e2 in loops if s2 <= s and e <= e2) - 1
    print(f"  {'  ' * depth}{s:#x}-{e:#x} {(e - s) // 16 + 1:5d} instrs  LDL {l:3d} STL {st:3d}")
if "-v" in sys.argv:
    for a, i in rows:
        if re.search(r"\b(LDL|STL)", i):
            print(f"{a:#x} {i[:70]}")
