"""Static SASS statistics of the built library: instructions, code size and opcode counts per kernel."""
import collections, re, subprocess, sys, os
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "ship_track_estimators_b200/csrc/libste_ukf.so")
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
for f in re.split(r"\n\s*Function : ", txt)[1:]:
    name = f.split("\n")[0]
    ops = re.findall(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", f)
    c = collections.Counter(ops)
    fp64 = c['DFMA'] + c['DMUL'] + c['DADD'] + c['DSETP']
    short = re.sub(r"_ZN3ste\d+", "", name)[:44]
    print(f"{short:44s} instr {len(ops):6d} ({len(ops)*16/1024:4.0f} KB) fp64 {fp64:5d} DFMA {c['DFMA']:5d} DMUL {c['DMUL']:4d} DADD {c['DADD']:4d} MUFU {c['MUFU']:3d} UMOV {c['UMOV']:4d} LDS {c['LDS']:3d} STS {c['STS']:3d} LDL {c['LDL']:3d} STL {c['STL']:3d} CALL {c['CALL']}")
