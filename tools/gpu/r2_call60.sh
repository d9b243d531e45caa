#!/bin/bash
# Round 2, GPU call 60: the round's last tree - parity suite, smoke, default bench, reference arm
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > $O/r2c60_pytest.log 2>&1; echo "pytest rc $?"; tail -2 $O/r2c60_pytest.log | cut -c1-200
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
( time timeout 900 python bench.py > $O/r2c60_bench_c5.json 2> $O/r2c60_bench_c5.err ) 2> $O/r2c60.time; echo "bench rc $?"; grep real $O/r2c60.time
timeout 600 python bench.py --impl reference > $O/r2c60_ref.json 2> $O/r2c60_ref.err; echo "ref rc $?"
python - <<PY
import json
l=[x for x in open("$O/r2c60_bench_c5.json") if x.startswith("{")][-1]
d=json.loads(l); r=d["roofline"]; p=d.get("partitioned_schedule") or {}
print("c5 value %.4e"%d["value"], "ms/step %.3f"%d["ms_per_step"], "frac", round(r["whole_step"]["frac"],4), r["forward_ms"], r["backward_ms"], "job", d["job"]["wall_s"], "e2e %.3e"%d["e2e"]["value"], d["clocks"]["reasons"])
print("partitioned", {k:p[k] for k in p if k!="schedule"})
r=json.loads(open("$O/r2c60_ref.json").read().strip().split("\n")[-1]); print("ref", r["value"], r["ms_per_step"])
PY
