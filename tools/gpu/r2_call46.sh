#!/bin/bash
# Round 2, GPU call 46: final tree - parity suite, smoke, default bench, reference arm
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > $O/r2c46_pytest.log 2>&1; echo "pytest rc $?"; tail -2 $O/r2c46_pytest.log | cut -c1-200
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py > $O/r2c46_bench.json 2> $O/r2c46_bench.err; echo "bench rc $?"
timeout 600 python bench.py --impl reference > $O/r2c46_ref.json 2> $O/r2c46_ref.err; echo "ref rc $?"
python - <<PY
import json
l=[x for x in open("$O/r2c46_bench.json") if x.startswith("{")][-1]
d=json.loads(l); print("c5 value %.4e"%d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["whole_step"]["frac"], "e2e %.3e"%d["e2e"]["value"], "job", d["job"]["wall_s"], d["clocks"])
r=json.loads(open("$O/r2c46_ref.json").read().strip().split("\n")[-1]); print("ref", r["value"], r["steps"], r["ms_per_step"])
PY
