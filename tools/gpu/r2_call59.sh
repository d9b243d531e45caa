#!/bin/bash
# Round 2, GPU call 59: SM-partitioned schedule - bit-identity test, default bench with the partitioned_schedule key
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "partitioned or fused" > $O/r2c59_pytest.log 2>&1; echo "pytest rc $?"; tail -5 $O/r2c59_pytest.log | cut -c1-300
( time timeout 900 python bench.py --no-cpu-baseline > $O/r2c59_bench_c5.json 2> $O/r2c59_bench_c5.err ) 2> $O/r2c59.time; echo "bench rc $?"; grep real $O/r2c59.time; tail -3 $O/r2c59_bench_c5.err | cut -c1-300
python - <<PY
import json
l=[x for x in open("$O/r2c59_bench_c5.json") if x.startswith("{")][-1]
d=json.loads(l); r=d["roofline"]
print("c5 value %.4e"%d["value"], "ms/step %.3f"%d["ms_per_step"], "frac", round(r["whole_step"]["frac"],4), "job", d["job"]["wall_s"], "e2e %.3e"%d["e2e"]["value"])
print(json.dumps(d.get("partitioned_schedule"))[:700])
PY
for n in 40 56; do timeout 600 python bench.py --no-cpu-baseline --no-job --e2e-headline-only --smoother-sms $n 2>/dev/null | python -c "
import sys,json
d=json.loads([x for x in sys.stdin if x.startswith('{')][-1]); p=d.get('partitioned_schedule'); print($n, {k:p[k] for k in p if k!='schedule'})"; done
