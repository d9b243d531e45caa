#!/bin/bash
# Round 2, GPU call 34: config 4 with every tile of the 1 M-track fleet (--full-job)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
( time timeout 1500 python bench.py --config c4 --full-job --no-cpu-baseline --e2e-headline-only > $O/r2c34_bench_c4_full.json 2> $O/r2c34_bench_c4_full.err ) 2> $O/r2c34_time.log; echo "rc $?"; cat $O/r2c34_time.log | tr '\n' ' '; echo
python - <<PY
import json
l=[x for x in open("$O/r2c34_bench_c4_full.json") if x.startswith("{")][-1]
d=json.loads(l); print("c4 full value %.4e"%d["value"], "tiles", d["steps"], "fwd/bwd ms", d["roofline"]["forward_ms"], d["roofline"]["backward_ms"], d["summary"])
PY
