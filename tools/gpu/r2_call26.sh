#!/bin/bash
# Round 2, GPU call 26: without the peeled sweep (sign / skip changes only)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
QP="python tools/quick_perf.py --tracks 113664 --steps 512 --packed --no-metrics --no-probe"
timeout 300 $QP --label nopeel > $O/r2c26_qp.log 2>&1
timeout 300 $QP --label nopeel_again >> $O/r2c26_qp.log 2>&1
grep -h fwd_ms $O/r2c26_qp.log | cut -c1-130
timeout 900 python bench.py --config c4 --steps 5 --warmup 3 --no-cpu-baseline --e2e-headline-only > $O/r2c26_bench_c4.json 2> $O/r2c26_bench_c4.err; echo "c4 rc $?"
python - <<PY
import json
l=[x for x in open("$O/r2c26_bench_c4.json") if x.startswith("{")][-1]
d=json.loads(l); print("c4 value %.4e"%d["value"], d["roofline"]["forward_ms"], d["roofline"]["backward_ms"])
PY
