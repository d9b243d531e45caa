#!/bin/bash
# Round 2, GPU call 45: medium-displacement tier for launches with long steps (config 4, real ship data)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
QP="python tools/quick_perf.py --tracks 113664 --steps 512 --packed --no-metrics --no-probe"
timeout 300 $QP --label medium_tree > $O/r2c45_qp.log 2>&1; grep -h fwd_ms $O/r2c45_qp.log | cut -c1-130
timeout 900 python bench.py --config c4 --steps 5 --warmup 3 --no-cpu-baseline --e2e-headline-only > $O/r2c45_bench_c4.json 2> $O/r2c45_bench_c4.err; echo "c4 rc $?"
timeout 600 python bench.py --config c2 --steps 5 --warmup 3 --no-cpu-baseline > $O/r2c45_bench_c2.json 2> $O/r2c45_bench_c2.err; echo "c2 rc $?"
python - <<PY
import json
for f in ("c4","c2"):
    l=[x for x in open("$O/r2c45_bench_%s.json"%f) if x.startswith("{")][-1]
    d=json.loads(l); print(f, "value %.4e"%d["value"], d["roofline"]["forward_ms"], d["roofline"]["backward_ms"], d.get("parity_vs_reference"))
PY
timeout 2400 python -m pytest tests -m gpu -q -x > $O/r2c45_pytest.log 2>&1; echo "pytest rc $?"; tail -3 $O/r2c45_pytest.log | cut -c1-300
