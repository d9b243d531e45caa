#!/bin/bash
# Round 2, GPU call 35: track_metrics as one thread per (track, row)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py -m gpu -q -x -k "metrics or pipelined_output or fleet" > $O/r2c35_pytest.log 2>&1; echo "pytest rc $?"; tail -2 $O/r2c35_pytest.log
timeout 300 python tools/quick_perf.py --tracks 113664 --steps 512 --packed --no-probe > $O/r2c35_qp.log 2>&1; grep track_metrics $O/r2c35_qp.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --e2e-headline-only > $O/r2c35_bench.json 2> $O/r2c35_bench.err; echo "bench rc $?"
python - <<PY
import json
l=[x for x in open("$O/r2c35_bench.json") if x.startswith("{")][-1]
d=json.loads(l); print("c5 value %.4e"%d["value"], "job", d["job"]["wall_s"], d["job"]["value"])
PY
