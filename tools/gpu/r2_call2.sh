#!/bin/bash
# Round 2, GPU call 2: full parity suite (all tests), benches with the 15-plane tape, e2e probe.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -s > $O/r2c2_pytest.log 2>&1; echo "pytest rc $?" >> $O/r2c2_pytest.log
tail -8 $O/r2c2_pytest.log
QP="python tools/quick_perf.py --tracks 113664 --steps 512 --packed --no-metrics"
timeout 300 $QP --label default > $O/r2c2_qp_default.log 2>&1
grep -h fwd_ms $O/r2c2_qp_*.log | cut -c1-330
timeout 600 python tools/e2e_probe.py > $O/r2c2_e2e_probe.log 2>&1; cat $O/r2c2_e2e_probe.log
timeout 600 python bench.py --steps 5 --warmup 3 > $O/r2c2_bench_c5.json 2> $O/r2c2_bench_c5.err; echo "c5 rc $?"
timeout 900 python bench.py --config c4 --steps 5 --warmup 3 > $O/r2c2_bench_c4.json 2> $O/r2c2_bench_c4.err; echo "c4 rc $?"
ls -la $O | grep r2c2
