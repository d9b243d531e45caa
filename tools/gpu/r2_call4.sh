#!/bin/bash
# Round 2, GPU call 4: parity suite (N4 tests, backward fix), quick perf, c5 + c4 bench.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -s > $O/r2c4_pytest.log 2>&1; echo "pytest rc $?" >> $O/r2c4_pytest.log
grep -v "^  " $O/r2c4_pytest.log | tail -12 | cut -c1-300
QP="python tools/quick_perf.py --tracks 113664 --steps 512 --packed --no-metrics"
timeout 300 $QP --label default > $O/r2c4_qp_default.log 2>&1
grep -h fwd_ms $O/r2c4_qp_*.log | cut -c1-330
timeout 600 python bench.py --steps 5 --warmup 3 > $O/r2c4_bench_c5.json 2> $O/r2c4_bench_c5.err; echo "c5 rc $?"
timeout 900 python bench.py --config c4 --steps 5 --warmup 3 > $O/r2c4_bench_c4.json 2> $O/r2c4_bench_c4.err; echo "c4 rc $?"
