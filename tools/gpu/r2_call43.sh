#!/bin/bash
# Round 2, GPU call 43: forward of one tile and backward of another on two streams (tails overlapping)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python tools/quick_perf.py --tracks 113664 --steps 512 --packed --no-metrics --no-probe --two-streams > $O/r2c43_qp.log 2>&1
timeout 300 python tools/quick_perf.py --tracks 151552 --steps 1024 --packed --no-metrics --no-probe --two-streams >> $O/r2c43_qp.log 2>&1
grep -h "two_streams\|fwd_ms" $O/r2c43_qp.log | cut -c1-200
