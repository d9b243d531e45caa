#!/bin/bash
# Round 2, GPU call 24: the one-launch roles kernel again, with the trimmed forward code
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
QP="python tools/quick_perf.py --packed --no-metrics --no-probe --fused --tracks 113664 --steps 512"
timeout 300 $QP --label roles > $O/r2c24_qp.log 2>&1
timeout 300 $QP --fused-bwd-tracks 128 --label roles_fwd_only >> $O/r2c24_qp.log 2>&1
timeout 300 $QP --fused-fwd-tracks 128 --label roles_bwd_only >> $O/r2c24_qp.log 2>&1
grep -h "fused_ms" $O/r2c24_qp.log | cut -c1-200
