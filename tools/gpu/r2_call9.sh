#!/bin/bash
# Round 2, GPU call 9: forward-kernel variants (lock-step multi-column pair loop, block shapes).
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
QP="python tools/quick_perf.py --tracks 113664 --steps 512 --packed --no-metrics --no-probe"
timeout 300 $QP --label default > $O/r2c9_qp_default.log 2>&1
for v in pc2 pc4 pc2b2 pc4b2 t64 pf2 bwd3 bwd3pf2; do
  STE_UKF_LIB=$PWD/gpurun_in/variants/libste_$v.so timeout 300 $QP --label $v > $O/r2c9_qp_$v.log 2>&1
done
timeout 300 $QP --label default_again > $O/r2c9_qp_default2.log 2>&1
grep -h fwd_ms $O/r2c9_qp_*.log | cut -c1-130
