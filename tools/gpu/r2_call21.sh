#!/bin/bash
# Round 2, GPU call 21: occupancy / block-shape / lock-step variants of the trimmed forward kernel
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
QP="python tools/quick_perf.py --tracks 113664 --steps 512 --packed --no-metrics --no-probe"
rm -f $O/r2c21_qp.log
for v in default pc2 b2 th64 th96 default; do
  if [ $v = default ]; then timeout 300 $QP --label $v >> $O/r2c21_qp.log 2>&1
  else STE_UKF_LIB=$PWD/gpurun_in/variants/libste_$v.so timeout 300 $QP --label $v >> $O/r2c21_qp.log 2>&1; fi
done
grep -h fwd_ms $O/r2c21_qp.log | cut -c1-130
