#!/bin/bash
# Round 2, GPU call 38: ptxas --register-usage-level 2 and 8 against the default 5
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
QP="python tools/quick_perf.py --tracks 113664 --steps 512 --packed --no-metrics --no-probe"
rm -f $O/r2c38_qp.log
for v in default ru2 ru8 default ru2 ru8; do
  if [ $v = default ]; then timeout 300 $QP --label $v >> $O/r2c38_qp.log 2>&1
  else STE_UKF_LIB=$PWD/gpurun_in/variants/libste_$v.so timeout 300 $QP --label $v >> $O/r2c38_qp.log 2>&1; fi
done
grep -h fwd_ms $O/r2c38_qp.log | cut -c1-130
