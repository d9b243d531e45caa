#!/bin/bash
# Round 2, GPU call 17: A/B of three instruction-count cuts in the forward kernel (merge copies of the offset trig, selects of the
# rotation parameters, max/min of the "almost diagonal" test)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
QP="python tools/quick_perf.py --tracks 113664 --steps 512 --packed --no-metrics --no-probe"
for v in default m3 m_off m_sel m_tr default m3; do
  if [ $v = default ]; then timeout 300 $QP --label $v >> $O/r2c17_qp.log 2>&1
  else STE_UKF_LIB=$PWD/gpurun_in/variants/libste_$v.so timeout 300 $QP --label $v >> $O/r2c17_qp.log 2>&1; fi
done
grep -h fwd_ms $O/r2c17_qp.log | cut -c1-130
