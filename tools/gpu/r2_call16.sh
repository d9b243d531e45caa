#!/bin/bash
# Round 2, GPU call 16: why the co-resident roles kernel is slower than the two passes back to back
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
QP="python tools/quick_perf.py --packed --no-metrics --no-probe --fused --tracks 113664 --steps 512"
timeout 300 $QP --fused-bwd-tracks 128 --label roles_fwd_only > $O/r2c16_qp_fwd_only.log 2>&1
timeout 300 $QP --fused-fwd-tracks 128 --label roles_bwd_only > $O/r2c16_qp_bwd_only.log 2>&1
timeout 300 $QP --fused-bwd-tracks 28416 --label roles_bwd_quarter > $O/r2c16_qp_bwd_quarter.log 2>&1
STE_UKF_LIB=$PWD/gpurun_in/variants/libste_roles2.so timeout 300 $QP --label roles2 > $O/r2c16_qp_roles2.log 2>&1
grep -h "fused_ms" $O/r2c16_qp_*.log | cut -c1-300
SMALL="python tools/quick_perf.py --tracks 75776 --steps 64 --packed --no-metrics --no-probe --reps 1 --fused"
timeout 300 $SMALL > $O/r2c16_small_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ukf_roles -s 1 -c 1 -f -o $O/r2c16_prof_roles $SMALL > $O/r2c16_ncu_full.log 2>&1
echo "ncu rc $?"
