#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
QP="python tools/quick_perf.py --tracks 113664 --steps 512 --packed --no-metrics --no-probe"
timeout 300 $QP --label default > $O/r2c14_qp_default.log 2>&1
STE_UKF_LIB=$PWD/gpurun_in/variants/libste_lit.so timeout 300 $QP --label lit > $O/r2c14_qp_lit.log 2>&1
timeout 300 $QP --label default_again > $O/r2c14_qp_default2.log 2>&1
STE_UKF_LIB=$PWD/gpurun_in/variants/libste_lit.so timeout 300 $QP --label lit_again > $O/r2c14_qp_lit2.log 2>&1
grep -h fwd_ms $O/r2c14_qp_*.log | cut -c1-150
