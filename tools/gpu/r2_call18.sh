#!/bin/bash
# Round 2, GPU call 18: forward-pass trims (one-Newton reciprocals, cosine series, merged Joseph products, hoisted small-offset test,
# course modulo fast path) - parity suite + timing
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
QP="python tools/quick_perf.py --tracks 113664 --steps 512 --packed --no-metrics --no-probe"
timeout 300 $QP --label trims > $O/r2c18_qp.log 2>&1
STE_UKF_LIB=$PWD/gpurun_in/variants/libste_m3.so timeout 300 $QP --label m3 >> $O/r2c18_qp.log 2>&1
timeout 300 $QP --label trims_again >> $O/r2c18_qp.log 2>&1
grep -h fwd_ms $O/r2c18_qp.log | cut -c1-130
timeout 2400 python -m pytest tests -m gpu -q -x > $O/r2c18_pytest.log 2>&1; echo "pytest rc $?"
grep -v "^  " $O/r2c18_pytest.log | tail -4 | cut -c1-300
