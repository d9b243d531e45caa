#!/bin/bash
# Round 2, GPU call 15: co-resident roles kernel (forward blocks of tile i+1 + backward blocks of tile i in one launch)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k fused > $O/r2c15_pytest.log 2>&1; echo "pytest rc $?"; tail -3 $O/r2c15_pytest.log
QP="python tools/quick_perf.py --packed --no-metrics --no-probe --fused"
timeout 300 $QP --tracks 113664 --steps 512 --label roles > $O/r2c15_qp_roles.log 2>&1
STE_FUSED_PER_THREAD=1 timeout 300 $QP --tracks 113664 --steps 512 --label per_thread > $O/r2c15_qp_per_thread.log 2>&1
timeout 300 $QP --tracks 151552 --steps 1024 --label roles_tile > $O/r2c15_qp_roles_tile.log 2>&1
grep -h "fused_ms\|fwd_ms" $O/r2c15_qp_*.log | cut -c1-400
