#!/bin/bash
# Round 2, GPU call 41: instruction / DRAM counts of the kernels on a config-4-shaped tile
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
M=$(python -c "import sys; sys.path.insert(0,'tools'); import ncu_counts; print(ncu_counts.METRICS)")
timeout 300 python tools/c4_counts.py run > $O/r2c41_plain.log 2>&1; tail -1 $O/r2c41_plain.log
timeout 900 ncu --metrics $M --clock-control none -k regex:'ukf_forward|urtss_backward' --csv --log-file $O/r2c41_counts.csv python tools/c4_counts.py run > $O/r2c41_ncu.log 2>&1; echo "ncu rc $?"; tail -1 $O/r2c41_ncu.log
