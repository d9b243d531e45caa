#!/bin/bash
# Round 2, GPU call 44: full capture of the forward kernel on the config-4-shaped tile
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python tools/c4_counts.py run > $O/r2c44_plain.log 2>&1 && \
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:ukf_forward -c 1 -f -o $O/r2c44_prof_fwd_c4 python tools/c4_counts.py run > $O/r2c44_ncu.log 2>&1
echo "ncu rc $?"; tail -2 $O/r2c44_ncu.log
