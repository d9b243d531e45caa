#!/bin/bash
# Round 2, GPU call 58: green-context split at the bench tile's width (151 552 tracks = 1 184 blocks), finer splits
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python tools/green_ctx_probe.py --tracks 151552 --steps 512 --splits 40,44,48,52,56,60 > gpurun_out/r2c58_green.log 2>&1; echo "rc $?"; grep -v Warning gpurun_out/r2c58_green.log | tail -9 | cut -c1-260
