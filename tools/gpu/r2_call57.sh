#!/bin/bash
# Round 2, GPU call 57: the two passes on disjoint SM sets (green contexts): does the DRAM-bound smoother need all 148 SMs?
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python tools/green_ctx_probe.py --splits 24,32,40,48,56,64,80,104 > gpurun_out/r2c57_green.log 2>&1; echo "rc $?"; grep -v Warning gpurun_out/r2c57_green.log | tail -14 | cut -c1-260
