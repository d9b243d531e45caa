#!/bin/bash
# Round 2, GPU call 33: the examples
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "examples or cli_end or fleet" > gpurun_out/r2c33_pytest.log 2>&1; echo "pytest rc $?"
tail -15 gpurun_out/r2c33_pytest.log | cut -c1-400
