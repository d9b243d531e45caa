#!/bin/bash
# Round 2, GPU call 22: parity suite with the dimension-generic smoother and the new tier-edge cases
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -s > $O/r2c22_pytest.log 2>&1; echo "pytest rc $?"
grep -v "^  " $O/r2c22_pytest.log | tail -6 | cut -c1-300
grep "n = 5 track" $O/r2c22_pytest.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
