#!/bin/bash
# Round 2, GPU call 5: device CSV ingest (test + rows/s), c3 bench.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q -s -k "csv or ingest or fleet" > $O/r2c5_pytest.log 2>&1; echo "pytest rc $?" >> $O/r2c5_pytest.log
grep -v "^  " $O/r2c5_pytest.log | tail -30 | cut -c1-300
timeout 1500 python tools/ingest_bench.py --rows 20000000 --ships 20000 > $O/r2c5_ingest.log 2>&1; echo "ingest rc $?"; cat $O/r2c5_ingest.log | cut -c1-600
timeout 600 python bench.py --config c3 --steps 5 --warmup 3 > $O/r2c5_bench_c3.json 2> $O/r2c5_bench_c3.err; echo "c3 rc $?"
