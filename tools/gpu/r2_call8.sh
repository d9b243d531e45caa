#!/bin/bash
# Round 2, GPU call 8: backward L2 prefetch A/B, then the parity suite and c5 bench on the chosen build.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
QP="python tools/quick_perf.py --tracks 113664 --steps 512 --packed --no-metrics --no-probe"
timeout 300 $QP --label prefetch > $O/r2c8_qp_pf.log 2>&1
STE_UKF_LIB=$PWD/gpurun_in/variants/libste_nopf.so timeout 300 $QP --label noprefetch > $O/r2c8_qp_nopf.log 2>&1
timeout 300 $QP --label prefetch_again > $O/r2c8_qp_pf2.log 2>&1
grep -h fwd_ms $O/r2c8_qp_*.log | cut -c1-260
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $O/r2c8_bench_c5.json 2> $O/r2c8_bench_c5.err; echo "c5 rc $?"
timeout 900 python -m pytest tests -m gpu -q -x -k "golden or polar or fused or bench_shape or synthetic" > $O/r2c8_pytest.log 2>&1; echo "pytest rc $?" >> $O/r2c8_pytest.log
grep -v "^  " $O/r2c8_pytest.log | tail -4 | cut -c1-300
