#!/bin/bash
# Round 2, GPU call 27: final record - parity suite, counts + full captures of both kernels, bench c5 / c3 / c4 / reference arm, launch list
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -s > $O/r2c27_pytest.log 2>&1; echo "pytest rc $?"
grep -v "^  " $O/r2c27_pytest.log | tail -3 | cut -c1-300
SMALL="python tools/quick_perf.py --tracks 75776 --steps 64 --packed --no-metrics --no-probe --reps 1"
SMALLF="python tools/quick_perf.py --tracks 75776 --steps 64 --no-metrics --no-probe --reps 1"
M=$(python -c "import sys; sys.path.insert(0,'tools'); import ncu_counts; print(ncu_counts.METRICS)")
timeout 300 $SMALL > $O/r2c27_small_plain.log 2>&1 && \
timeout 900 ncu --metrics $M --clock-control none -k regex:'ukf_forward|urtss_backward' --csv --log-file $O/r2c27_counts.csv $SMALL > $O/r2c27_ncu_counts.log 2>&1
echo "ncu counts rc $?"
timeout 300 $SMALLF > $O/r2c27_smallf_plain.log 2>&1 && \
timeout 900 ncu --metrics $M --clock-control none -k regex:'ukf_forward|urtss_backward' --csv --log-file $O/r2c27_counts_full.csv $SMALLF > $O/r2c27_ncu_counts_full.log 2>&1
echo "ncu counts full-cov rc $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ukf_forward -s 1 -c 1 -f -o $O/r2c27_prof_fwd $SMALL > $O/r2c27_ncu_full_fwd.log 2>&1
echo "ncu full fwd rc $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:urtss_backward -s 1 -c 1 -f -o $O/r2c27_prof_bwd $SMALL > $O/r2c27_ncu_full_bwd.log 2>&1
echo "ncu full bwd rc $?"
