#!/bin/bash
# Round 2, GPU call 11: final record - parity suite, three bench configs, counts of the final kernels.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -s > $O/r2c11_pytest.log 2>&1; echo "pytest rc $?" >> $O/r2c11_pytest.log
grep -v "^  " $O/r2c11_pytest.log | tail -4 | cut -c1-300
SMALL="python tools/quick_perf.py --tracks 75776 --steps 64 --packed --no-metrics --no-probe --reps 1"
SMALLF="python tools/quick_perf.py --tracks 75776 --steps 64 --no-metrics --no-probe --reps 1"
M=$(python -c "import sys; sys.path.insert(0,'tools'); import ncu_counts; print(ncu_counts.METRICS)")
timeout 300 $SMALL > $O/r2c11_small_plain.log 2>&1 && \
timeout 900 ncu --metrics $M --clock-control none -k regex:'ukf_forward|urtss_backward' --csv --log-file $O/r2c11_counts.csv $SMALL > $O/r2c11_ncu_counts.log 2>&1
echo "ncu counts rc $?"
timeout 300 $SMALLF > $O/r2c11_smallf_plain.log 2>&1 && \
timeout 900 ncu --metrics $M --clock-control none -k regex:'ukf_forward|urtss_backward' --csv --log-file $O/r2c11_counts_full.csv $SMALLF > $O/r2c11_ncu_counts_full.log 2>&1
echo "ncu counts full-cov rc $?"
timeout 300 $SMALL > $O/r2c11_small_plain3.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:urtss_backward -s 1 -c 1 -f -o $O/r2c11_prof_bwd $SMALL > $O/r2c11_ncu_full_bwd.log 2>&1
echo "ncu full bwd rc $?"
BL="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-job --e2e-headline-only"
timeout 300 $BL > $O/r2c11_plain_bench.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'ukf_|urtss_|track_metrics' -c 40 --csv --log-file $O/r2c11_launches.csv $BL > $O/r2c11_ncu_launches.log 2>&1
echo "ncu launches rc $?"
timeout 600 python bench.py --config c3 --steps 5 --warmup 3 > $O/r2c11_bench_c3.json 2> $O/r2c11_bench_c3.err; echo "c3 rc $?"
timeout 900 python bench.py --config c4 --steps 5 --warmup 3 > $O/r2c11_bench_c4.json 2> $O/r2c11_bench_c4.err; echo "c4 rc $?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/r2c11_ref.json 2> $O/r2c11_ref.err; echo "ref rc $?"
