#!/bin/bash
# Round 2, GPU call 47: evidence of the round's last tree - counts, full captures of both kernels, launch list of the bench command
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
SMALL="python tools/quick_perf.py --tracks 75776 --steps 64 --packed --no-metrics --no-probe --reps 1"
SMALLF="python tools/quick_perf.py --tracks 75776 --steps 64 --no-metrics --no-probe --reps 1"
M=$(python -c "import sys; sys.path.insert(0,'tools'); import ncu_counts; print(ncu_counts.METRICS)")
timeout 300 $SMALL > $O/r2c47_small_plain.log 2>&1 && \
timeout 900 ncu --metrics $M --clock-control none -k regex:'ukf_forward|urtss_backward' --csv --log-file $O/r2c47_counts.csv $SMALL > $O/r2c47_ncu_counts.log 2>&1
echo "ncu counts rc $?"
timeout 300 $SMALLF > $O/r2c47_smallf_plain.log 2>&1 && \
timeout 900 ncu --metrics $M --clock-control none -k regex:'ukf_forward|urtss_backward' --csv --log-file $O/r2c47_counts_full.csv $SMALLF > $O/r2c47_ncu_counts_full.log 2>&1
echo "ncu counts full-cov rc $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ukf_forward -s 1 -c 1 -f -o $O/r2c47_prof_fwd $SMALL > $O/r2c47_ncu_full_fwd.log 2>&1
echo "ncu full fwd rc $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:urtss_backward -s 1 -c 1 -f -o $O/r2c47_prof_bwd $SMALL > $O/r2c47_ncu_full_bwd.log 2>&1
echo "ncu full bwd rc $?"
BL="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-job --e2e-headline-only"
timeout 300 $BL > $O/r2c47_plain_bench.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'ukf_|urtss_|track_metrics' -c 40 --csv --log-file $O/r2c47_launches.csv $BL > $O/r2c47_ncu_launches.log 2>&1
echo "ncu launches rc $?"
timeout 300 python tools/quick_perf.py --tracks 113664 --steps 512 --packed --no-metrics --no-probe --label final > $O/r2c47_qp.log 2>&1; grep -h fwd_ms $O/r2c47_qp.log | cut -c1-200
