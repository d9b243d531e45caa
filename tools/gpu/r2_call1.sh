#!/bin/bash
# Round 2, GPU call 1: parity suite, the three bench configs, kernel variants, ncu counts + one full capture.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $O/r2c1_smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q -s > $O/r2c1_pytest.log 2>&1; echo "pytest rc $?" >> $O/r2c1_pytest.log
tail -5 $O/r2c1_pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 > $O/r2c1_bench_c5.json 2> $O/r2c1_bench_c5.err; echo "c5 rc $?"
timeout 600 python bench.py --config c3 --steps 5 --warmup 3 > $O/r2c1_bench_c3.json 2> $O/r2c1_bench_c3.err; echo "c3 rc $?"
timeout 900 python bench.py --config c4 --steps 5 --warmup 3 > $O/r2c1_bench_c4.json 2> $O/r2c1_bench_c4.err; echo "c4 rc $?"
QP="python tools/quick_perf.py --tracks 113664 --steps 512 --packed --no-metrics"
timeout 300 $QP --label default > $O/r2c1_qp_default.log 2>&1
for v in b2 b4 pu2; do
  STE_UKF_LIB=$PWD/gpurun_in/variants/libste_$v.so timeout 300 $QP --no-probe --label $v > $O/r2c1_qp_$v.log 2>&1
done
grep -h fwd_ms $O/r2c1_qp_*.log | cut -c1-330
SMALL="python tools/quick_perf.py --tracks 75776 --steps 64 --packed --no-metrics --no-probe --reps 1"
M=$(python -c "import sys; sys.path.insert(0,'tools'); import ncu_counts; print(ncu_counts.METRICS)")
timeout 300 $SMALL > $O/r2c1_small_plain.log 2>&1 && \
timeout 900 ncu --metrics $M --clock-control none -k regex:'ukf_forward|urtss_backward' --csv --log-file $O/r2c1_counts.csv $SMALL > $O/r2c1_ncu_counts.log 2>&1
echo "ncu counts rc $?"
timeout 300 $SMALL > $O/r2c1_small_plain2.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ukf_forward -s 1 -c 1 -f -o $O/r2c1_prof_fwd $SMALL > $O/r2c1_ncu_full.log 2>&1
echo "ncu full rc $?"
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-job --e2e-headline-only > $O/r2c1_plain_bench.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'ukf_|urtss_|track_metrics' -c 40 --csv --log-file $O/r2c1_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-job --e2e-headline-only > $O/r2c1_ncu_launches.log 2>&1
echo "ncu launches rc $?"
ls -la $O | grep r2c1
