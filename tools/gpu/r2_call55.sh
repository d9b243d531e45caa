#!/bin/bash
# Round 2, GPU call 55: A/B of block-synchronous stepping (the warps of a block start every step together: __syncthreads_or in the time loop)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
QP="python tools/quick_perf.py --tracks 113664 --steps 512 --packed --no-metrics --no-probe"
for v in base sync; do
  if [ $v = sync ]; then export STE_UKF_LIB=$PWD/gpurun_in/libste_ukf_sync.so; else unset STE_UKF_LIB; fi
  timeout 300 $QP --label $v >> $O/r2c55_qp.log 2>&1
done
grep -h fwd_ms $O/r2c55_qp.log | cut -c1-120
for v in base sync base sync; do
  if [ $v = sync ]; then export STE_UKF_LIB=$PWD/gpurun_in/libste_ukf_sync.so; else unset STE_UKF_LIB; fi
  timeout 900 python bench.py --config c4 --steps 5 --warmup 3 --no-cpu-baseline --e2e-headline-only > $O/r2c55_bench_c4_$v.json 2> $O/r2c55_bench_c4_$v.err; echo "c4 $v rc $?"
  python - <<PY
import json
l=[x for x in open("$O/r2c55_bench_c4_$v.json") if x.startswith("{")][-1]
d=json.loads(l); print("c4", "$v", "value %.4e"%d["value"], d["roofline"]["forward_ms"], d["roofline"]["backward_ms"])
PY
done
unset STE_UKF_LIB
M=$(python -c "import sys; sys.path.insert(0,'tools'); import ncu_counts; print(ncu_counts.METRICS)"),sm__icc_request_hit_rate.pct,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio
STE_UKF_LIB=$PWD/gpurun_in/libste_ukf_sync.so timeout 900 ncu --metrics $M --clock-control none -k regex:'ukf_forward' --csv --log-file $O/r2c55_c4_counts_sync.csv python tools/c4_counts.py run > $O/r2c55_c4_ncu.log 2>&1; echo "ncu rc $?"
grep -h "icc_request_hit_rate\|no_instruction\|stalled_barrier\|gpu__time_duration" $O/r2c55_c4_counts_sync.csv | cut -d, -f13-15
