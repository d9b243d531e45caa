#!/bin/bash
# Round 2, GPU call 51: A/B of the single-copy gating loop (base = call 50 tree): a pure code-size change: uniform shape and config 4
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
QP="python tools/quick_perf.py --tracks 113664 --steps 512 --packed --no-metrics --no-probe"
for v in base new base new; do
  if [ $v = base ]; then export STE_UKF_LIB=$PWD/gpurun_in/libste_ukf_base.so; else unset STE_UKF_LIB; fi
  timeout 300 $QP --label $v >> $O/r2c51_qp.log 2>&1
done
grep -h fwd_ms $O/r2c51_qp.log | cut -c1-120
for v in base new; do
  if [ $v = base ]; then export STE_UKF_LIB=$PWD/gpurun_in/libste_ukf_base.so; else unset STE_UKF_LIB; fi
  timeout 900 python bench.py --config c4 --steps 5 --warmup 3 --no-cpu-baseline --e2e-headline-only > $O/r2c51_bench_c4_$v.json 2> $O/r2c51_bench_c4_$v.err; echo "c4 $v rc $?"
done
unset STE_UKF_LIB
python - <<PY
import json
for v in ("base","new"):
    l=[x for x in open("$O/r2c51_bench_c4_%s.json"%v) if x.startswith("{")][-1]
    d=json.loads(l); print(v, "value %.4e"%d["value"], d["roofline"]["forward_ms"], d["roofline"]["backward_ms"])
PY
