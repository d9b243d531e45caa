#!/bin/bash
# Round 2, GPU call 40: shorter offset series for tiny latitude / distance offsets, hypot excess without its e^6 term
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
QP="python tools/quick_perf.py --tracks 113664 --steps 512 --packed --no-metrics --no-probe"
rm -f $O/r2c40_qp.log
for v in default hyp default hyp; do
  if [ $v = default ]; then timeout 300 $QP --label $v >> $O/r2c40_qp.log 2>&1
  else STE_UKF_LIB=$PWD/gpurun_in/variants/libste_$v.so timeout 300 $QP --label $v >> $O/r2c40_qp.log 2>&1; fi
done
grep -h fwd_ms $O/r2c40_qp.log | cut -c1-130
STE_UKF_LIB=$PWD/gpurun_in/variants/libste_hyp.so timeout 900 python bench.py --config c4 --steps 5 --warmup 3 --no-cpu-baseline --e2e-headline-only > $O/r2c40_bench_c4.json 2> $O/r2c40_bench_c4.err; echo "c4 rc $?"
python - <<PY
import json
l=[x for x in open("$O/r2c40_bench_c4.json") if x.startswith("{")][-1]
d=json.loads(l); print("c4 value %.4e"%d["value"], d["roofline"]["forward_ms"], d["roofline"]["backward_ms"])
PY
