#!/bin/bash
# Round 2, GPU call 53: A/B of the Jacobi stage order (0,1)(2,3) -> (0,3)(1,2) -> (0,2)(1,3) (base = call 52 tree), then the parity suite
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
QP="python tools/quick_perf.py --tracks 113664 --steps 512 --packed --no-metrics --no-probe"
for v in base new base new; do
  if [ $v = base ]; then export STE_UKF_LIB=$PWD/gpurun_in/libste_ukf_base.so; else unset STE_UKF_LIB; fi
  timeout 300 $QP --label $v >> $O/r2c53_qp.log 2>&1
done
grep -h fwd_ms $O/r2c53_qp.log | cut -c1-120
for v in base new; do
  if [ $v = base ]; then export STE_UKF_LIB=$PWD/gpurun_in/libste_ukf_base.so; else unset STE_UKF_LIB; fi
  timeout 900 python bench.py --config c4 --steps 5 --warmup 3 --no-cpu-baseline --e2e-headline-only > $O/r2c53_bench_c4_$v.json 2> $O/r2c53_bench_c4_$v.err; echo "c4 $v rc $?"
  timeout 600 python bench.py --config c2 --steps 5 --warmup 3 --no-cpu-baseline > $O/r2c53_bench_c2_$v.json 2> $O/r2c53_bench_c2_$v.err; echo "c2 $v rc $?"
done
unset STE_UKF_LIB
python - <<PY
import json
for c in ("c4","c2"):
  for v in ("base","new"):
    l=[x for x in open("$O/r2c53_bench_%s_%s.json"%(c,v)) if x.startswith("{")][-1]
    d=json.loads(l); print(c, v, "value %.4e"%d["value"], d["roofline"]["forward_ms"], d["roofline"]["backward_ms"], d.get("parity_vs_reference"))
PY
timeout 2400 python -m pytest tests -m gpu -q > $O/r2c53_pytest.log 2>&1; echo "pytest rc $?"; tail -2 $O/r2c53_pytest.log | cut -c1-200
