#!/bin/bash
# Round 2, GPU call 28: bench record of the trimmed kernels (c5 default run, launch list, c3, c4, reference arm)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
( time timeout 900 python bench.py > $O/r2c28_bench_c5.json 2> $O/r2c28_bench_c5.err ) 2> $O/r2c28_bench_c5.time; echo "c5 rc $?"; cat $O/r2c28_bench_c5.time | tr '\n' ' '; echo
BL="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-job --e2e-headline-only"
timeout 300 $BL > $O/r2c28_plain_bench.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'ukf_|urtss_|track_metrics' -c 40 --csv --log-file $O/r2c28_launches.csv $BL > $O/r2c28_ncu_launches.log 2>&1
echo "ncu launches rc $?"
timeout 600 python bench.py --config c3 --steps 5 --warmup 3 > $O/r2c28_bench_c3.json 2> $O/r2c28_bench_c3.err; echo "c3 rc $?"
timeout 900 python bench.py --config c4 --steps 5 --warmup 3 > $O/r2c28_bench_c4.json 2> $O/r2c28_bench_c4.err; echo "c4 rc $?"
( time timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/r2c28_ref.json 2> $O/r2c28_ref.err ) 2> $O/r2c28_ref.time; echo "ref rc $?"; cat $O/r2c28_ref.time | tr '\n' ' '; echo
for f in c5 c3 c4; do python - <<PY
import json
l=[x for x in open("$O/r2c28_bench_$f.json") if x.startswith("{")][-1]
d=json.loads(l); r=d["roofline"]
print("$f", "value %.3e"%d["value"], "ms/step %.2f"%d["ms_per_step"], "frac", round(r.get("whole_step",{}).get("frac",r["frac"]),4), "fwd/bwd ms", r.get("forward_ms"), r.get("backward_ms"), "e2e %.3e"%d["e2e"]["value"], "job", d.get("job",{}).get("wall_s"))
PY
done
