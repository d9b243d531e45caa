#!/bin/bash
# Round 2, GPU call 52: the round's last tree - parity suite, smoke, bench c5 (default), reference arm, c4
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > $O/r2c52_pytest.log 2>&1; echo "pytest rc $?"; tail -2 $O/r2c52_pytest.log | cut -c1-200
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
( time timeout 900 python bench.py > $O/r2c52_bench_c5.json 2> $O/r2c52_bench_c5.err ) 2> $O/r2c52_bench_c5.time; echo "bench rc $?"; grep real $O/r2c52_bench_c5.time
( time timeout 600 python bench.py --impl reference > $O/r2c52_ref.json 2> $O/r2c52_ref.err ) 2> $O/r2c52_ref.time; echo "ref rc $?"; grep real $O/r2c52_ref.time
timeout 900 python bench.py --config c4 --steps 5 --warmup 3 > $O/r2c52_bench_c4.json 2> $O/r2c52_bench_c4.err; echo "c4 rc $?"
for f in c5 c4; do python - <<PY
import json
l=[x for x in open("$O/r2c52_bench_$f.json") if x.startswith("{")][-1]
d=json.loads(l); r=d["roofline"]
print("$f", "value %.4e"%d["value"], "ms/step %.3f"%d["ms_per_step"], "frac", round(r.get("whole_step",{}).get("frac",r["frac"]),4), "fwd/bwd ms", r.get("forward_ms"), r.get("backward_ms"), "e2e %.3e"%d["e2e"]["value"], "job", d.get("job",{}).get("wall_s"), d["clocks"]["reasons"])
PY
done
python -c "
import json; r=json.loads(open('$O/r2c52_ref.json').read().strip().split('\n')[-1]); print('ref', r['value'], r['steps'], r['ms_per_step'], r['cpu_baseline']['sample'][:90])"
