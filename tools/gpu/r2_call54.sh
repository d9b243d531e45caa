#!/bin/bash
# Round 2, GPU call 54: records of the round's last tree (after the stage-order change of call 53, whose call ran the parity suite):
# smoke, bench c5 (default) / reference arm / c3 / c4 / c2, config-4 counts, launch list of the bench command
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py > $O/r2c54_bench_c5.json 2> $O/r2c54_bench_c5.err; echo "bench rc $?"
timeout 600 python bench.py --impl reference > $O/r2c54_ref.json 2> $O/r2c54_ref.err; echo "ref rc $?"
timeout 600 python bench.py --config c3 --steps 5 --warmup 3 > $O/r2c54_bench_c3.json 2> $O/r2c54_bench_c3.err; echo "c3 rc $?"
M=$(python -c "import sys; sys.path.insert(0,'tools'); import ncu_counts; print(ncu_counts.METRICS)")
timeout 300 python tools/c4_counts.py run > $O/r2c54_c4_plain.log 2>&1; tail -1 $O/r2c54_c4_plain.log
timeout 900 ncu --metrics $M --clock-control none -k regex:'ukf_forward|urtss_backward' --csv --log-file $O/r2c54_c4_counts.csv python tools/c4_counts.py run > $O/r2c54_c4_ncu.log 2>&1; echo "ncu c4 counts rc $?"
timeout 900 python bench.py --config c4 --steps 5 --warmup 3 > $O/r2c54_bench_c4.json 2> $O/r2c54_bench_c4.err; echo "c4 rc $?"
timeout 600 python bench.py --config c2 --steps 5 --warmup 3 > $O/r2c54_bench_c2.json 2> $O/r2c54_bench_c2.err; echo "c2 rc $?"
BL="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-job --e2e-headline-only"
timeout 300 $BL > $O/r2c54_plain_bench.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'ukf_|urtss_|track_metrics' -c 40 --csv --log-file $O/r2c54_launches.csv $BL > $O/r2c54_ncu_launches.log 2>&1
echo "ncu launches rc $?"
for f in c5 c3 c4 c2; do python - <<PY
import json
l=[x for x in open("$O/r2c54_bench_$f.json") if x.startswith("{")][-1]
d=json.loads(l); r=d["roofline"]
print("$f", "value %.4e"%d["value"], "ms/step %.3f"%d["ms_per_step"], "frac", round(r.get("whole_step",{}).get("frac",r["frac"]),4), "fwd/bwd ms", r.get("forward_ms"), r.get("backward_ms"), "e2e %.3e"%d["e2e"]["value"], "job", d.get("job",{}).get("wall_s"), d["clocks"]["reasons"])
PY
done
python -c "
import json; r=json.loads(open('$O/r2c54_ref.json').read().strip().split('\n')[-1]); print('ref', r['value'], r['steps'], r['ms_per_step'])"
