#!/bin/bash
# Round 2, GPU call 62: is the partitioned schedule power-limited at the bench shape?  Probe at 151 552 x 1 024 with clocks / power sampled
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --id=0 --query-gpu=timestamp,clocks.sm,power.draw,clocks_event_reasons.sw_power_cap --format=csv,noheader,nounits -lms 20 > $O/r2c62_smi.csv 2>/dev/null &
SMI=$!
timeout 600 python tools/green_ctx_probe.py --tracks 151552 --steps 1024 --splits 48 --reps 8 > $O/r2c62_green.log 2>&1; echo "rc $?"
kill $SMI
grep "^{" $O/r2c62_green.log | cut -c1-300
python - <<PY
import csv
rows=[r for r in csv.reader(open("$O/r2c62_smi.csv")) if len(r)>=4]
clk=[float(r[1]) for r in rows]; pw=[float(r[2]) for r in rows]
busy=[(c,p,r[3].strip()) for c,p,r in zip(clk,pw,rows) if p>400]
print("samples", len(rows), "under load", len(busy))
if busy:
    import statistics
    print("clock min/median/max under load", min(b[0] for b in busy), statistics.median(b[0] for b in busy), max(b[0] for b in busy))
    print("power median/max under load", statistics.median(b[1] for b in busy), max(b[1] for b in busy), "power-cap active share", sum(b[2]=="Active" for b in busy)/len(busy))
PY
