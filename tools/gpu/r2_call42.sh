#!/bin/bash
# Round 2, GPU call 42: bench c4 with its own kernel counts in the roofline block
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python bench.py --config c4 --steps 5 --warmup 3 > $O/r2c42_bench_c4.json 2> $O/r2c42_bench_c4.err; echo "c4 rc $?"
python - <<PY
import json
l=[x for x in open("$O/r2c42_bench_c4.json") if x.startswith("{")][-1]
d=json.loads(l); print("c4 value %.4e"%d["value"]); print({k:{a:(round(b,3) if isinstance(b,float) else b) for a,b in v.items()} for k,v in d["roofline"]["kernels"].items()})
PY
