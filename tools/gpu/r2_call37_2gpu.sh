#!/bin/bash
# Round 2, GPU call 37 (2 GPUs): configs 2 and 4 under torchrun with the final tree
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29537"
timeout 600 $TR bench.py --gpus 2 --config c2 --steps 5 --warmup 3 > $O/r2c37_bench_c2_2gpu.json 2> $O/r2c37_bench_c2_2gpu.err; echo "c2 x2 rc $?"
timeout 900 $TR bench.py --gpus 2 --config c4 --steps 4 --warmup 3 > $O/r2c37_bench_c4_2gpu.json 2> $O/r2c37_bench_c4_2gpu.err; echo "c4 x2 rc $?"
timeout 600 $TR bench.py --gpus 2 --config c3 --steps 5 --warmup 3 > $O/r2c37_bench_c3_2gpu.json 2> $O/r2c37_bench_c3_2gpu.err; echo "c3 x2 rc $?"
for f in c2 c4 c3; do python - <<PY
import json
try:
    l=[x for x in open("$O/r2c37_bench_${f}_2gpu.json") if x.startswith("{")][-1]
    d=json.loads(l); print("$f x2 value %.4e"%d["value"], "ms/step %.2f"%d["ms_per_step"], "e2e %.3e"%d["e2e"]["value"], d.get("parity_vs_reference"))
except Exception as e:
    print("$f failed", e); print(open("$O/r2c37_bench_${f}_2gpu.err").read()[-1500:])
PY
done
