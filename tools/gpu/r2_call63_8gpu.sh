#!/bin/bash
# Round 2, GPU call 63 (8 GPUs): default bench under the driver's torchrun launch on the round's last tree
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29563"
( time timeout 600 $TR bench.py --gpus 8 --steps 5 --warmup 3 > $O/r2c63_bench_c5_8gpu.json 2> $O/r2c63_bench_c5_8gpu.err ) 2> $O/r2c63.time; echo "c5 x8 rc $?"; grep real $O/r2c63.time
python - <<PY
import json
try:
    l=[x for x in open("$O/r2c63_bench_c5_8gpu.json") if x.startswith("{")]
    d=json.loads(l[-1]); p=d.get("partitioned_schedule") or {}
    print("lines", len(l), "value %.4e"%d["value"], "n_gpus", d["n_gpus"], "ms/step %.2f"%d["ms_per_step"], "job", d.get("job",{}).get("wall_s"), "e2e %.3e"%d["e2e"]["value"], "partitioned", {k:p[k] for k in p if k!="schedule"})
except Exception as e:
    print("failed", e); print(open("$O/r2c63_bench_c5_8gpu.err").read()[-1500:])
PY
