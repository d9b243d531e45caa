#!/bin/bash
# Round 2, GPU call 61 (2 GPUs): default bench under the driver's torchrun launch with the partitioned_schedule key; then one GPU
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561"
( time timeout 900 $TR bench.py --gpus 2 --steps 5 --warmup 3 > $O/r2c61_bench_c5_2gpu.json 2> $O/r2c61_bench_c5_2gpu.err ) 2> $O/r2c61.time; echo "c5 x2 rc $?"; grep real $O/r2c61.time
timeout 600 python bench.py --no-cpu-baseline --no-job --e2e-headline-only > $O/r2c61_bench_c5_1gpu.json 2> $O/r2c61_bench_c5_1gpu.err; echo "c5 x1 rc $?"
python - <<PY
import json
for f in ("2gpu","1gpu"):
    try:
        l=[x for x in open("$O/r2c61_bench_c5_%s.json"%f) if x.startswith("{")]
        d=json.loads(l[-1]); p=d.get("partitioned_schedule") or {}
        print(f, "lines", len(l), "value %.4e"%d["value"], "n_gpus", d["n_gpus"], "ms/step %.2f"%d["ms_per_step"], "job", d.get("job",{}).get("wall_s"), "partitioned", {k:p[k] for k in p if k!="schedule"})
    except Exception as e:
        print(f, "failed", e); print(open("$O/r2c61_bench_c5_%s.err"%f).read()[-1500:])
PY
