#!/bin/bash
# Round 2, GPU call 23 (8 GPUs): bench.py under torchrun as the driver launches it, final kernels: N = 8, 4, 2 back to back
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
for N in 8 4 2; do
  TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2952$N"
  timeout 600 $TR bench.py --gpus $N --steps 5 --warmup 3 --e2e-headline-only > $O/r2c23_bench_c5_${N}gpu.json 2> $O/r2c23_bench_c5_${N}gpu.err; echo "c5 x$N rc $?"
  python - <<PY
import json
l=[x for x in open("$O/r2c23_bench_c5_${N}gpu.json") if x.startswith("{")][-1]
d=json.loads(l)
print($N, "value %.4e"%d["value"], "ms/step %.2f"%d["ms_per_step"], "e2e %.3e"%d["e2e"]["value"], "job s", d.get("job",{}).get("wall_s"))
PY
done
