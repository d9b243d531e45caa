#!/bin/bash
# Round 2, GPU call 56 (2 GPUs): the driver's launch of both arms on the round's last tree (torchrun, defaults)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29556"
( time timeout 600 $TR bench.py --impl reference --gpus 2 --steps 5 --warmup 3 > $O/r2c56_ref_2gpu.json 2> $O/r2c56_ref_2gpu.err ) 2> $O/r2c56_ref.time; echo "ref x2 rc $?"; grep real $O/r2c56_ref.time
( time timeout 900 $TR bench.py --gpus 2 --steps 5 --warmup 3 > $O/r2c56_bench_c5_2gpu.json 2> $O/r2c56_bench_c5_2gpu.err ) 2> $O/r2c56_c5.time; echo "c5 x2 rc $?"; grep real $O/r2c56_c5.time
timeout 900 $TR bench.py --gpus 2 --config c4 --steps 4 --warmup 3 > $O/r2c56_bench_c4_2gpu.json 2> $O/r2c56_bench_c4_2gpu.err; echo "c4 x2 rc $?"
python - <<PY
import json
for f in ("bench_c5","bench_c4","ref"):
    try:
        l=[x for x in open("$O/r2c56_%s_2gpu.json"%f) if x.startswith("{")]
        d=json.loads(l[-1]); print(f, "lines", len(l), "value %.4e"%d["value"], "n_gpus", d["n_gpus"], "ms/step %.2f"%d["ms_per_step"], "e2e %.3e"%d["e2e"]["value"], d.get("job",{}).get("wall_s"))
    except Exception as e:
        print(f, "failed", e); print(open("$O/r2c56_%s_2gpu.err"%f).read()[-1500:])
PY
