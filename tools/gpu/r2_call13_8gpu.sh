#!/bin/bash
# Round 2, GPU call 13 (8 GPUs): bench.py under torchrun as the driver launches it.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519"
timeout 900 $TR bench.py --gpus 8 --steps 5 --warmup 3 > $O/r2c13_bench_c5_8gpu.json 2> $O/r2c13_bench_c5_8gpu.err; echo "c5 x8 rc $?"
tail -c 1500 $O/r2c13_bench_c5_8gpu.json; echo; grep -v "OMP_NUM_THREADS\|\*\*\*\*" $O/r2c13_bench_c5_8gpu.err | tail -5
