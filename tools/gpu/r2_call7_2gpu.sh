#!/bin/bash
# Round 2, GPU call 7 (2 GPUs): bench.py under torchrun as the driver launches it - c5 with its whole-job pass, c4, reference arm.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517"
timeout 900 $TR bench.py --gpus 2 --steps 5 --warmup 3 > $O/r2c7_bench_c5_2gpu.json 2> $O/r2c7_bench_c5_2gpu.err; echo "c5 x2 rc $?"
timeout 900 $TR bench.py --gpus 2 --config c4 --steps 4 --warmup 3 > $O/r2c7_bench_c4_2gpu.json 2> $O/r2c7_bench_c4_2gpu.err; echo "c4 x2 rc $?"
timeout 600 $TR bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > $O/r2c7_ref_2gpu.json 2> $O/r2c7_ref_2gpu.err; echo "ref x2 rc $?"
tail -c 600 $O/r2c7_bench_c5_2gpu.json; echo; tail -c 300 $O/r2c7_bench_c5_2gpu.err; tail -c 300 $O/r2c7_bench_c4_2gpu.err; tail -c 400 $O/r2c7_ref_2gpu.json
