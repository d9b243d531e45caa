#!/bin/bash
# Round 2, GPU call 31: bench.py --config c2 (the reference's batch example on its own ship data), single-step class API cost
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python bench.py --config c2 --steps 5 --warmup 3 > $O/r2c31_bench_c2.json 2> $O/r2c31_bench_c2.err; echo "c2 rc $?"; tail -3 $O/r2c31_bench_c2.err
cut -c1-3000 $O/r2c31_bench_c2.json | tail -1
timeout 300 python tools/class_api_probe.py > $O/r2c31_class_api.log 2>&1; tail -1 $O/r2c31_class_api.log
