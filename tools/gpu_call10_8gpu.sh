#!/bin/bash
# Round-2 call 10 (8 GPUs, one box): BASELINE.json configs[3] / configs[4] on 8 GPUs through the product API, and bench.py at N = 8.
set -u
OUT=gpurun_out/r2_c10
mkdir -p "$OUT"
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533"
timeout 600 $T tools/sweep.py --no-n1000 > "$OUT/sweep_8gpu.md" 2> "$OUT/sweep_8gpu.err"
timeout 300 $T bench.py --gpus 8 --steps 20 --warmup 3 > "$OUT/bench_8gpu.json" 2> "$OUT/bench_8gpu.err"
