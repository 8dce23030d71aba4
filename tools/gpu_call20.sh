#!/bin/bash
# Round-2 call 20 (1 GPU): streaming (evict-first) stores in the kernel_conv GEMM adopted -- GPU suite, bench, sustained power profile.
set -u
OUT=gpurun_out/r2_c20
mkdir -p "$OUT"; rm -f "$OUT"/*
timeout 900 python -m pytest tests -m gpu -q -x > "$OUT/gpu_tests.log" 2>&1; echo "pytest rc=$?" >> "$OUT/gpu_tests.log"
timeout 400 python bench.py --steps 20 --warmup 3 > "$OUT/bench_1gpu.json" 2> "$OUT/bench_1gpu.err"
timeout 120 python tests/gpu_scripts/power_profile.py 2>&1 | grep -v "Warn\|WeightNorm" > "$OUT/power_profile.txt"
