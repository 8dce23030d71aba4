#!/bin/bash
# Round-2 call 16 (1 GPU): verification of HEAD -- the whole GPU suite, smoke, the bench line and the 1 s utterance.
set -u
OUT=gpurun_out/r2_c16
mkdir -p "$OUT"; rm -f "$OUT"/*
timeout 1500 python -m pytest tests -m gpu -q > "$OUT/gpu_tests.log" 2>&1; echo "pytest rc=$?" >> "$OUT/gpu_tests.log"
python __graft_entry__.py smoke > "$OUT/smoke.log" 2>&1; echo "smoke rc=$?" >> "$OUT/smoke.log"
timeout 400 python bench.py --steps 20 --warmup 3 > "$OUT/bench_1gpu.json" 2> "$OUT/bench_1gpu.err"
timeout 150 python bench.py --steps 20 --warmup 3 --no-cpu --batch 1 --frames 86 > "$OUT/bench_1x86.json" 2> "$OUT/bench_1x86.err"
