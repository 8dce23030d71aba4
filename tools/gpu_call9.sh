#!/bin/bash
# Round-2 call 9 (1 GPU): the measurements the round's documents quote -- GPU test-suite, bench (both arms), ncu launch list + full-set capture of one
# reverse step, the 1-GPU sweep of BASELINE.json's configs.
set -u
OUT=gpurun_out/r2_c9
mkdir -p "$OUT"
timeout 1500 python -m pytest tests -m gpu -q > "$OUT/gpu_tests.log" 2>&1; echo "pytest rc=$?" >> "$OUT/gpu_tests.log"
timeout 400 python bench.py --steps 20 --warmup 3 > "$OUT/bench_1gpu.json" 2> "$OUT/bench_1gpu.err"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > "$OUT/bench_reference.json" 2> "$OUT/bench_reference.err"
timeout 150 python bench.py --steps 20 --warmup 3 --no-cpu --batch 1 --frames 86 > "$OUT/bench_1x86.json" 2> "$OUT/bench_1x86.err"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file "$OUT/launches.csv" python bench.py --steps 1 --warmup 1 --no-cpu --opt graphs=0 > "$OUT/ncu_list.log" 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -s 120 -c 32 -o "$OUT/step" python bench.py --steps 1 --warmup 1 --no-cpu --opt graphs=0 > "$OUT/ncu_full.log" 2>&1
timeout 900 python tools/sweep.py > "$OUT/sweep_1gpu.md" 2> "$OUT/sweep_1gpu.err"
python __graft_entry__.py smoke > "$OUT/smoke.log" 2>&1; echo "smoke rc=$?" >> "$OUT/smoke.log"
