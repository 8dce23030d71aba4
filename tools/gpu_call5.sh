#!/bin/bash
# Round-2 call 5: k_lvc_p with the critical-path roles on the high warp ids, bounded-poll loops not unrolled; sweep tool (1 GPU).
set -u
OUT=gpurun_out/r2_c5
mkdir -p "$OUT"

timeout 300 python tests/gpu_lvcp_check.py > "$OUT/lvcp_check.log" 2>&1; echo "rc=$?" >> "$OUT/lvcp_check.log"
if ! grep -q "PARITY OK" "$OUT/lvcp_check.log"; then echo "quick parity failed" > "$OUT/summary.txt"; tail -30 "$OUT/lvcp_check.log" >> "$OUT/summary.txt"; fi
timeout 1200 python -m pytest tests -m gpu -q > "$OUT/gpu_tests.log" 2>&1; echo "pytest rc=$?" >> "$OUT/gpu_tests.log"
B="python bench.py --steps 10 --warmup 3"
timeout 300 $B > "$OUT/bench_default.json" 2> "$OUT/bench_default.err"
timeout 150 $B --no-cpu --opt graphs=0 > "$OUT/bench_nograph.json" 2> "$OUT/bench_nograph.err"
timeout 150 $B --no-cpu --batch 1 --frames 86 > "$OUT/bench_1x86.json" 2> "$OUT/bench_1x86.err"
timeout 150 $B --no-cpu --batch 1 --frames 86 --opt graphs=0 > "$OUT/bench_1x86_nograph.json" 2> "$OUT/bench_1x86_nograph.err"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > "$OUT/bench_reference.json" 2> "$OUT/bench_reference.err"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file "$OUT/launches.csv" python bench.py --steps 1 --warmup 1 --no-cpu --opt graphs=0 > "$OUT/ncu_list.log" 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_lvc_p -s 12 -c 2 -o "$OUT/lvcp" python bench.py --steps 1 --warmup 1 --no-cpu --opt graphs=0 > "$OUT/ncu_full.log" 2>&1
timeout 400 python tools/sweep.py --no-n1000 > "$OUT/sweep_1gpu.md" 2> "$OUT/sweep_1gpu.err"
grep -h '"value"' "$OUT"/bench_*.json | python -c "
import sys, json
for l in sys.stdin:
    try:
        j = json.loads(l); print(round(j['value'] / 1e6, 2), 'M samples/s', j.get('experiment', ''), j.get('impl',''), round(j['ms_per_step'], 3), 'e2e', (j.get('e2e') or {}).get('ms_per_step'), 'e2e_ref_noise', (j.get('e2e_reference_noise') or {}).get('ms_per_step'), {k: round(v, 3) for k, v in j.get('kernel_ms_per_step', {}).items()})
    except Exception as e:
        print('unparsed line', e)
" >> "$OUT/summary.txt" 2>&1
