#!/bin/bash
# Round-2 call 2: the piece-row LVC path (k_lvc_p) on hardware: quick parity, full GPU suite, bench with the option on / off, ncu.
set -u
OUT=gpurun_out/r2_c2
mkdir -p "$OUT"
timeout 300 python tests/gpu_lvcp_check.py > "$OUT/lvcp_check.log" 2>&1; echo "rc=$?" >> "$OUT/lvcp_check.log"
if ! grep -q "PARITY OK" "$OUT/lvcp_check.log"; then echo "quick parity failed; stopping" > "$OUT/summary.txt"; tail -30 "$OUT/lvcp_check.log" >> "$OUT/summary.txt"; exit 1; fi
timeout 900 python -m pytest tests -m gpu -q -x > "$OUT/gpu_tests.log" 2>&1; echo "pytest rc=$?" >> "$OUT/gpu_tests.log"
B="python bench.py --steps 10 --warmup 3 --no-cpu"
timeout 150 $B > "$OUT/bench_default.json" 2> "$OUT/bench_default.err"
timeout 150 $B --opt lvc_p=0 > "$OUT/bench_lvc_p0.json" 2> "$OUT/bench_lvc_p0.err"
timeout 150 python bench.py --steps 10 --warmup 3 --no-cpu --batch 1 --frames 86 > "$OUT/bench_1x86.json" 2> "$OUT/bench_1x86.err"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file "$OUT/launches.csv" python bench.py --steps 1 --warmup 1 --no-cpu > "$OUT/ncu_list.log" 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_lvc_p -s 12 -c 2 -o "$OUT/lvcp" python bench.py --steps 1 --warmup 1 --no-cpu > "$OUT/ncu_full.log" 2>&1
grep -h '"value"' "$OUT"/bench_*.json | python -c "
import sys, json
for l in sys.stdin:
    try:
        j = json.loads(l); print(round(j['value'] / 1e6, 2), 'M samples/s', j.get('experiment', ''), round(j['ms_per_step'], 3), {k: round(v, 3) for k, v in j['kernel_ms_per_step'].items()})
    except Exception as e:
        print('unparsed line', e)
" > "$OUT/summary.txt" 2>&1
