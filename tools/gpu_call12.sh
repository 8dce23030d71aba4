#!/bin/bash
# Round-2 call 12 (lean A/B): quick parity of the piece-row path + bench (+ an optional second bench with the options in $FD_AB_OPT).
set -u
OUT=gpurun_out/r2_c12
mkdir -p "$OUT"; rm -f "$OUT"/*
timeout 300 python tests/gpu_scripts/lvcp_check.py > "$OUT/lvcp_check.log" 2>&1; echo "rc=$?" >> "$OUT/lvcp_check.log"
B="python bench.py --steps 10 --warmup 3 --no-cpu"
timeout 300 $B > "$OUT/bench_default.json" 2> "$OUT/bench_default.err"
if [ -n "${FD_AB_OPT:-}" ]; then timeout 300 $B --opt $FD_AB_OPT > "$OUT/bench_ab.json" 2> "$OUT/bench_ab.err"; fi
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "boundaries or stages or tensor_core_mode or 8x861" > "$OUT/gpu_tests.log" 2>&1; echo "rc=$?" >> "$OUT/gpu_tests.log"
for f in "$OUT"/bench_*.json; do
  python - "$f" >> "$OUT/summary.txt" 2>&1 <<'PY'
import sys, json
try:
    j = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1].split('/')[-1], round(j['value'] / 1e6, 2), 'M samples/s', round(j['ms_per_step'], 3), 'ms', 'e2e', round(j['e2e']['ms_per_step'], 3), {k: round(v, 3) for k, v in j.get('kernel_ms_per_step', {}).items()})
except Exception as e:
    print(sys.argv[1], 'unparsed', e)
PY
done
