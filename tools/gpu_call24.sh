#!/bin/bash
# Round-2 call 24 (1 GPU, last of the round): k_upsample_p4 epilogue with 256-bit sector stores -- parity subset + A/B bench against a -DU4_STORE32=0 build.
set -u
OUT=gpurun_out/r2_c24
mkdir -p "$OUT"; rm -f "$OUT"/*
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "stages or tensor_core_mode or 8x861 or options_agree or boundaries" > "$OUT/gpu_tests.log" 2>&1; echo "rc=$?" >> "$OUT/gpu_tests.log"
B="python bench.py --steps 10 --warmup 3 --no-cpu"
timeout 120 $B > "$OUT/bench_a_store32.json" 2>/dev/null
FASTDIFF_B200_LIB=$PWD/fastdiff_b200/csrc/libfd_ab_u416.so timeout 120 $B > "$OUT/bench_b_store16.json" 2>/dev/null
for f in "$OUT"/bench_*.json; do
  python - "$f" >> "$OUT/summary.txt" 2>&1 <<'PY'
import sys, json
try:
    j = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1].split('/')[-1], round(j['value'] / 1e6, 2), 'M samples/s', round(j['ms_per_step'], 3), 'ms', 'e2e', round(j['e2e']['ms_per_step'], 3), j.get('clocks', {}).get('sm_mhz'), {k: round(v, 3) for k, v in j.get('kernel_ms_per_step', {}).items()})
except Exception as e:
    print(sys.argv[1], 'unparsed', e)
PY
done
