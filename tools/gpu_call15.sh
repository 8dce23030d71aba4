#!/bin/bash
# Round-2 call 15: kernel_conv GEMM, resident-frame-tile form (kc_res = 1) -- parity, A/B bench against the whole-stage ring (kc_res = 0) and against
# a -DKC_EARLY_RELEASE=0 build on one box, role timeline.
#   (build here first:  nvcc ... -DKC_EARLY_RELEASE=0 -o fastdiff_b200/csrc/libfd_ab_norel.so ; nvcc ... -DKC_TIMELINE=1 -o fastdiff_b200/csrc/libfd_ab_tl.so)
set -u
OUT=gpurun_out/r2_c15
mkdir -p "$OUT"; rm -f "$OUT"/*
timeout 500 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "forms or boundaries or options_agree or stages or tensor_core_mode or 8x861" > "$OUT/gpu_tests.log" 2>&1; echo "rc=$?" >> "$OUT/gpu_tests.log"
B="python bench.py --steps 10 --warmup 3 --no-cpu"
NOREL=$PWD/fastdiff_b200/csrc/libfd_ab_norel.so
timeout 200 $B > "$OUT/bench_a1_res.json" 2>/dev/null
timeout 200 $B --opt kc_res=0 > "$OUT/bench_b1_ring.json" 2>/dev/null
FASTDIFF_B200_LIB=$NOREL timeout 200 $B > "$OUT/bench_c1_res_norel.json" 2>/dev/null
timeout 200 $B > "$OUT/bench_a2_res.json" 2>/dev/null
timeout 200 $B --opt kc_res=0 > "$OUT/bench_b2_ring.json" 2>/dev/null
FASTDIFF_B200_LIB=$PWD/fastdiff_b200/csrc/libfd_ab_tl.so timeout 200 python tests/gpu_scripts/kc_timeline.py > "$OUT/kc_timeline.txt" 2>&1
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
