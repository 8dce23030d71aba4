#!/bin/bash
# Round-2 call 7: merged-N k_lvc_p (SWIZZLE_64B kernel image) on hardware: quick parity, bench, role timeline, ncu of k_lvc_p, GPU suite.
set -u
OUT=gpurun_out/r2_c7
mkdir -p "$OUT"
timeout 300 python tests/gpu_lvcp_check.py > "$OUT/lvcp_check.log" 2>&1; echo "rc=$?" >> "$OUT/lvcp_check.log"
B="python bench.py --steps 10 --warmup 3"
timeout 300 $B --no-cpu > "$OUT/bench_default.json" 2> "$OUT/bench_default.err"
timeout 150 $B --no-cpu --batch 1 --frames 86 > "$OUT/bench_1x86.json" 2> "$OUT/bench_1x86.err"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_lvc_p -s 12 -c 2 -o "$OUT/lvcp" python bench.py --steps 1 --warmup 1 --no-cpu --opt graphs=0 > "$OUT/ncu_full.log" 2>&1
timeout 1200 python -m pytest tests -m gpu -q -x > "$OUT/gpu_tests.log" 2>&1; echo "pytest rc=$?" >> "$OUT/gpu_tests.log"
FD_NVCC_EXTRA="-DLP_TIMELINE=1" python -c "import __graft_entry__ as g; g.build_cuda(force=True)" > "$OUT/build.log" 2>&1
timeout 200 python tests/gpu_lp_timeline.py > "$OUT/lp_timeline.txt" 2>&1
python -c "import __graft_entry__ as g; g.build_cuda(force=True)" >> "$OUT/build.log" 2>&1
for f in "$OUT"/bench_*.json; do
  python - "$f" >> "$OUT/summary.txt" 2>&1 <<'PY'
import sys, json
try:
    j = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1].split('/')[-1], round(j['value'] / 1e6, 2), 'M samples/s', round(j['ms_per_step'], 3), 'ms', {k: round(v, 3) for k, v in j.get('kernel_ms_per_step', {}).items()}, 'roofline', j.get('roofline') and (j['roofline']['kernel'], round(j['roofline']['frac'], 4)))
except Exception as e:
    print(sys.argv[1], 'unparsed', e)
PY
done
