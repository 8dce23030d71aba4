#!/bin/bash
# Round-2 call 8: k_lvc_p with the storer warp + 4 stages; role timelines of k_lvc_p and of the kernel_conv GEMM.
set -u
OUT=gpurun_out/r2_c8
mkdir -p "$OUT"
timeout 300 python tests/gpu_scripts/lvcp_check.py > "$OUT/lvcp_check.log" 2>&1; echo "rc=$?" >> "$OUT/lvcp_check.log"
B="python bench.py --steps 10 --warmup 3"
timeout 300 $B --no-cpu > "$OUT/bench_default.json" 2> "$OUT/bench_default.err"
FD_NVCC_EXTRA="-DLP_TIMELINE=1 -DKC_TIMELINE=1 -DUT_TIMELINE=1" python -c "import __graft_entry__ as g; g.build_cuda(force=True)" > "$OUT/build.log" 2>&1
timeout 200 python tests/gpu_scripts/lp_timeline.py > "$OUT/lp_timeline.txt" 2>&1
timeout 200 python tests/gpu_scripts/kc_timeline.py 0 > "$OUT/kc_timeline_serial.txt" 2>&1
timeout 200 python tests/gpu_scripts/kc_timeline.py 1 > "$OUT/kc_timeline_overlap.txt" 2>&1
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
