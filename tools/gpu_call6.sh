#!/bin/bash
# Round-2 call 6: store-pattern microbenchmark (GEMM epilogue), mbarrier suspend-hint sweep, k_lvc_p role timeline.
set -u
OUT=gpurun_out/r2_c6
mkdir -p "$OUT"
( cd tests/microbench && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o store_patterns store_patterns.cu && timeout 120 ./store_patterns ) > "$OUT/store_patterns.txt" 2>&1
B="python bench.py --steps 10 --warmup 3 --no-cpu"
rebuild() { FD_NVCC_EXTRA="$1" python -c "import __graft_entry__ as g; g.build_cuda(force=True)" >> "$OUT/build.log" 2>&1; }
for H in 0 300 1000 4000 20000; do
  rebuild "-DMBAR_HINT_NS=$H"
  timeout 200 $B > "$OUT/bench_hint$H.json" 2> "$OUT/bench_hint$H.err"
  timeout 100 $B --batch 1 --frames 86 > "$OUT/bench_1x86_hint$H.json" 2> "$OUT/bench_1x86_hint$H.err"
done
rebuild "-DLP_TIMELINE=1"
timeout 200 python tests/gpu_lp_timeline.py > "$OUT/lp_timeline_h0.txt" 2>&1
rebuild "-DLP_TIMELINE=1 -DMBAR_HINT_NS=1000"
timeout 200 python tests/gpu_lp_timeline.py > "$OUT/lp_timeline_h1000.txt" 2>&1
rebuild ""
for f in "$OUT"/bench_*.json; do
  python - "$f" >> "$OUT/summary.txt" 2>&1 <<'PY'
import sys, json
try:
    j = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1].split('/')[-1], round(j['value'] / 1e6, 2), 'M samples/s', round(j['ms_per_step'], 3), 'ms', {k: round(v, 3) for k, v in j.get('kernel_ms_per_step', {}).items()})
except Exception as e:
    print(sys.argv[1], 'unparsed', e)
PY
done
