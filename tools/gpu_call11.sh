#!/bin/bash
# Round-2 call 11: mbarrier suspend-time hint again, now that k_lvc_p is bound by its SIMT issue slots (per-class kernel times from gpu_lvcp_check.py).
set -u
OUT=gpurun_out/r2_c11
mkdir -p "$OUT"
for H in 0 200 1000 5000; do
  FD_NVCC_EXTRA="-DMBAR_HINT_NS=$H" python -c "import __graft_entry__ as g; g.build_cuda(force=True)" > "$OUT/build.log" 2>&1
  timeout 200 python tests/gpu_lvcp_check.py 2>&1 | grep -E "lvc_p=1|PARITY" > "$OUT/hint$H.txt"
  timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu > "$OUT/bench_hint$H.json" 2> /dev/null
done
python -c "import __graft_entry__ as g; g.build_cuda(force=True)" >> "$OUT/build.log" 2>&1
for H in 0 200 1000 5000; do echo "hint $H: $(cat $OUT/hint$H.txt | cut -c1-330)"; python -c "
import json,sys
j=json.loads(open('$OUT/bench_hint$H.json').read().strip().splitlines()[-1]); print('   bench', round(j['ms_per_step'],3), 'ms', {k: round(v,3) for k,v in j['kernel_ms_per_step'].items()})"; done > "$OUT/summary.txt" 2>&1
