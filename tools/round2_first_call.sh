#!/bin/bash
# First GPU call of round 2 (run from the repo root through gpurun, ~25 GPU-minutes):
#   gpurun --timeout 2400 -- 'bash tools/round2_first_call.sh'
# 1. the GPU test-suite on the default path; 2. parity + per-class kernel times of the run-time options that are OFF by default
# (tc_b0, b2_skipbuf, kc_stage, lvc_pipe; tests/gpu_options_check.py); 3. bench.py for the default and for each option; 4. one rebuild + bench per
# compile-time switch.  Everything lands in gpurun_out/r2_first/.  Each step has its own timeout: an experimental kernel that hangs
# costs its step, not the call.
set -u
OUT=gpurun_out/r2_first
mkdir -p "$OUT"
B="python bench.py --steps 10 --warmup 3 --no-cpu"

timeout 900 python -m pytest tests -m gpu -q > "$OUT/gpu_tests.log" 2>&1; echo "pytest rc=$?" >> "$OUT/gpu_tests.log"
for o in tc_b0 b2_skipbuf kc_stage lvc_pipe pipe_rows; do
    timeout 150 python tests/gpu_options_check.py $o > "$OUT/options_check_$o.log" 2>&1; echo "rc=$?" >> "$OUT/options_check_$o.log"
done

timeout 120 $B > "$OUT/bench_default.json" 2> "$OUT/bench_default.err"
for o in tc_b0=1 tc_b0=2 b2_skipbuf=1 kc_stage=1 lvc_pipe=1; do
    timeout 120 $B --opt $o > "$OUT/bench_$o.json" 2> "$OUT/bench_$o.err"; echo "$o rc=$?" >> "$OUT/bench_rc.log"
done
timeout 120 $B --opt b2_skipbuf=1 --opt lvc_pipe=1 > "$OUT/bench_pipe_rows.json" 2> "$OUT/bench_pipe_rows.err"
timeout 120 $B --opt tc_b0=1 --opt b2_skipbuf=1 --opt lvc_pipe=1 --opt kc_stage=1 > "$OUT/bench_all_options.json" 2> "$OUT/bench_all_options.err"

for d in LH_ROW_SPREAD FINAL_BATCH_LOADS FD_VEC256 LH_PREFETCH_EPI LH_NO_END_SYNC; do
    FD_NVCC_EXTRA="-D$d=1" python -c "import __graft_entry__ as g; g.build_cuda(force=True)" > "$OUT/build_$d.log" 2>&1 || continue
    timeout 120 $B > "$OUT/bench_$d.json" 2> "$OUT/bench_$d.err"; echo "$d rc=$?" >> "$OUT/bench_rc.log"
    timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "(tensor_core_mode_vs_oracle and tc_3xf16) or full_size" > "$OUT/parity_$d.log" 2>&1
done
python -c "import __graft_entry__ as g; g.build_cuda(force=True)" > "$OUT/build_default.log" 2>&1
grep -h '"value"' "$OUT"/bench_*.json | python -c "
import sys, json
for l in sys.stdin:
    try:
        j = json.loads(l); print(round(j['value'] / 1e6, 2), 'M samples/s', j.get('experiment', ''), j['ms_per_step'])
    except Exception as e:
        print('unparsed line', e)
" > "$OUT/summary.txt" 2>&1
ls "$OUT" > /dev/null
