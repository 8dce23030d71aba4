OUT=gpurun_out/r2_pdl; mkdir -p $OUT; rm -f $OUT/*
B="python bench.py --steps 10 --warmup 3 --no-cpu"
for cfg in "graphs=1 pdl=0" "graphs=0 pdl=0" "graphs=0 pdl=1" "graphs=1 pdl=1"; do
  set -- $cfg
  n=$(echo $cfg | tr ' =' '__')
  timeout 200 $B --opt $1 --opt $2 > $OUT/b_$n.json 2>/dev/null
  timeout 100 $B --batch 1 --frames 86 --opt $1 --opt $2 > $OUT/s_$n.json 2>/dev/null
done
for f in $OUT/*.json; do python -c "
import json,sys
j=json.loads(open('$f').read().strip().splitlines()[-1]); print('$f'.split('/')[-1], round(j['ms_per_step'],3), 'e2e', round(j['e2e']['ms_per_step'],3))"; done > $OUT/summary.txt
