#!/bin/bash
# Energy A/B of compile-time variants of the library on one box: tests/gpu_scripts/power_profile.py for each lib
#   fastdiff_b200/csrc/libfd_ab_<variant>.so (built here with the switch under test; "default*" = the in-tree library).
#   VARIANTS="default noef default2" PREFIXES="2 3 4 5" bash tools/gpu_energy_ab.sh
set -u
OUT=gpurun_out/r2_energy; mkdir -p $OUT; rm -f $OUT/*
for v in ${VARIANTS:-default}; do
  L=$PWD/fastdiff_b200/csrc/libfd_ab_$v.so; [ -f $L ] || L=$PWD/fastdiff_b200/csrc/libfastdiff_b200.so
  echo "== $v" >> $OUT/energy_ab.txt
  FASTDIFF_B200_LIB=$L timeout 120 python tests/gpu_scripts/power_profile.py ${PREFIXES:-1 5 99} 2>&1 | grep -v "Warn\|WeightNorm" >> $OUT/energy_ab.txt
done
