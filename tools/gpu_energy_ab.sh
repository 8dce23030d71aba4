mkdir -p gpurun_out/r2_c19; rm -f gpurun_out/r2_c19/*
for v in default cs shfl hint default2; do
  L=$PWD/fastdiff_b200/csrc/libfd_ab_$v.so; [ -f $L ] || L=$PWD/fastdiff_b200/csrc/libfastdiff_b200.so
  echo "== $v" >> gpurun_out/r2_c19/energy_ab.txt
  FASTDIFF_B200_LIB=$L timeout 120 python tests/gpu_scripts/power_profile.py 1 5 99 2>&1 | grep -v "Warn\|WeightNorm" >> gpurun_out/r2_c19/energy_ab.txt
done
