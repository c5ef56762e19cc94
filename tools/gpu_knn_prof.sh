#!/bin/bash
# diagnostic build with the search kernel's cycle/event counters (never the shipped library)
mkdir -p gpurun_out
cp imagescry_b200/lib/libimagescry_b200.so /tmp/lib_keep.so
ISX_NVCC_EXTRA="-DISX_KNN_PROFILE $1" python -m imagescry_b200._build --force > gpurun_out/prof_build.log 2>&1
for args in "--d 256 --k 10" "--d 64 --k 10" "--d 256 --k 100"; do
  echo "== $args"
  timeout 300 python tools/run_case.py knn $args --iters 1 --warm 0 2>&1 | grep -v "^$" | tail -3
done > gpurun_out/knn_prof.log 2>&1
cp /tmp/lib_keep.so imagescry_b200/lib/libimagescry_b200.so
cat gpurun_out/knn_prof.log
