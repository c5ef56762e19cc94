#!/bin/bash
# A/B of a search-kernel variant selected by a macro: the shipped library (tests + timings) against a
# rebuild with $1 (nvcc defines, e.g. -DISX_KNN_DEFER=0).  The shipped library is restored afterwards.
mkdir -p gpurun_out
cases=("--d 256 --k 10" "--d 64 --k 10" "--d 128 --k 16" "--d 1280 --k 10")
run_cases() {
  for args in "${cases[@]}"; do
    timeout 300 python tools/run_case.py knn $args --iters 5 --warm 3 2>&1 | tail -1
  done
  timeout 300 python tools/run_case.py graph --n 500000 --d 256 --k 10 --iters 3 --warm 1 2>&1 | tail -1
}
cp imagescry_b200/lib/libimagescry_b200.so /tmp/lib_keep.so
{
  timeout 900 python -m pytest tests/test_gpu_knn.py tests/test_gpu_guard_bands.py -x -q -m gpu 2>&1 | tail -3
  echo "== shipped"; run_cases
  ISX_NVCC_EXTRA="$1" python -m imagescry_b200._build --force > gpurun_out/ab_build.log 2>&1 || { echo "build failed"; tail -20 gpurun_out/ab_build.log; }
  echo "== variant $1"; run_cases
} 2>&1 | tee gpurun_out/knn_ab.log
cp /tmp/lib_keep.so imagescry_b200/lib/libimagescry_b200.so
