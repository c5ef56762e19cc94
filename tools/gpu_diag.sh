#!/usr/bin/env bash
# Run every GPU test file in its own process (a faulting kernel must not hide the others' results)
# and keep the full logs under gpurun_out/.  Usage on the GPU box:  bash tools/gpu_diag.sh [files...]
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.csv 2>&1
files=("$@")
if [ ${#files[@]} -eq 0 ]; then files=(tests/test_gpu_preprocess.py tests/test_gpu_knn.py tests/test_gpu_project.py); fi
rc_all=0
for f in "${files[@]}"; do
  name=$(basename "$f" .py)
  timeout 600 python -m pytest "$f" -q -m gpu --tb=short -p no:cacheprovider > "gpurun_out/${name}.log" 2>&1
  rc=$?
  echo "$f -> exit $rc" | tee -a gpurun_out/diag_summary.txt
  tail -n 25 "gpurun_out/${name}.log"
  [ $rc -ne 0 ] && rc_all=1
done
exit $rc_all
