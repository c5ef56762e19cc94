#!/usr/bin/env bash
# ncu evidence for the stage-1 kernels after the occupancy / band-geometry tuning.
set -u
mkdir -p gpurun_out
run() {  # name, kernel regex, skip, count, command...
  local name=$1 regex=$2 skip=$3 count=$4; shift 4
  "$@" > "gpurun_out/${name}_plain.log" 2>&1 &&
  ncu --set full --clock-control none --import-source on -k "regex:${regex}" -s "${skip}" -c "${count}" -o "gpurun_out/r2c_${name}" "$@" > "gpurun_out/${name}_ncu.log" 2>&1
  echo "${name}: rc=$? $(tail -n 1 gpurun_out/${name}_plain.log | head -c 300)"
}
run pre_nhwc_f32 "stats_u8_stream|apply_u8_lut" 4 2 python tools/run_case.py preprocess --case nhwc_f32 --batch 4096 --iters 1 --warm 2
run pre_resize384 "resize_u8_c3" 4 2 python tools/run_case.py preprocess --case resize384 --batch 2048 --iters 1 --warm 2
