#!/bin/bash
# ncu captures of the kernels added late in round 2: misaligned rows / short tails (generic 1D kernel), short rows
# (packed kernel), gradient / Hessian in one launch (multi-output 2D kernel).   usage: gpurun -- 'bash tools/r2_profile_new.sh'
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
cap() { # name kernel-regex skip command...
  local name=$1 k=$2 s=$3; shift 3
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s $s -c 1 -f -o gpurun_out/r2_prof_$name "$@" > gpurun_out/r2_ncu_$name.log 2>&1
  tail -1 gpurun_out/r2_ncu_$name.log | cut -c1-160
}
cap l4097 sg1d_kernel 3 python tools/run_shape.py 4097
cap l64 sg1d_packed 3 python tools/run_shape.py 64
cap l250 sg1d_packed 3 python tools/run_shape.py 250
cap grad5 sep_kernel 4 python tools/run_wrapper.py 2 2 4096     # launches alternate gradient / hessian: skip 4 -> a gradient launch
cap hess5 sep_kernel 5 python tools/run_wrapper.py 2 2 4096
