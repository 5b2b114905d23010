#!/bin/bash
# A/B of library builds on C4: usage  bash tools/r2_ab_c4.sh lib1.so lib2.so ...   ("-" = the in-tree build)
cd "$(dirname "$0")/.."
J='import json,sys; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["value"], d["roofline"]["frac"], d["roofline"].get("frac_best_step"), d.get("parity",{}).get("ok"), d["clocks"]["sm_mhz"])'
for rep in 1 2; do
for L in "$@"; do
  if [ "$L" = "-" ]; then unset SAVGOL_B200_LIB; else export SAVGOL_B200_LIB=$PWD/$L; fi
  for IM in 64 256; do
  echo -n "c4/$IM $L: "; SG_C4_IMAGES=$IM timeout 300 python bench.py --workload c4 --steps 10 --warmup 3 --no-cpu --no-e2e --no-sustained 2>&1 | tail -1 | python -c "$J"
  done
done
done
unset SAVGOL_B200_LIB
