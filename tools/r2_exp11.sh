#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
J='import json,sys; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["value"], d["roofline"]["frac"], d["roofline"].get("frac_best_step"), d.get("parity",{}).get("ok"), d["clocks"]["sm_mhz"], (d.get("sustained") or {}).get("frac_hbm"))'
for rep in 1 2; do
for L in - variants/libsavgol_b200_addminb3.so; do
  if [ "$L" = "-" ]; then unset SAVGOL_B200_LIB; else export SAVGOL_B200_LIB=$PWD/$L; fi
  echo -n "c4/64 $L: "; SG_C4_IMAGES=64 timeout 300 python bench.py --workload c4 --steps 20 --warmup 5 --no-cpu --no-e2e 2>&1 | tail -1 | python -c "$J"
done
done
unset SAVGOL_B200_LIB
echo "== c2 / c3 after the rule change"
for W in c2 c3; do timeout 300 python bench.py --workload $W --steps 20 --warmup 5 --no-cpu --no-e2e --no-sustained 2>&1 | tail -1 | python -c "$J"; done
